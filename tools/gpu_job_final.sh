#!/bin/bash
# Final profile job of round 2 (run through gpurun): bench line, reference arm, launch list + DRAM bytes, ncu full capture of the
# wide search kernel on the long searches, light profiles, sanitizer.  TAG names the output files (gpurun_out/TAG_*).
mkdir -p gpurun_out
T=${1:-r02fin}
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench rc $?"; head -c 700 gpurun_out/${T}_bench_n1.json; echo
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref_n1.json 2> gpurun_out/${T}_bench_ref_n1.err; echo "ref rc $?"; head -c 400 gpurun_out/${T}_bench_ref_n1.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${T}_bench_under_ncu.log 2>&1; echo "launch list rc $?"
timeout 900 ncu --set full --clock-control none --import-source on -k k_plan --launch-skip 2 --launch-count 1 -f -o gpurun_out/${T}_kplan_long python tools/profile_long.py 6000 > gpurun_out/${T}_prof_long.log 2>&1; echo "ncu long rc $?"; tail -1 gpurun_out/${T}_prof_long.log
AVP_B200_LIB=$PWD/automatedvaletparking_b200/libavp_b200_lp.so timeout 300 python tools/gpu_light_profile.py c2 > gpurun_out/${T}_light_c2.log 2>&1; echo "light c2 rc $?"; head -3 gpurun_out/${T}_light_c2.log | cut -c1-400
AVP_B200_LIB=$PWD/automatedvaletparking_b200/libavp_b200_lp.so timeout 300 python tools/gpu_light_profile.py c3 > gpurun_out/${T}_light_c3.log 2>&1; echo "light c3 rc $?"; head -3 gpurun_out/${T}_light_c3.log | cut -c1-400
AVP_QUANTUM=8 AVP_FORCE_YIELD=1 AVP_TWO_PHASE=1 AVP_NARROW_BUDGET=6 timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python tools/gpu_sanitize.py > gpurun_out/${T}_memcheck.log 2>&1; echo "memcheck rc $? $(grep -E 'ERROR SUMMARY|sanitize batch' gpurun_out/${T}_memcheck.log | tr '\n' ' ')"
AVP_QUANTUM=8 AVP_FORCE_YIELD=1 AVP_TWO_PHASE=1 AVP_NARROW_BUDGET=6 timeout 400 compute-sanitizer --tool synccheck --print-limit 20 python tools/gpu_sanitize.py > gpurun_out/${T}_synccheck.log 2>&1; echo "synccheck rc $? $(grep -E 'ERROR SUMMARY|sanitize batch' gpurun_out/${T}_synccheck.log | tr '\n' ' ')"
