#!/bin/bash
mkdir -p gpurun_out
T=r02s11
( time timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1 ) 2>&1 | grep real; echo "tests rc $(tail -1 gpurun_out/${T}_tests.log)"
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 900 ncu --set full --clock-control none --import-source on -k k_plan --launch-count 1 -f -o gpurun_out/${T}_kplan python tools/profile_run.py 1024 4000 > gpurun_out/${T}_prof_kplan.log 2>&1; echo "ncu kplan rc $?"; tail -1 gpurun_out/${T}_prof_kplan.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "ref rc $?"; head -c 600 gpurun_out/${T}_bench_ref.json
AVP_QUANTUM=8 AVP_FORCE_YIELD=1 timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 200 python tools/gpu_sanitize.py 1 > gpurun_out/${T}_racecheck.log 2>&1; echo "racecheck rc $? $(grep -E 'RACECHECK SUMMARY|sanitize batch' gpurun_out/${T}_racecheck.log | tr '\n' ' ')"
AVP_QUANTUM=8 AVP_FORCE_YIELD=1 timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/gpu_sanitize.py > gpurun_out/${T}_memcheck.log 2>&1; echo "memcheck rc $? $(grep -E 'ERROR SUMMARY|sanitize batch' gpurun_out/${T}_memcheck.log | tr '\n' ' ')"
AVP_QUANTUM=8 AVP_FORCE_YIELD=1 timeout 600 compute-sanitizer --tool synccheck --print-limit 20 python tools/gpu_sanitize.py > gpurun_out/${T}_synccheck.log 2>&1; echo "synccheck rc $? $(grep -E 'ERROR SUMMARY|sanitize batch' gpurun_out/${T}_synccheck.log | tr '\n' ' ')"
