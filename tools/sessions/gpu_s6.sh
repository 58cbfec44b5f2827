#!/bin/bash
mkdir -p gpurun_out
T=r02s6
for rep in 1 2; do
  for v in "" _sl100 _sl400; do
    AVP_B200_LIB=$PWD/automatedvaletparking_b200/libavp_b200$v.so timeout 300 python tools/gpu_sweep.py c2 "" 2>&1 | sed "s/^c2  /c2 [base$v]/" | cut -c1-200 | tee -a gpurun_out/${T}_ab.log
  done
done
AVP_B200_LIB=$PWD/automatedvaletparking_b200/libavp_b200_lp.so timeout 600 python tools/gpu_light_profile.py c3 > gpurun_out/${T}_light_c3.log 2>&1; echo "light c3 rc $?"; cat gpurun_out/${T}_light_c3.log | cut -c1-250
AVP_QUANTUM=8 timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/gpu_sanitize.py > gpurun_out/${T}_memcheck.log 2>&1; echo "memcheck rc $? $(grep -E 'ERROR SUMMARY|sanitize batch' gpurun_out/${T}_memcheck.log | tr '\n' ' ')"
AVP_QUANTUM=8 timeout 600 compute-sanitizer --tool synccheck --print-limit 20 python tools/gpu_sanitize.py > gpurun_out/${T}_synccheck.log 2>&1; echo "synccheck rc $? $(grep -E 'ERROR SUMMARY|sanitize batch' gpurun_out/${T}_synccheck.log | tr '\n' ' ')"
AVP_QUANTUM=8 timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 200 python tools/gpu_sanitize.py 1 > gpurun_out/${T}_racecheck.log 2>&1; echo "racecheck rc $? $(grep -E 'RACECHECK SUMMARY|sanitize batch' gpurun_out/${T}_racecheck.log | tr '\n' ' ')"
grep -c "Race reported" gpurun_out/${T}_racecheck.log
