#!/bin/bash
mkdir -p gpurun_out
T=r02s5
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_split.py -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc $? $(tail -2 gpurun_out/${T}_tests.log | tr '\n' ' ')"
AVP_B200_LIB=$PWD/automatedvaletparking_b200/libavp_b200_fine.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "plan_all_benchmark or config2_full" > gpurun_out/${T}_tests_fine.log 2>&1; echo "tests(fine) rc $? $(tail -2 gpurun_out/${T}_tests_fine.log | tr '\n' ' ')"
for rep in 1 2; do
  for v in "" _nopf _nostage _fine; do
    AVP_B200_LIB=$PWD/automatedvaletparking_b200/libavp_b200$v.so timeout 300 python tools/gpu_sweep.py c2 "" AVP_PLAN_BLOCK=640 2>&1 | sed "s/^c2  /c2 [base$v]/" | cut -c1-200 | tee -a gpurun_out/${T}_ab.log
  done
done
