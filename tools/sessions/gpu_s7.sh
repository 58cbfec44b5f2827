#!/bin/bash
mkdir -p gpurun_out
T=r02s7
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "scheduling or f_g_h or root_expansion or config4_full" > gpurun_out/${T}_tests.log 2>&1; echo "tests rc $? $(tail -2 gpurun_out/${T}_tests.log | tr '\n' ' ')"
for rep in 1 2; do
  for v in "" _sl100 _sl400; do
    AVP_B200_LIB=$PWD/automatedvaletparking_b200/libavp_b200$v.so timeout 300 python tools/gpu_sweep.py c2 "" 2>&1 | sed "s/^c2  /c2 [base$v]/" | cut -c1-200 | tee -a gpurun_out/${T}_ab.log
  done
done
AVP_B200_LIB=$PWD/automatedvaletparking_b200/libavp_b200_lp.so timeout 600 python tools/gpu_light_profile.py c3 > gpurun_out/${T}_light_c3.log 2>&1; echo "light c3 rc $?"; cat gpurun_out/${T}_light_c3.log | cut -c1-250
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config3_full" > gpurun_out/${T}_tests_c3full.log 2>&1; echo "tests c3 full rc $? $(tail -2 gpurun_out/${T}_tests_c3full.log | tr '\n' ' ')"
