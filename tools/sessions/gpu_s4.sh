#!/bin/bash
mkdir -p gpurun_out
T=r02s4
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rasters or plan_all_benchmark or edge_inputs" > gpurun_out/${T}_tests.log 2>&1; echo "tests rc $? $(tail -2 gpurun_out/${T}_tests.log | tr '\n' ' ')"
AVP_B200_LIB=$PWD/automatedvaletparking_b200/libavp_b200_lp.so timeout 300 python tools/gpu_light_profile.py c2 > gpurun_out/${T}_light.log 2>&1; echo "light rc $?"; cat gpurun_out/${T}_light.log | cut -c1-250
timeout 600 python tools/gpu_sweep.py c2 "" AVP_PLAN_BLOCK=640 "AVP_SPREAD_MAX=74" "AVP_PLAN_BLOCK=640 AVP_SPREAD_MAX=74" "" > gpurun_out/${T}_sweep_c2.log 2>&1; echo "sweep c2 rc $?"; cat gpurun_out/${T}_sweep_c2.log | cut -c1-260
AVP_PLAN_BLOCK=640 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "plan_all_benchmark or config2_full" > gpurun_out/${T}_tests_b640.log 2>&1; echo "tests(block 640) rc $? $(tail -2 gpurun_out/${T}_tests_b640.log | tr '\n' ' ')"
