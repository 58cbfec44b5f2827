#!/bin/bash
mkdir -p gpurun_out
T=r02s20
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "plan_all_benchmark or perturbed_batch or synthetic_stress or edge_inputs or scheduling or config2_full or config4_full" > gpurun_out/${T}_tests.log 2>&1; echo "tests rc $? $(tail -2 gpurun_out/${T}_tests.log | tr '\n' ' ')"
timeout 900 python tools/gpu_sweep.py c2 "" AVP_TWO_PHASE=0 AVP_NARROW_BUDGET=16 AVP_NARROW_BUDGET=32 AVP_NARROW_BUDGET=128 AVP_NARROW_BUDGET=256 "" AVP_TWO_PHASE=0 > gpurun_out/${T}_sweep_c2.log 2>&1; cat gpurun_out/${T}_sweep_c2.log | cut -c1-215
AVP_B200_LIB=$PWD/automatedvaletparking_b200/libavp_b200_lp.so timeout 300 python tools/gpu_light_profile.py c2 > gpurun_out/${T}_light.log 2>&1; echo "light rc $?"; tail -5 gpurun_out/${T}_light.log | cut -c1-400
timeout 900 python tools/gpu_sweep.py c3 "" AVP_TWO_PHASE=0 > gpurun_out/${T}_sweep_c3.log 2>&1; cat gpurun_out/${T}_sweep_c3.log | cut -c1-215
timeout 900 python tools/gpu_sweep.py c4 "" AVP_TWO_PHASE=0 > gpurun_out/${T}_sweep_c4.log 2>&1; cat gpurun_out/${T}_sweep_c4.log | cut -c1-215
