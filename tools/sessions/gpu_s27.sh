#!/bin/bash
mkdir -p gpurun_out
T=r02s27
L=$PWD/automatedvaletparking_b200
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "plan_all_benchmark or perturbed_batch or scheduling or config2_full or edge_inputs" > gpurun_out/${T}_tests.log 2>&1; echo "tests rc $? $(tail -2 gpurun_out/${T}_tests.log | tr '\n' ' ')"
timeout 600 python tools/gpu_sweep.py c2 "" AVP_NARROW_BLOCK=128 AVP_NARROW_BUDGET=32 AVP_NARROW_BUDGET=64 AVP_NARROW_BUDGET=256 AVP_NARROW_BUDGET=512 "" AVP_NARROW_BLOCK=128 > gpurun_out/${T}_sweep_c2.log 2>&1; cat gpurun_out/${T}_sweep_c2.log | cut -c1-215
timeout 900 python tools/gpu_sweep.py c3 "" AVP_NARROW_BLOCK=128 AVP_NARROW_BUDGET=64 AVP_NARROW_BUDGET=256 > gpurun_out/${T}_sweep_c3.log 2>&1; cat gpurun_out/${T}_sweep_c3.log | cut -c1-215
timeout 300 python tools/gpu_sweep.py c4 "" AVP_NARROW_BLOCK=128 > gpurun_out/${T}_sweep_c4.log 2>&1; cat gpurun_out/${T}_sweep_c4.log | cut -c1-215
