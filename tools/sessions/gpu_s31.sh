#!/bin/bash
mkdir -p gpurun_out
T=r02s31
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "plan_all_benchmark or perturbed_batch or scheduling or config2_full or config4_synthetic or edge_inputs or visited_grid" > gpurun_out/${T}_tests.log 2>&1; echo "tests rc $? $(tail -2 gpurun_out/${T}_tests.log | tr '\n' ' ')"
timeout 600 python tools/gpu_sweep.py c2 "" AVP_FUSE_EAGER=0 "" AVP_FUSE_EAGER=0 "" AVP_FUSE_EAGER=0 > gpurun_out/${T}_sweep_c2.log 2>&1; cat gpurun_out/${T}_sweep_c2.log | cut -c1-215
