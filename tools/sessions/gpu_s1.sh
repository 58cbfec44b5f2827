#!/bin/bash
# session 1 of round 2: parity of the new single-launch search, then scheduling sweeps on the three workloads
mkdir -p gpurun_out
T=r02s1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc $? $(tail -3 gpurun_out/${T}_tests.log | tr '\n' ' ')"
AVP_QUANTUM=5 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "plan_all_benchmark or perturbed_batch or synthetic_stress or edge_inputs" > gpurun_out/${T}_tests_q5.log 2>&1; echo "tests(q=5) rc $? $(tail -2 gpurun_out/${T}_tests_q5.log | tr '\n' ' ')"
AVP_QUANTUM=16 AVP_SLOTS=150 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config2_full or config4" > gpurun_out/${T}_tests_slots.log 2>&1; echo "tests(q=16, 150 slots) rc $? $(tail -2 gpurun_out/${T}_tests_slots.log | tr '\n' ' ')"
AVP_PLAN_BLOCK=256 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "plan_all_benchmark or config2_full" > gpurun_out/${T}_tests_b256.log 2>&1; echo "tests(block 256) rc $? $(tail -2 gpurun_out/${T}_tests_b256.log | tr '\n' ' ')"
timeout 600 python tools/gpu_sweep.py c2 "" AVP_QUANTUM=128 AVP_QUANTUM=256 AVP_QUANTUM=1024 AVP_QUANTUM=4096 AVP_SPREAD=0 "AVP_SPREAD_MAX=90" "AVP_SPREAD_MAX=130" AVP_PLAN_BLOCK=256 "AVP_PLAN_BLOCK=256 AVP_SPREAD=0" "" > gpurun_out/${T}_sweep_c2.log 2>&1; echo "sweep c2 rc $?"; cat gpurun_out/${T}_sweep_c2.log | cut -c1-260
timeout 900 python tools/gpu_sweep.py c4 "" AVP_PLAN_BLOCK=256 AVP_SPREAD=0 "AVP_PLAN_BLOCK=256 AVP_SPREAD=0" AVP_QUANTUM=2048 > gpurun_out/${T}_sweep_c4.log 2>&1; echo "sweep c4 rc $?"; cat gpurun_out/${T}_sweep_c4.log | cut -c1-260
timeout 1500 python tools/gpu_sweep.py c3 "" AVP_PLAN_BLOCK=256 > gpurun_out/${T}_sweep_c3.log 2>&1; echo "sweep c3 rc $?"; cat gpurun_out/${T}_sweep_c3.log | cut -c1-260
