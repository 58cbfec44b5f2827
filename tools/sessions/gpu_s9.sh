#!/bin/bash
mkdir -p gpurun_out
T=r02s9
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "scheduling or root_expansion" > gpurun_out/${T}_tests.log 2>&1; echo "tests rc $? $(tail -2 gpurun_out/${T}_tests.log | tr '\n' ' ')"
timeout 900 python tools/gpu_sweep.py c2 "" AVP_OVERFLOW_ODD=0 AVP_QUANTUM=128 "AVP_QUANTUM=128 AVP_OVERFLOW_ODD=0" AVP_QUANTUM=256 "AVP_QUANTUM=256 AVP_OVERFLOW_ODD=0" "" AVP_OVERFLOW_ODD=0 "AVP_QUANTUM=128 AVP_SPREAD_MAX=148" > gpurun_out/${T}_sweep_c2.log 2>&1; echo "sweep rc $?"; cat gpurun_out/${T}_sweep_c2.log | cut -c1-200
SWEEP_RANK=3 timeout 900 python tools/gpu_sweep.py c2 "" AVP_OVERFLOW_ODD=0 AVP_QUANTUM=128 "AVP_QUANTUM=128 AVP_OVERFLOW_ODD=0" > gpurun_out/${T}_sweep_c2_rank3.log 2>&1; echo "sweep rank3 rc $?"; cat gpurun_out/${T}_sweep_c2_rank3.log | cut -c1-200
