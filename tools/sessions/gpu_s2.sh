#!/bin/bash
mkdir -p gpurun_out
T=r02s2
AVP_QUANTUM=16 AVP_SLOTS=150 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config2_full or config4" > gpurun_out/${T}_tests_slots.log 2>&1; echo "tests(q=16, 150 slots) rc $? $(tail -2 gpurun_out/${T}_tests_slots.log | tr '\n' ' ')"
AVP_B200_LIB=$PWD/automatedvaletparking_b200/libavp_b200_prof.so AVP_TRACE_POP=5000 timeout 300 python tools/gpu_pipe_profile.py > gpurun_out/${T}_pipe_phase.log 2>&1; echo "phase rc $?"; head -34 gpurun_out/${T}_pipe_phase.log | cut -c1-200
timeout 900 python tools/gpu_sweep.py c4 "" AVP_PLAN_BLOCK=256 > gpurun_out/${T}_sweep_c4.log 2>&1; echo "sweep c4 rc $?"; cat gpurun_out/${T}_sweep_c4.log | cut -c1-260
timeout 1500 python tools/gpu_sweep.py c3 "" AVP_PLAN_BLOCK=256 "AVP_SPREAD=0" > gpurun_out/${T}_sweep_c3.log 2>&1; echo "sweep c3 rc $?"; cat gpurun_out/${T}_sweep_c3.log | cut -c1-260
timeout 1500 python tools/gpu_sweep.py c4s "" AVP_PLAN_BLOCK=256 > gpurun_out/${T}_sweep_c4s.log 2>&1; echo "sweep c4s rc $?"; cat gpurun_out/${T}_sweep_c4s.log | cut -c1-260
