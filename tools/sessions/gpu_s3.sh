#!/bin/bash
mkdir -p gpurun_out
T=r02s3
timeout 900 python -m pytest tests/test_gpu_split.py tests/test_gpu_dropin.py -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc $? $(tail -3 gpurun_out/${T}_tests.log | tr '\n' ' ')"
AVP_B200_LIB=$PWD/automatedvaletparking_b200/libavp_b200_prof.so AVP_TRACE_POP=5000 timeout 300 python tools/gpu_pipe_profile.py > gpurun_out/${T}_pipe_phase.log 2>&1; echo "phase rc $?"; head -16 gpurun_out/${T}_pipe_phase.log | cut -c1-200; tail -3 gpurun_out/${T}_pipe_phase.log | cut -c1-400
timeout 1500 python tools/gpu_sweep.py c4s "" AVP_PLAN_BLOCK=256 > gpurun_out/${T}_sweep_c4s.log 2>&1; echo "sweep c4s rc $?"; cat gpurun_out/${T}_sweep_c4s.log | cut -c1-260
