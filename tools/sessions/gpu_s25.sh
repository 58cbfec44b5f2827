#!/bin/bash
mkdir -p gpurun_out
T=r02s25
L=$PWD/automatedvaletparking_b200
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not config3_full_size and not config4_full_size" > gpurun_out/${T}_tests.log 2>&1; echo "tests rc $? $(tail -2 gpurun_out/${T}_tests.log | tr '\n' ' ')"
for rep in 1 2; do
for v in "" _head _sellate; do
  AVP_B200_LIB=$L/libavp_b200$v.so timeout 300 python tools/gpu_sweep.py c2 "" >> gpurun_out/${T}_ab.log 2>&1; echo "variant '$v': $(tail -1 gpurun_out/${T}_ab.log | cut -c1-200)"
done; done
AVP_B200_LIB=$L/libavp_b200_lp.so timeout 300 python tools/gpu_light_profile.py c2 > gpurun_out/${T}_light.log 2>&1; echo "light rc $?"; head -3 gpurun_out/${T}_light.log | cut -c1-400; tail -1 gpurun_out/${T}_light.log
for v in "" _head; do
  AVP_B200_LIB=$L/libavp_b200$v.so timeout 300 python tools/gpu_sweep.py c3 "" >> gpurun_out/${T}_ab_c3.log 2>&1; echo "variant '$v': $(tail -1 gpurun_out/${T}_ab_c3.log | cut -c1-200)"
done
