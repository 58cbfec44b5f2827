#!/bin/bash
mkdir -p gpurun_out
T=r02s29
L=$PWD/automatedvaletparking_b200
AVP_B200_LIB=$L/libavp_b200_byfam.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "plan_all_benchmark or perturbed_batch" > gpurun_out/${T}_tests.log 2>&1; echo "tests(byfam) rc $? $(tail -2 gpurun_out/${T}_tests.log | tr '\n' ' ')"
for rep in 1 2 3; do
for v in "" _byfam _smo1k; do
  AVP_B200_LIB=$L/libavp_b200$v.so timeout 300 python tools/gpu_sweep.py c2 "" >> gpurun_out/${T}_ab.log 2>&1; echo "variant '$v': $(tail -1 gpurun_out/${T}_ab.log | cut -c1-120 | tr '\n' '|')"
done; done
for v in "" _byfam _smo1k; do
  AVP_B200_LIB=$L/libavp_b200$v.so timeout 300 python tools/gpu_sweep.py c3 "" >> gpurun_out/${T}_ab_c3.log 2>&1; echo "variant '$v': $(tail -1 gpurun_out/${T}_ab_c3.log | cut -c1-200)"
done
