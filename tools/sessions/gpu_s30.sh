#!/bin/bash
mkdir -p gpurun_out
T=r02s30
L=$PWD/automatedvaletparking_b200
( time timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1 ) 2>&1 | grep real; echo "tests rc $(tail -1 gpurun_out/${T}_tests.log)"
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
for rep in 1 2 3; do
for v in "" _qserial; do
  AVP_B200_LIB=$L/libavp_b200$v.so timeout 300 python tools/gpu_sweep.py c2 "" >> gpurun_out/${T}_ab.log 2>&1; echo "variant '$v': $(tail -1 gpurun_out/${T}_ab.log | cut -c1-120 | tr '\n' '|')"
done; done
for v in "" _qserial; do
  AVP_B200_LIB=$L/libavp_b200$v.so timeout 300 python tools/gpu_sweep.py c3 "" >> gpurun_out/${T}_ab_c3.log 2>&1; echo "variant '$v': $(tail -1 gpurun_out/${T}_ab_c3.log | cut -c1-200)"
done
