#!/bin/bash
mkdir -p gpurun_out
T=r02s28
L=$PWD/automatedvaletparking_b200
for rep in 1 2 3; do
for v in "" _selfine; do
  AVP_B200_LIB=$L/libavp_b200$v.so timeout 300 python tools/gpu_sweep.py c2 "" AVP_NARROW_BUDGET=256 >> gpurun_out/${T}_ab.log 2>&1; echo "variant '$v': $(tail -2 gpurun_out/${T}_ab.log | cut -c1-120 | tr '\n' '|')"
done; done
timeout 600 python tools/gpu_sweep.py c2 "" AVP_NARROW_BUDGET=192 AVP_NARROW_BUDGET=256 AVP_NARROW_BUDGET=320 "AVP_NARROW_BUDGET=256 AVP_QUANTUM=256" "AVP_NARROW_BUDGET=256 AVP_QUANTUM=1024" AVP_SPREAD_MAX=90 AVP_SPREAD_MAX=130 > gpurun_out/${T}_sweep_c2.log 2>&1; cat gpurun_out/${T}_sweep_c2.log | cut -c1-150
for v in "" _selfine; do
  AVP_B200_LIB=$L/libavp_b200$v.so timeout 300 python tools/gpu_sweep.py c3 "" >> gpurun_out/${T}_ab_c3.log 2>&1; echo "variant '$v': $(tail -1 gpurun_out/${T}_ab_c3.log | cut -c1-200)"
done
