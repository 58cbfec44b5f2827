#!/bin/bash
mkdir -p gpurun_out
T=r02s16
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "plan_all_benchmark or visited or perturbed_batch or synthetic_stress or edge_inputs or scheduling or f_g_h or root_expansion or config2_full or config4_full" > gpurun_out/${T}_tests.log 2>&1; echo "tests rc $? $(tail -2 gpurun_out/${T}_tests.log | tr '\n' ' ')"
for rep in 1 2; do
  for v in "" _sercommit _nosubgeom _prevlike; do
    AVP_B200_LIB=$PWD/automatedvaletparking_b200/libavp_b200$v.so timeout 300 python tools/gpu_sweep.py c2 "" AVP_PLAN_BLOCK=640 2>&1 | sed "s/^c2  /c2 [base$v]/" | cut -c1-200 | tee -a gpurun_out/${T}_ab.log
  done
done
AVP_B200_LIB=$PWD/automatedvaletparking_b200/libavp_b200_lp.so timeout 300 python tools/gpu_light_profile.py c2 > gpurun_out/${T}_light.log 2>&1; echo "light rc $?"; cat gpurun_out/${T}_light.log | cut -c1-330
