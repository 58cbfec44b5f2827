#!/bin/bash
mkdir -p gpurun_out
T=r02s17
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rs_optimal or expand_pure or plan_all_benchmark or visited or perturbed_batch or synthetic_stress or edge_inputs or scheduling or f_g_h or root_expansion or config2_full" > gpurun_out/${T}_tests.log 2>&1; echo "tests rc $? $(tail -2 gpurun_out/${T}_tests.log | tr '\n' ' ')"
for rep in 1 2 3; do
  for v in "" _sercommit _nosincos _prevlike; do
    AVP_B200_LIB=$PWD/automatedvaletparking_b200/libavp_b200$v.so timeout 300 python tools/gpu_sweep.py c2 "" 2>&1 | sed "s/^c2  /c2 [base$v]/" | cut -c1-200 | tee -a gpurun_out/${T}_ab.log
  done
done
