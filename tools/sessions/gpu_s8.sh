#!/bin/bash
mkdir -p gpurun_out
T=r02s8
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "scheduling or f_g_h or root_expansion or config4_full" > gpurun_out/${T}_tests.log 2>&1; echo "tests rc $? $(tail -2 gpurun_out/${T}_tests.log | tr '\n' ' ')"
for rep in 1 2; do
  for v in "" _pf0 _pf2 _pf3 _smo8k; do
    AVP_B200_LIB=$PWD/automatedvaletparking_b200/libavp_b200$v.so timeout 300 python tools/gpu_sweep.py c2 "" 2>&1 | sed "s/^c2  /c2 [base$v]/" | cut -c1-200 | tee -a gpurun_out/${T}_ab.log
  done
done
( time timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err ) 2>&1 | grep real; echo "bench rc $?"; head -c 3000 gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config3_full" > gpurun_out/${T}_tests_c3full.log 2>&1; echo "tests c3 full rc $? $(tail -2 gpurun_out/${T}_tests_c3full.log | tr '\n' ' ')"
