#!/bin/bash
mkdir -p gpurun_out
T=r02s10
AVP_B200_LIB=$PWD/automatedvaletparking_b200/libavp_b200_lp.so timeout 300 python tools/gpu_light_profile.py c2 > gpurun_out/${T}_light.log 2>&1; echo "light rc $?"; cat gpurun_out/${T}_light.log | cut -c1-330
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -m gpu -x -q -k "scheduling or dropin or stepwise or leaf or case20 or path_planning" > gpurun_out/${T}_tests.log 2>&1; echo "tests rc $? $(tail -2 gpurun_out/${T}_tests.log | tr '\n' ' ')"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_plan --launch-count 1 -f -o gpurun_out/${T}_kplan python tools/profile_run.py 1024 4000 > gpurun_out/${T}_prof_kplan.log 2>&1; echo "ncu kplan rc $?"; tail -2 gpurun_out/${T}_prof_kplan.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dij_eager --launch-count 1 -f -o gpurun_out/${T}_kdij python tools/profile_run.py 1024 4000 > gpurun_out/${T}_prof_kdij.log 2>&1; echo "ncu kdij rc $?"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${T}_bench_under_ncu.log 2>&1; echo "launch list rc $?"
ls -la gpurun_out | grep ${T}
