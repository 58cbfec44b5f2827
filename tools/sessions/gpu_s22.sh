#!/bin/bash
mkdir -p gpurun_out
T=r02s22
timeout 300 python tools/profile_long.py 6000 > gpurun_out/${T}_long_plain.log 2>&1; echo "plain rc $?"; tail -2 gpurun_out/${T}_long_plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k k_plan --launch-skip 2 --launch-count 1 -f -o gpurun_out/${T}_kplan_long python tools/profile_long.py 6000 > gpurun_out/${T}_prof_long.log 2>&1; echo "ncu rc $?"; tail -2 gpurun_out/${T}_prof_long.log
ls -la gpurun_out/${T}*
