#!/bin/bash
mkdir -p gpurun_out
T=r02s34
timeout 600 ncu --set full --clock-control none --import-source on -k k_dij_eager --launch-count 1 -f -o gpurun_out/${T}_kdij python tools/profile_run.py 1024 64 > gpurun_out/${T}_prof_kdij.log 2>&1; echo "ncu rc $?"; tail -1 gpurun_out/${T}_prof_kdij.log
