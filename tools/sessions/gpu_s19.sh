#!/bin/bash
mkdir -p gpurun_out
T=r02s19
AVP_B200_LIB=$PWD/automatedvaletparking_b200/libavp_b200_lp.so timeout 300 python tools/gpu_light_profile.py c2 > gpurun_out/${T}_light.log 2>&1; echo "light rc $?"; cat gpurun_out/${T}_light.log | cut -c1-400
AVP_SPREAD=0 AVP_B200_LIB=$PWD/automatedvaletparking_b200/libavp_b200_lp.so timeout 300 python tools/gpu_light_profile.py c2 > gpurun_out/${T}_light_nospread.log 2>&1; echo "light(nospread) rc $?"; tail -5 gpurun_out/${T}_light_nospread.log | cut -c1-400
