#!/bin/bash
# build a variant of the library next to the default one: tools/build_variant.sh <suffix> [extra nvcc flags...]
#   tools/build_variant.sh prof -DAVP_PROFILE      -> automatedvaletparking_b200/libavp_b200_prof.so (cycle counters compiled in)
# select it with AVP_B200_LIB=$PWD/automatedvaletparking_b200/libavp_b200_<suffix>.so (tools/gpu_pipe_profile.py, A/B runs)
set -e
cd "$(dirname "$0")/.."
S=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off -shared "$@" \
  -o automatedvaletparking_b200/libavp_b200_$S.so automatedvaletparking_b200/csrc/avp_api.cu
echo built automatedvaletparking_b200/libavp_b200_$S.so
