# A/B job (run through gpurun): parity tests on the working-tree library, then the pipelined kernel's phase profile
# for the working-tree library and for every other build named on the command line (AVP_B200_LIB), alternating.
#   bash tools/gpu_job_ab.sh TAG [other.so ...]
mkdir -p gpurun_out
T=${1:-r}; shift
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc $?"; tail -3 gpurun_out/${T}_tests.log
for i in 1 2; do
  AVP_TRACE_POP=5000 timeout 300 python tools/gpu_pipe_profile.py > gpurun_out/${T}_new_$i.log 2>&1; echo "new rc $? $(head -1 gpurun_out/${T}_new_$i.log)"
  for L in "$@"; do
    N=$(basename $L .so)
    AVP_B200_LIB=$PWD/$L AVP_TRACE_POP=5000 timeout 300 python tools/gpu_pipe_profile.py > gpurun_out/${T}_${N}_$i.log 2>&1; echo "$N rc $? $(head -1 gpurun_out/${T}_${N}_$i.log)"
  done
done
sed -n 2,14p gpurun_out/${T}_new_2.log; sed -n 27,45p gpurun_out/${T}_new_2.log
