#!/bin/bash
mkdir -p gpurun_out
T=r02fin3
( time timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1 ) 2>&1 | grep real; echo "tests rc $(tail -1 gpurun_out/${T}_tests.log)"
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench rc $?"; head -c 300 gpurun_out/${T}_bench_n1.json; echo
AVP_QUANTUM=8 AVP_FORCE_YIELD=1 AVP_TWO_PHASE=1 AVP_NARROW_BUDGET=6 timeout 420 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 200 python tools/gpu_sanitize.py 1 > gpurun_out/${T}_racecheck.log 2>&1; echo "racecheck rc $? $(grep -E 'RACECHECK SUMMARY|sanitize batch' gpurun_out/${T}_racecheck.log | tr '\n' ' ')"
