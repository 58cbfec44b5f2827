#!/bin/bash
mkdir -p gpurun_out
T=r02fin4
timeout 280 python -m pytest tests -m gpu -x -q -k "not config3_full_size and not config4_full_size and not config3_all_cases" > gpurun_out/${T}_tests.log 2>&1; echo "tests rc $? $(tail -2 gpurun_out/${T}_tests.log | tr '\n' ' ')"
timeout 60 python __graft_entry__.py --smoke 2>&1 | tail -1
