#!/bin/bash
mkdir -p gpurun_out
T=r02fin
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err; echo "n2 rc $?"; tail -c 1500 gpurun_out/${T}_bench_n2.json | head -c 900; echo
