#!/bin/bash
mkdir -p gpurun_out
T=r02s15
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 600 python tools/gpu_sweep.py c2 "" AVP_PLAN_BLOCK=640 "" AVP_PLAN_BLOCK=640 > gpurun_out/${T}_sweep_c2.log 2>&1; cat gpurun_out/${T}_sweep_c2.log | cut -c1-200
timeout 900 python bench.py --workload c3 --steps 1 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/${T}_c3_n1.json 2> gpurun_out/${T}_c3_n1.err; echo "c3 n1 rc $?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload c3 --steps 1 --warmup 3 --no-extra > gpurun_out/${T}_c3_n2.json 2> gpurun_out/${T}_c3_n2.err; echo "c3 n2 rc $?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/${T}_n2.json 2> gpurun_out/${T}_n2.err; echo "default n2 rc $?"
python - <<'PY'
import json
for f in ("c3_n1", "c3_n2", "n2"):
    try:
        d = json.loads(open(f"gpurun_out/r02s15_{f}.json").read().strip().splitlines()[-1])
        c = d["config"]
        print(f, "value %.3e e2e %.3e ms %.1f scaling %s checksum %s per-rank %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["scaling"], c["records_checksum"], [round(x, 1) for x in c["ms_per_step_per_rank"]]))
        for k, w in d.get("workloads", {}).items():
            print("   ", k, "value %.3e ms %.1f checksum %s per-rank %s" % (w["value"], w["ms_per_step"], w["config"]["records_checksum"], [round(x, 1) for x in w["config"]["ms_per_step_per_rank"]]))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/${T}_c3_n2.err gpurun_out/${T}_n2.err
