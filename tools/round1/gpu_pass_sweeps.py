"""Round-2 sweeps on the bench workload (run through gpurun): pass-1 occupancy, pass-1 pop budget, SM-pair threshold.
Every configuration is a fresh process (the switches are read at avp_create / plan time):
    python tools/gpu_pass_sweeps.py            # prints one line per configuration: search ms, (pass 1, pass 2, pending, CTA width)
"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUN = ("import sys, os; sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, 'tests'));"
       "import bench; from automatedvaletparking_b200.batch import DevicePlanner;"
       "dp = DevicePlanner(max_pops=20000); scs = bench.make_scenarios(int(os.environ.get('SWEEP_RANK', '0')), 1024, dp); dp.load(scs);"
       "dp.plan_resident(256, 0); ms = dp.plan_resident(256, 0); print('%%.1f' %% ms, dp.last_search_passes())") % (ROOT, ROOT)
configs = [{}] + [{"AVP_NARROW_PER_SM": str(k)} for k in (1, 2, 3)] + [{"AVP_POP_BUDGET": str(b)} for b in (512, 256, 128)] + \
          [{"AVP_POP_BUDGET": "256", "AVP_SPREAD_MAX": "148"}, {"SWEEP_RANK": "1"}, {"SWEEP_RANK": "1", "AVP_SPREAD": "0"}, {"SWEEP_RANK": "7"}, {"SWEEP_RANK": "7", "AVP_SPREAD": "0"}]
for c in configs:
    env = dict(os.environ, AVP_HOST_TIMEOUT_S="120", **c)
    out = subprocess.run([sys.executable, "-c", RUN], capture_output=True, text=True, env=env, timeout=400)
    print(c or "default", "->", out.stdout.strip() or out.stderr.strip()[-300:], flush=True)
