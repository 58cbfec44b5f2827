"""No-GPU checks: the C-ABI library builds for sm_100a, loads, and exports every symbol the
header declares; host-side logic (scenario packing, config struct layout, recipes)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from automatedvaletparking_b200 import scenarios as scn
from automatedvaletparking_b200.hostcfg import AvpConfig, AvpPlanSummary, SUMMARY_DTYPE, make_avp_config


def test_library_exports_every_declared_symbol(native_built):
    from automatedvaletparking_b200 import _native
    hdr = open(os.path.join(ROOT, "include", "avp_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(avp_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    L = _native.lib()
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/avp_b200.h but not exported"
    assert set(_native.EXPORTS) == declared


def test_product_has_no_cpu_fallback(native_built):
    """avp_create must fail (not fall back) when no CUDA device is present."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from automatedvaletparking_b200.batch import DevicePlanner, AvpError
    with pytest.raises(AvpError):
        DevicePlanner()


def test_product_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "automatedvaletparking_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                code = "\n".join(l for l in txt.splitlines() if not l.strip().startswith(("#", "//", "*", "/*")))
                assert "oracle_lib" not in code and "libavp_oracle" not in code and "avp_oracle" not in code, f


def test_struct_layouts_match_c(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "avp_b200.h"\nint main(){printf("%zu %zu\\n", sizeof(avp_config), sizeof(avp_plan_summary));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    a, b = map(int, subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split())
    assert a == ctypes.sizeof(AvpConfig) and b == ctypes.sizeof(AvpPlanSummary) == SUMMARY_DTYPE.itemsize


def test_config_matches_reference_constants():
    c = make_avp_config()
    assert c.n_substeps == 3 and c.steering_angle_num == 5
    assert list(c.steer[:5]) == [-0.75, -0.375, 0.0, 0.375, 0.75]
    assert c.min_radius_turn == 3.9765932159382564
    assert c.collision_mode == 0


def test_case_parsing_and_packing():
    s = scn.benchmark_case(1)
    assert len(s.obs) == 3 and all(o.shape == (4, 2) for o in s.obs)
    assert s.pose[0] == -16.0199004975124
    b = scn.pack([scn.benchmark_case(1), scn.benchmark_case(19)])
    assert b.obs_off.tolist() == [0, 3, 3 + len(scn.benchmark_case(19).obs)]
    assert b.vert_off[-1] == b.verts.shape[0] and b.nv.sum() == b.verts.shape[0]
    # CSV round trip is bit-exact (repr(float))
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "x.csv")
        scn.write_case_csv(s, p)
        t = scn.read_case_csv(p)
    assert t.pose == s.pose and all(np.array_equal(a, c) for a, c in zip(t.obs, s.obs))


def test_recipes_are_seeded_and_valid():
    a = scn.perturbed_set(scn.benchmark_case(1), 8, seed=1)
    b = scn.perturbed_set(scn.benchmark_case(1), 8, seed=1)
    assert [x.pose for x in a] == [y.pose for y in b]
    for s in a:
        assert abs(s.x0 - scn.benchmark_case(1).x0) <= 2 and np.hypot(s.x0 - s.xf, s.y0 - s.yf) >= 1
    syn = scn.synthetic_set(1, 3, seed=4)
    assert len(syn) == 3 and len(syn[0].obs) == 256 and syn[0].boundary == (0.0, 20.0, 0.0, 20.0)


def test_sincos_port_is_bit_identical_to_libm(tmp_path):
    """csrc/avp_sincos.h (host build of the same source the device compiles) vs libm."""
    exe = tmp_path / "sincos_bits"
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-mfma", "-o", str(exe),
                    os.path.join(ROOT, "tests", "csrc", "sincos_bits.c"), "-lm"], check=True)
    r = subprocess.run([str(exe), "1500000", "2024"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-500:]
    assert "sin mismatches 0, cos mismatches 0" in r.stdout


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the oracle port on the host cores) prints ONE JSON line with the keys the driver reads."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "successors/s" and d["higher_is_better"] is True
    assert d["metric"] == "collision-checked successor evaluations/sec" and d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("Case1 obstacle map") and "model" not in d["config"]
