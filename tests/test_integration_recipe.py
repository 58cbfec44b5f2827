"""INTEGRATION.md §1, executed: the shim directory in front of the reference root on sys.path, then the
reference's own import block (main.py:13-21).  CPU only (imports, no compute).  The part that needs
/root/reference is skipped where the reference is absent (the GPU box)."""
import os
import subprocess
import sys
import textwrap

import pytest

from conftest import ROOT

SHIM = os.path.join(ROOT, "automatedvaletparking_b200", "dropin")
REF = "/root/reference"

PRELUDE = textwrap.dedent("""
    import sys, types
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.animation", "matplotlib.patches", "shapely", "shapely.geometry",
                 "cvxopt", "pyomo", "pyomo.environ", "pyomo.dae", "imageio", "casadi"):
        m = types.ModuleType(name); m.__path__ = []; sys.modules.setdefault(name, m)
    sys.modules["matplotlib.pyplot"].grid = lambda *a, **k: None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    for n in ("matrix", "solvers"):
        setattr(sys.modules["cvxopt"], n, object())
    class _Any(types.ModuleType):
        def __getattr__(self, k):
            if k.startswith("__"): raise AttributeError(k)
            return type(k, (), {})
    for name in ("pyomo.environ", "pyomo.dae", "matplotlib.patches", "matplotlib.animation", "shapely.geometry", "matplotlib.pyplot"):
        a = _Any(name); a.__dict__.update(sys.modules[name].__dict__); sys.modules[name] = a
""")


def _run(code):
    return subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)


def test_shim_directory_holds_only_the_hot_path_packages():
    names = sorted(d for d in os.listdir(SHIM) if os.path.isdir(os.path.join(SHIM, d)) and not d.startswith("__"))
    assert names == ["collision_check", "config", "map", "path_plan"]


def test_shim_modules_resolve_to_the_package_classes():
    code = PRELUDE + textwrap.dedent(f"""
        sys.path.insert(0, {ROOT!r})
        sys.path.insert(0, {SHIM!r})
        from path_plan import path_planner
        from map import costmap
        from collision_check import collision_check
        from config import read_config
        import automatedvaletparking_b200.path_plan.path_planner as real
        assert path_planner.PathPlanner is real.PathPlanner
        assert costmap.Map.__module__ == "automatedvaletparking_b200.map.costmap"
        assert collision_check.distance_checker.__module__.startswith("automatedvaletparking_b200.")
        assert read_config.read_config("config")["map_discrete_size"] == 0.1
        print("ok")
    """)
    r = _run(code)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is only present in the build container")
def test_reference_main_import_block_under_the_recipe():
    lines = open(os.path.join(REF, "main.py"), encoding="utf-8").read().splitlines()
    block = "\n".join(ln for ln in lines[:24] if ln.startswith(("from ", "import ")))
    assert "from path_plan import path_planner" in block and "from optimization import" in block
    code = PRELUDE + textwrap.dedent(f"""
        sys.path.insert(0, {ROOT!r})                 # the package itself
        sys.path.insert(0, {REF!r})                  # downstream stages, main.py
        sys.path.insert(0, {SHIM!r})                 # map/, collision_check/, path_plan/, config/
    """) + block + textwrap.dedent(f"""
        import automatedvaletparking_b200.path_plan.path_planner as real
        assert path_planner.PathPlanner is real.PathPlanner
        assert costmap.Map.__module__ == "automatedvaletparking_b200.map.costmap"
        for mod in (path_optimazition, ocp_optimization, velocity_planner, path_interpolation, sys.modules["animation.animation"],
                    sys.modules["animation.record_solution"], sys.modules["animation.curve_plot"]):
            assert mod.__file__.startswith({REF!r}), mod.__file__
        assert hasattr(path_optimazition.path_opti, "get_result") and hasattr(ocp_optimization.ocp_optimization, "solution")
        print("ok")
    """)
    r = _run(code)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), (r.stdout[-500:], r.stderr[-3000:])
