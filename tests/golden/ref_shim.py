"""Import the UNMODIFIED reference modules from /root/reference (read-only).

The reference's hot-path modules import matplotlib / shapely at module scope but
never call them unless `draw_collision: True` (path_planner.py:81-84), so empty
stub modules are sufficient.  Used only by the fixture
generators in this directory; nothing under tests/ that runs on the GPU box
imports this file (/root/reference does not exist there).
"""
import sys
import types

REF_ROOT = "/root/reference"


def install():
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.animation",
                 "matplotlib.patches", "shapely", "shapely.geometry"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []  # behave like a package
            sys.modules[name] = m
    # `from matplotlib.pyplot import grid` (compute_h.py:13)
    sys.modules["matplotlib.pyplot"].grid = lambda *a, **k: None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].animation = sys.modules["matplotlib.animation"]
    sys.modules["shapely"].geometry = sys.modules["shapely.geometry"]
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)


def default_config():
    install()
    from config import read_config
    return read_config.read_config(config_name="config")
