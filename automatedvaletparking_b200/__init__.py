"""B200-native hybrid-A* hot path behind the reference's planner call surface.

Layout:
  csrc/                 CUDA kernels + the C ABI (include/avp_b200.h) -> libavp_b200.so
  _native.py            ctypes binding of the C ABI (fails loudly if the .so is missing)
  hostcfg.py            config dict + Vehicle constants -> avp_config
  scenarios.py          Case CSV parsing / scenario batches / synthetic scenario recipes
  batch.py              plan_batch(): whole searches for many scenarios on one GPU
  distributed.py        scenario sharding across ranks + NCCL all-gather of results
  map/, collision_check/, path_plan/, config/   drop-in modules with the reference's
                        module paths, class names and method signatures
"""
__all__ = ["__version__"]
__version__ = "0.1.0"
