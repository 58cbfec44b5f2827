"""Shim with the reference's module path `path_plan/hybrid_a_star.py`: re-exports automatedvaletparking_b200.path_plan.hybrid_a_star
(this directory, and only this directory, goes in front of the reference root on sys.path)."""
from automatedvaletparking_b200.path_plan.hybrid_a_star import *  # noqa: F401,F403
import automatedvaletparking_b200.path_plan.hybrid_a_star as _impl
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
