"""Shim with the reference's module path `collision_check/collision_check.py`: re-exports automatedvaletparking_b200.collision_check.collision_check
(this directory, and only this directory, goes in front of the reference root on sys.path)."""
from automatedvaletparking_b200.collision_check.collision_check import *  # noqa: F401,F403
import automatedvaletparking_b200.collision_check.collision_check as _impl
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
