"""Shim with the reference's module path `config/read_config.py`: re-exports automatedvaletparking_b200.config.read_config
(this directory, and only this directory, goes in front of the reference root on sys.path)."""
from automatedvaletparking_b200.config.read_config import *  # noqa: F401,F403
import automatedvaletparking_b200.config.read_config as _impl
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
