"""Scenario sharding across ranks and the all-gather of finished trajectories.

Scenarios are independent (each owns its raster, h table, open/closed sets), so the data path
has NO collective; one all-gather of fixed-stride result records at the end makes every rank
hold every trajectory.  One process per GPU; torch.distributed is plumbing only (NCCL on the
GPU box over NVLink/NVSwitch, gloo in CPU tests).
"""
from typing import List, Sequence, Tuple

import numpy as np

from .hostcfg import SUMMARY_DTYPE


def cost_proxy(scenario) -> float:
    """Cheap ordering key: start-goal distance times map area (Dijkstra work grows with both)."""
    span_x = abs(scenario.x0 - scenario.xf) + 24.0
    span_y = abs(scenario.y0 - scenario.yf) + 24.0
    return float(np.hypot(scenario.x0 - scenario.xf, scenario.y0 - scenario.yf) * span_x * span_y)


def shard_indices(n: int, rank: int, world: int, keys: Sequence[float] = None) -> np.ndarray:
    """Deal scenario ids round-robin after sorting by the cost proxy (expensive first), so every
    rank gets the same mix of cheap and expensive searches."""
    order = np.arange(n) if keys is None else np.argsort(-np.asarray(keys, dtype=np.float64), kind="stable")
    return order[rank::world]


def unshard(parts: List[np.ndarray], n: int, world: int, keys: Sequence[float] = None) -> np.ndarray:
    """Inverse of shard_indices for gathered per-rank arrays (axis 0): returns rows in scenario order."""
    order = np.arange(n) if keys is None else np.argsort(-np.asarray(keys, dtype=np.float64), kind="stable")
    out = np.empty((n,) + parts[0].shape[1:], dtype=parts[0].dtype)
    for r in range(world):
        idx = order[r::world]
        out[idx] = parts[r][:len(idx)]
    return out


class _DevArray:
    """Zero-copy view of a device buffer owned by libavp_b200 (for torch.as_tensor)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def gather_results(summaries: np.ndarray, paths: np.ndarray, per_rank: int, group=None) -> Tuple[np.ndarray, np.ndarray]:
    """Host-buffer all-gather (gloo or nccl via staging): every rank passes its (padded to per_rank
    rows) summaries/paths and receives the rank-major concatenation."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    s = np.zeros(per_rank, dtype=SUMMARY_DTYPE)
    s[:len(summaries)] = summaries
    p = np.zeros((per_rank,) + paths.shape[1:], dtype=np.float64)
    p[:len(paths)] = paths
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    ts = torch.from_numpy(s.view(np.uint8).reshape(-1)).to(dev)
    tp = torch.from_numpy(p.reshape(-1)).to(dev)
    gs = torch.empty(world * ts.numel(), dtype=torch.uint8, device=dev)
    gp = torch.empty(world * tp.numel(), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(gs, ts, group=group)
    dist.all_gather_into_tensor(gp, tp, group=group)
    S = gs.cpu().numpy().view(SUMMARY_DTYPE).reshape(world, per_rank)
    P = gp.cpu().numpy().reshape((world, per_rank) + paths.shape[1:])
    return S, P


def gather_results_device(dp, per_rank: int, cap_path: int, group=None):
    """NCCL all-gather straight from the library's device result buffers (no host staging):
    returns device tensors (world*per_rank*sizeof(summary) bytes, world*per_rank*cap_path*3 f64)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sp, pp, n, cp = dp.result_device_pointers()
    assert cp == cap_path and n <= per_rank
    dev = torch.device("cuda", torch.cuda.current_device())
    ts = torch.zeros(per_rank * SUMMARY_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    tp = torch.zeros(per_rank * cap_path * 3, dtype=torch.float64, device=dev)
    ts[: n * SUMMARY_DTYPE.itemsize] = torch.as_tensor(_DevArray(sp, n * SUMMARY_DTYPE.itemsize), device=dev)
    tp[: n * cap_path * 3] = torch.as_tensor(_DevArray(pp, n * cap_path * 3 * 8), device=dev).view(torch.float64)
    gs = torch.empty(world * ts.numel(), dtype=torch.uint8, device=dev)
    gp = torch.empty(world * tp.numel(), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(gs, ts, group=group)
    dist.all_gather_into_tensor(gp, tp, group=group)
    return gs, gp


def records_checksum(summaries: np.ndarray, paths: np.ndarray) -> str:
    """SHA-256 (first 16 hex digits) over the result records of a job in scenario order: the summaries and the rows of every
    returned path.  Scenario sharding must not change it: bench.py prints it for every N (strong-scaling workloads)."""
    import hashlib
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(summaries).tobytes())
    cap = paths.shape[1]
    for k in range(len(summaries)):
        h.update(np.ascontiguousarray(paths[k, :min(int(summaries["n_final"][k]), cap)]).tobytes())
    return h.hexdigest()[:16]
