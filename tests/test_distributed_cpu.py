"""world_size-2 gloo tests (CPU) of the N>1 host logic: scenario sharding and the all-gather of result
records.  The planner is replaced by a deterministic stub (no GPU here); on the GPU box bench.py runs
the same gather over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from automatedvaletparking_b200 import distributed as avd
from automatedvaletparking_b200 import scenarios as scn
from automatedvaletparking_b200.hostcfg import SUMMARY_DTYPE


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _stub_results(ids, cap_path=8):
    """what a rank would produce for its scenario ids: recognisable, deterministic records"""
    s = np.zeros(len(ids), dtype=SUMMARY_DTYPE)
    s["status"] = ids % 3
    s["n_pops"] = 10 + ids
    s["global_index"] = 100 * ids
    p = np.zeros((len(ids), cap_path, 3))
    p[:, :, 0] = ids[:, None] + 0.25
    return s, p


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    scs = [scn.Scenario(float(i), 0.0, 0.1, float(i) + 3 + (i % 5), 4.0, 0.2, [], None, f"s{i}") for i in range(n)]
    keys = [avd.cost_proxy(s) for s in scs]
    mine = avd.shard_indices(n, rank, world, keys)
    s, p = _stub_results(mine)
    per_rank = (n + world - 1) // world
    S, P = avd.gather_results(s, p, per_rank)
    full_s = avd.unshard([S[r] for r in range(world)], n, world, keys)
    full_p = avd.unshard([P[r] for r in range(world)], n, world, keys)
    q.put((rank, mine.tolist(), full_s["global_index"].tolist(), full_p[:, 0, 0].tolist(), avd.records_checksum(full_s, full_p)))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [10, 7])
def test_shard_and_gather_two_ranks(n):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    out.sort()
    ids0, ids1 = out[0][1], out[1][1]
    assert sorted(ids0 + ids1) == list(range(n)) and not set(ids0) & set(ids1)          # a partition
    s1, p1 = _stub_results(np.arange(n))                                              # what ONE rank produces for the whole job
    for _, _, gi, px, cks in out:                                                     # every rank holds every record, in scenario order
        assert gi == [100 * i for i in range(n)]
        assert px == [i + 0.25 for i in range(n)]
        assert cks == avd.records_checksum(s1, p1)                                    # sharding does not change the job's records


def test_shard_count_invariance():
    """1/2/4/8 ranks -> the same set of scenarios, balanced mix of cheap and expensive ones"""
    n = 64
    keys = np.random.default_rng(0).uniform(1, 100, n)
    for world in (1, 2, 4, 8):
        parts = [avd.shard_indices(n, r, world, keys) for r in range(world)]
        assert sorted(np.concatenate(parts).tolist()) == list(range(n))
        loads = [keys[p].sum() for p in parts]
        assert max(loads) / min(loads) < 1.35
        rec = [np.stack([p.astype(np.float64), keys[p]], 1) for p in parts]
        per = (n + world - 1) // world
        rec = [np.concatenate([r, np.zeros((per - len(r), 2))]) for r in rec]
        full = avd.unshard(rec, n, world, keys)
        assert np.array_equal(full[:, 0], np.arange(n)) and np.array_equal(full[:, 1], keys)
