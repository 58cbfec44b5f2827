"""CPU: the oracle's split_path (oracle/avp_oracle.c orc_split_path) against vectors of the UNMODIFIED reference
(tests/golden/leaf_split.npz, made by gen_split_golden.py; tests/golden/cases/*.npz hold split_lens / out_final_path of the
finished BenchmarkCases)."""
import ctypes
import os

import numpy as np

import oracle_lib as O
from automatedvaletparking_b200 import scenarios as scn
from conftest import GOLDEN


def test_np_dot_of_two_vectors_is_one_fma():
    """scipy.spatial.distance.cosine uses np.dot; on 2-vectors the installed BLAS evaluates fma(a1, b1, a0 * b0)"""
    g = np.load(os.path.join(GOLDEN, "leaf_split.npz"))
    libm = ctypes.CDLL("libm.so.6")
    libm.fma.restype = ctypes.c_double
    libm.fma.argtypes = [ctypes.c_double] * 3
    U, V = g["dot_u"], g["dot_v"]
    got = np.array([libm.fma(U[i, 1], V[i, 1], U[i, 0] * V[i, 0]) for i in range(len(U))])
    assert np.array_equal(got, g["dot_uv"])


def test_split_path_matches_reference_vectors(cfg):
    g = np.load(os.path.join(GOLDEN, "leaf_split.npz"))
    maps = {}
    n_err = n_ext = 0
    for k in range(int(g["n"])):
        case = int(g[f"p{k}_case"])
        if case not in maps:
            maps[case] = O.OracleMap(scn.benchmark_case(case))
        r = maps[case].split_path(cfg, g[f"p{k}_in"])
        assert r["status"] == int(g[f"p{k}_status"]), k
        if r["status"] == 0:
            assert list(r["seg_len"]) == list(g[f"p{k}_lens"]) and r["change_gear"] == int(g[f"p{k}_cg"]), k
            assert np.array_equal(r["out_final_path"], g[f"p{k}_out"]), k
            n_ext += len(r["out_final_path"]) > len(g[f"p{k}_in"]) + r["change_gear"]
        else:
            n_err += 1
    assert n_err >= 10 and n_ext >= 50


def test_split_path_of_the_finished_benchmark_cases(cfg):
    n = 0
    for c in range(1, 21):
        g = np.load(os.path.join(GOLDEN, "cases", f"Case{c}.npz"))
        if str(g["status"]) != "ok":
            continue
        final = np.concatenate([g["astar_path"], np.stack([g["rs_x"], g["rs_y"], g["rs_yaw"]], 1)[1:]], 0)
        r = O.OracleMap(scn.benchmark_case(c)).split_path(cfg, final)
        assert r["status"] == 0 and list(r["seg_len"]) == list(g["split_lens"]) and r["change_gear"] == int(g["change_gear"])
        assert np.array_equal(r["out_final_path"], g["out_final_path"])
        n += 1
    assert n >= 16
