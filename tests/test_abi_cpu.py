"""CPU: the C-ABI library loads, exports every symbol include/rqcd.h declares, refuses to run without a
GPU (no CPU fallback), and its pure-host entry points behave."""
import ctypes as C
import math
import os
import re

import pytest

from rapidquadcoptercollisiondetection_b200 import abi
from rapidquadcoptercollisiondetection_b200.api import merge_summaries

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "rqcd.h")).read()
    return sorted(set(re.findall(r"\b(rqcd_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = abi.load()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"librqcd.so does not export {n}"
    assert set(names) == set(abi.SYMBOLS), "abi.py mirror and include/rqcd.h disagree"


def test_struct_mirrors_match_header_sizes():
    # rqcd_obstacle: 2 x int32 + (3+1+3+4+18+2) doubles; rqcd_summary: 9 x 8 bytes + double + int64
    assert C.sizeof(abi.Obstacle) == 8 + 31 * 8
    assert C.sizeof(abi.Summary) == 11 * 8
    assert C.sizeof(abi.BcSoa) == 40 and C.sizeof(abi.CoeffSoa) == 32 and C.sizeof(abi.Out) == 88


def test_no_device_means_no_service():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = abi.load()
    h = C.c_void_p()
    st = lib.rqcd_create(C.byref(h), 0)
    assert st == abi.ERR_NO_DEVICE and not h.value
    assert b"no CPU fallback" in lib.rqcd_strerror(st)


def test_merge_summaries_is_the_only_exchange_step():
    inf = math.inf
    shards = [
        {"traj_counts": [5, 1, 0, 0], "pair_counts": [50, 1, 0, 0], "pairs_checked": 51, "best_cost": 3.5, "best_index": 7},
        {"traj_counts": [0, 6, 0, 0], "pair_counts": [10, 6, 0, 0], "pairs_checked": 16, "best_cost": inf, "best_index": -1},
        {"traj_counts": [4, 1, 1, 0], "pair_counts": [40, 1, 1, 0], "pairs_checked": 42, "best_cost": 3.5, "best_index": 3},
        {"traj_counts": [2, 0, 0, 0], "pair_counts": [20, 0, 0, 0], "pairs_checked": 20, "best_cost": 9.0, "best_index": 1},
    ]
    m = merge_summaries(shards)
    assert m["traj_counts"] == [11, 8, 1, 0] and m["pair_counts"] == [120, 8, 1, 0] and m["pairs_checked"] == 129
    assert m["best_cost"] == 3.5 and m["best_index"] == 3  # tie on cost -> lowest index
    e = merge_summaries([])
    assert e["best_index"] == -1 and e["best_cost"] == inf and e["pairs_checked"] == 0


def test_shard_range_matches_the_python_host_logic():
    from rapidquadcoptercollisiondetection_b200 import sharding
    lib = abi.load()
    for n in (0, 1, 7, 1000, (1 << 26) + 5, 1 << 40):
        for world in (1, 2, 3, 8):
            prev = 0
            for k in range(world):
                lo, hi = C.c_int64(), C.c_int64()
                lib.rqcd_shard_range(n, k, world, C.byref(lo), C.byref(hi))
                assert (lo.value, hi.value) == sharding.shard_range(n, k, world)
                assert lo.value == prev
                prev = hi.value
            assert prev == n
