"""CPU, world_size 2 over gloo: the N>1 path's host logic -- contiguous candidate shards, replicated obstacles,
all-gather of the per-shard summary records, merge -- gives the same summary as one shard over everything.
Each rank's shard is computed by the CPU oracle here (no GPU in this test); on the GPU box bench.py runs the same
exchange over NCCL with the CUDA path computing the shards."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_global, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from oracle import binding as B
    from rapidquadcoptercollisiondetection_b200 import abi, sharding, workloads as W
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    specs = W.static_obstacles(2, 4, 4)
    obs = abi.make_obstacles(specs)
    lo, hi = sharding.shard_range(n_global, rank, world)
    bc = W.synth_bc(1, lo, hi - lo)
    local = B.Oracle().generate_check(bc, obs, len(specs), 0.002, index_base=lo, matrix=False)["summary"]
    merged = sharding.exchange_summaries(local)
    q.put((rank, lo, hi, merged))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_exchange_equals_single_shard(oracle):
    from rapidquadcoptercollisiondetection_b200 import abi, sharding, workloads as W
    n_global, world = 3001, 2  # odd on purpose: ragged shards
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_global, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    specs = W.static_obstacles(2, 4, 4)
    whole = oracle.generate_check(W.synth_bc(1, 0, n_global), abi.make_obstacles(specs), len(specs), 0.002, matrix=False)["summary"]
    ranges = sorted((lo, hi) for _, lo, hi, _ in got)
    assert ranges[0][0] == 0 and ranges[-1][1] == n_global and ranges[0][1] == ranges[1][0]
    for _, _, _, merged in got:
        assert merged == whole
    assert whole["best_index"] >= 0


def test_pack_unpack_roundtrip():
    from rapidquadcoptercollisiondetection_b200 import sharding
    s = {"traj_counts": [1, 2, 3, 0], "pair_counts": [10, 20, 30, 0], "pairs_checked": 60, "best_cost": 1.25e-3, "best_index": 2 ** 40 + 5}
    assert sharding.unpack_summary(sharding.pack_summary(s)) == s
    e = {"traj_counts": [0] * 4, "pair_counts": [0] * 4, "pairs_checked": 0, "best_cost": float("inf"), "best_index": -1}
    assert sharding.unpack_summary(sharding.pack_summary(e)) == e
    assert [sharding.shard_range(10, r, 4) for r in range(4)] == [(0, 2), (2, 5), (5, 7), (7, 10)]
