"""Multi-GPU host logic: contiguous candidate-index shards, one per rank, obstacle set replicated; the only
exchange is an all-gather of the 88-byte per-shard summary records (verdict counts + best candidate), merged on
every rank by rqcd_merge_summaries. The transport is torch.distributed: NCCL over
NVLink on the GPU box, gloo in the CPU tests."""
from __future__ import annotations

import numpy as np

from .api import merge_summaries

RECORD_WORDS = 11  # traj_counts[4], pair_counts[4], pairs_checked, best_cost (bits), best_index


def shard_range(n_global: int, rank: int, world: int) -> tuple[int, int]:
    """[lo, hi) of rank's contiguous candidate range: [r*N/G, (r+1)*N/G)."""
    return n_global * rank // world, n_global * (rank + 1) // world


def pack_summary(s: dict) -> np.ndarray:
    rec = np.zeros(RECORD_WORDS, dtype=np.int64)
    rec[0:4] = np.asarray(s["traj_counts"], dtype=np.uint64).astype(np.int64)
    rec[4:8] = np.asarray(s["pair_counts"], dtype=np.uint64).astype(np.int64)
    rec[8] = s["pairs_checked"]
    rec[9] = np.array([s["best_cost"]], dtype=np.float64).view(np.int64)[0]
    rec[10] = s["best_index"]
    return rec


def unpack_summary(rec: np.ndarray) -> dict:
    rec = np.asarray(rec, dtype=np.int64)
    return {"traj_counts": [int(x) for x in rec[0:4]], "pair_counts": [int(x) for x in rec[4:8]],
            "pairs_checked": int(rec[8]), "best_cost": float(rec[9:10].view(np.float64)[0]), "best_index": int(rec[10])}


def exchange_summaries(local: dict, device=None, gather_buf=None) -> dict:
    """All-gather the ranks' summary records and merge them (deterministic: lowest cost, lowest index on ties).
    With one rank this is the identity."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    mine = torch.from_numpy(pack_summary(local))
    if device is not None:
        mine = mine.to(device)
    if gather_buf is None:
        gather_buf = torch.zeros((world, RECORD_WORDS), dtype=torch.int64, device=mine.device)
    dist.all_gather(list(gather_buf.unbind(0)), mine)  # works on both nccl and gloo
    recs = gather_buf.cpu().numpy()
    return merge_summaries([unpack_summary(r) for r in recs])
