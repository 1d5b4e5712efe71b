"""Decoy sharding across the GPUs of one box, and pool selection.

The reference's parallelism is one OS process per decoy (utils_trX2dy/utils.py:495-503)
followed by a pick of the best decoy (run_inference.py:61-69).  Here: one process per GPU
(torchrun), decoys dealt round-robin to ranks, no data-path collective; the only exchange is
an all-gather of per-decoy scalars (energy terms, score) for pool selection -- N x 8 B per
scalar, latency-bound on NVLink."""
from __future__ import annotations

import numpy as np


def shard(n_total, rank, world):
    """Global decoy indices of this rank: rank, rank+world, ...  (balanced to within one)."""
    return np.arange(rank, n_total, world, dtype=np.int64)


def shard_counts(n_total, world):
    return [len(range(r, n_total, world)) for r in range(world)]


def gather_scalars(local, n_total, rank, world, group=None, device=None):
    """All-gather per-decoy rows (n_local, k) float64 -> (n_total, k) in GLOBAL decoy order on
    every rank.  Uses torch.distributed when world > 1 (NCCL on GPU tensors, gloo on CPU)."""
    local = np.ascontiguousarray(local, dtype=np.float64)
    if local.ndim == 1:
        local = local[:, None]
    k = local.shape[1]
    if world == 1:
        return local.copy()
    import torch
    import torch.distributed as dist
    counts = shard_counts(n_total, world)
    nmax = max(counts)
    buf = torch.zeros((nmax, k), dtype=torch.float64, device=device or "cpu")
    buf[: local.shape[0]] = torch.from_numpy(local).to(buf.device)
    out = torch.empty((world * nmax, k), dtype=torch.float64, device=buf.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    out = out.cpu().numpy().reshape(world, nmax, k)
    full = np.empty((n_total, k))
    for r in range(world):
        full[r::world] = out[r, : counts[r]]
    return full


def select_pool(scores, k, lower_is_better=True):
    """Indices of the k best decoys (ties broken by index, so every rank agrees)."""
    scores = np.asarray(scores, dtype=np.float64)
    key = scores if lower_is_better else -scores
    order = np.lexsort((np.arange(len(key)), key))
    return order[:k]


def assign_blocks(lengths, n_decoys, world, block=32):
    """Batch mode (name_lst; BASELINE config 5): targets of different length, `n_decoys[t]` decoys each.
    The unit of work is a (target, decoy block) pair -- a block is up to `block` decoys, one warp-wide
    decoy group on the device -- with estimated cost ~ L^2 x decoys (restraint evaluations dominate).
    Blocks are dealt longest-first to the least loaded rank (ties: lowest rank), so every rank derives the
    same assignment without communication.  Returns per rank a list of (target, first decoy, count)."""
    items = []
    for t, (L, n) in enumerate(zip(lengths, n_decoys)):
        for d0 in range(0, int(n), block):
            cnt = min(block, int(n) - d0)
            items.append((float(L) * float(L) * cnt, t, d0, cnt))
    items.sort(key=lambda it: (-it[0], it[1], it[2]))
    load = [0.0] * world
    out = [[] for _ in range(world)]
    for cost, t, d0, cnt in items:
        r = min(range(world), key=lambda k: (load[k], k))
        load[r] += cost
        out[r].append((t, d0, cnt))
    merged = []
    for r in range(world):   # contiguous blocks of one target become one fold call
        blocks = sorted(out[r])
        m = []
        for t, d0, cnt in blocks:
            if m and m[-1][0] == t and m[-1][1] + m[-1][2] == d0:
                m[-1] = (t, m[-1][1], m[-1][2] + cnt)
            else:
                m.append((t, d0, cnt))
        merged.append(m)
    return merged
