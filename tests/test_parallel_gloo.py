"""N>1 host logic on CPU: world_size-2 gloo processes shard decoys, all-gather per-decoy
scalars and agree on the selected pool."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    sys.path.insert(0, ROOT)
    import trx2dyn  # noqa: F401
    from trx2dyn import parallel
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = parallel.shard(n_total, rank, world)
    # a "fold" stand-in: per-decoy terms that depend only on the GLOBAL decoy index
    terms = np.stack([np.sin(idx * 0.37) * 100, idx.astype(float)], axis=1)
    full = parallel.gather_scalars(terms, n_total, rank, world)
    pool = parallel.select_pool(full[:, 0], 5)
    q.put((rank, idx.tolist(), full.tolist(), pool.tolist()))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [10, 37])
def test_shard_gather_select_world2(n_total):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    all_idx = sorted(res[0][1] + res[1][1])
    assert all_idx == list(range(n_total))                       # a partition of the decoys
    assert abs(len(res[0][1]) - len(res[1][1])) <= 1
    want = np.stack([np.sin(np.arange(n_total) * 0.37) * 100, np.arange(n_total, dtype=float)], axis=1)
    for r in res:
        np.testing.assert_allclose(np.array(r[2]), want)          # global order on every rank
    assert res[0][3] == res[1][3] == np.argsort(want[:, 0], kind="stable")[:5].tolist()


def test_single_rank_is_identity():
    sys.path.insert(0, ROOT)
    import trx2dyn  # noqa: F401
    from trx2dyn import parallel
    x = np.arange(12.0).reshape(6, 2)
    np.testing.assert_array_equal(parallel.gather_scalars(x, 6, 0, 1), x)
    assert parallel.shard_counts(10, 4) == [3, 3, 2, 2]
    assert parallel.select_pool([3.0, 1.0, 1.0, 2.0], 3).tolist() == [1, 2, 3]


def test_batch_mode_assignment_is_a_balanced_partition():
    """BASELINE config 5: 64 targets L in [100, 500], 100 decoys each, on 8 ranks."""
    sys.path.insert(0, ROOT)
    import trx2dyn  # noqa: F401
    from trx2dyn import parallel
    rng = np.random.default_rng(5)
    lengths = rng.integers(100, 501, size=64)
    n_dec = [100] * 64
    plan = parallel.assign_blocks(lengths, n_dec, 8)
    seen = np.zeros((64, 100), dtype=int)
    load = []
    for r in range(8):
        cost = 0.0
        for t, d0, cnt in plan[r]:
            seen[t, d0:d0 + cnt] += 1
            cost += float(lengths[t]) ** 2 * cnt
        load.append(cost)
    assert np.all(seen == 1)                                          # every decoy of every target exactly once
    assert max(load) / (sum(load) / 8) < 1.02                         # within 2 % of perfect balance
    assert plan == parallel.assign_blocks(lengths, n_dec, 8)          # deterministic: ranks agree without talking
    one = parallel.assign_blocks([50], [7], 4)
    assert one[0] == [(0, 0, 7)] and one[1:] == [[], [], []]
