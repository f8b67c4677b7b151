"""World-size-2 gloo test (CPU) of the ensemble sharding plumbing: member partition, gather in member order and the
all-reduced RTD statistics.  The integration itself needs a GPU and is covered by the -m gpu tests."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world_size, port, n_members, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world_size)
    try:
        from openccm_b200.ensemble import gather_members, reduce_statistics, shard_bounds
        lo, hi = shard_bounds(n_members, rank, world_size)
        glob = torch.arange(n_members * 6, dtype=torch.float64).reshape(n_members, 2, 3) ** 1.5
        local = glob[lo:hi].clone()
        full = gather_members(local, n_members)
        stats = reduce_statistics(local, n_members)
        # members dealt by a stiffness estimate (boustrophedon) instead of contiguous blocks: gather + inverse permutation
        from openccm_b200.ensemble import balanced_order
        est = np.random.default_rng(5).uniform(0.0, 1.0, n_members)
        perm = balanced_order(est, world_size)
        dealt = glob[torch.as_tensor(perm[lo:hi])].clone()
        back = gather_members(dealt, n_members)[torch.as_tensor(np.argsort(perm))]
        # rooted gather: rank 0 receives every member, the others nothing
        rooted = gather_members(local, n_members, root=0)
        ok_root = torch.equal(rooted, glob) if rank == 0 else rooted is None
        ok = ok_root and torch.equal(back, glob) and torch.equal(full, glob) and torch.allclose(stats['mean'], glob.mean(0)) \
            and torch.equal(stats['min'], glob.min(0).values) and torch.equal(stats['max'], glob.max(0).values) \
            and torch.allclose(stats['var'], glob.var(0, unbiased=False))
        q.put((rank, bool(ok), (lo, hi)))
    finally:
        dist.destroy_process_group()


def test_gather_and_reduce_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    n_members = 7                                   # not divisible by 2: uneven shards
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_members, q)) for r in range(2)]
    for p in procs: p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs: p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    bounds = sorted(b for _, _, b in res)
    assert bounds == [(0, 3), (3, 7)]


def test_shard_bounds_cover_all_members():
    from openccm_b200.ensemble import shard_bounds
    for n in (1, 5, 8, 4096, 4097):
        for ws in (1, 2, 4, 8):
            b = [shard_bounds(n, r, ws) for r in range(ws)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(ws - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_balanced_order_deals_every_rank_the_same_mix():
    from openccm_b200.ensemble import balanced_order, shard_bounds
    rng = np.random.default_rng(0)
    for n, ws in ((7, 2), (4096, 8), (513, 4), (5, 8), (64, 1)):
        est = rng.lognormal(0.0, 1.0, n)
        perm = balanced_order(est, ws)
        assert sorted(perm.tolist()) == list(range(n))                     # a permutation
        sums = []
        for r in range(ws):
            lo, hi = shard_bounds(n, r, ws)
            mine = est[perm[lo:hi]]
            assert np.all(np.diff(mine) <= 0)                              # stiffest first inside a rank
            sums.append(mine.sum())
        if n >= 8 * ws:                                                    # every rank carries the same share of the work
            assert max(sums) / min(sums) < 1.05, sums
