"""Two ranks on two GPUs over NCCL: the ensemble API shards the members, integrates them, gathers the recorded outputs in
member order and all-reduces the RTD moment statistics — and gives the same numbers as one rank doing all members.
Skipped when fewer than two GPUs are visible."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _solve(tmp, M, world_size):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here); sys.path.insert(0, os.path.join(here, 'golden')); sys.path.insert(0, os.path.dirname(here))
    import pathlib
    from _util import setup_case
    from openccm_b200.ensemble import EnsembleSolver
    cp, cmesh, net, mn = setup_case('cstr_net9_tracer', 'B200-RK45', pathlib.Path(tmp))
    solver = EnsembleSolver('cstr', mn, cp, cmesh)
    rng = np.random.default_rng(0)
    qs = rng.uniform(0.5, 2.0, M)
    return solver.solve(M, q_scale=qs, save='outlets', rtd=True)


def _worker(rank, world_size, port, tmp, M, q):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    os.environ['RANK'] = str(rank); os.environ['WORLD_SIZE'] = str(world_size); os.environ['LOCAL_RANK'] = str(rank)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world_size, device_id=torch.device('cuda', rank))
    try:
        res = _solve(os.path.join(tmp, f'r{rank}'), M, world_size)
        q.put((rank, np.asarray(res.y), np.asarray(res.status), {k: np.asarray(v) for k, v in (res.moment_stats or {}).items()},
               None if res.moments is None else np.asarray(res.moments)))
    finally:
        dist.destroy_process_group()


def test_two_ranks_match_one_rank(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    M = 7
    os.makedirs(tmp_path / 'single'); os.makedirs(tmp_path / 'r0'); os.makedirs(tmp_path / 'r1')
    single = _solve(str(tmp_path / 'single'), M, 1)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, str(tmp_path), M, q)) for r in range(2)]
    for p in procs: p.start()
    out = sorted((q.get(timeout=300) for _ in procs), key=lambda r: r[0])
    for p in procs: p.join(timeout=60)
    for rank, y, status, stats, mom in out:
        assert np.all(status == 1)
        np.testing.assert_array_equal(y, np.asarray(single.y))             # gathered in member order, bit-identical members
        if mom is not None:
            np.testing.assert_array_equal(mom, np.asarray(single.moments))
        for k, v in stats.items():
            np.testing.assert_allclose(v, np.asarray(single.moment_stats[k]), rtol=1e-12, atol=1e-14)
