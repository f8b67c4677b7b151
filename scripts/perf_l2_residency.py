"""L2-residency experiment: how fast are the stage kernels when a member group's vectors stay in
the 126 MB L2?  Per-member stage times at M = 1 .. 512 members per launch (C4: 8 MB per vector and member, 11 vectors), timed
back to back on the same buffers, i.e. with everything a launch reads already in L2 when the group is small; and complete
ensemble solves with small batches against the one big batch.

    python scripts/perf_l2_residency.py
"""
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import torch
from openccm_b200 import synthetic as syn
from openccm_b200.device import DeviceSystem, MemberParams
from openccm_b200.system_solvers.export import export_system
from openccm_b200.workloads import cstr_ensemble_workload

N = 100_000
model, net, cp, cmesh = cstr_ensemble_workload(n_cstr=N)
host = export_system(model, net, cp, cmesh)
dev = DeviceSystem(host)
te = np.linspace(0, 10, 101)
save = torch.arange(160, dtype=torch.int32, device='cuda')
y0 = torch.as_tensor(host.to_device_layout(host.c0.reshape(-1)), device='cuda')
BYTES = {2: 32, 3: 48, 4: 56, 5: 64, 6: 72, 7: 32}
print('members | per-member stage time [us] (2..7) | sum | effective GB/s of the 304 B/element | vectors of the group [MB]')
for M in (1, 2, 4, 8, 16, 64, 512):
    ks, qs, ics = syn.ensemble_sweeps(M, host.tables.n_rxn)
    mem = MemberParams(M, torch.as_tensor(ks, device='cuda'), torch.as_tensor(qs, device='cuda'), None)
    y = (torch.as_tensor(ics, device='cuda')[:, None] * y0[None, :]).contiguous()
    with dev.open_rk45(y, (0.0, 10.0), rtol=1e-6, atol=1e-10, first_step=1e-4, t_eval=te, save_idx=save, members=mem) as run:
        run.run(max_launch_steps=2)
        st = {}
        for s in (2, 3, 4, 5, 6, 7):
            run.time_stage(s, 5)
            st[s] = run.time_stage(s, 20) * 1e3 / M
    tot = sum(st.values())
    print(f'{M:7d} | ' + ' '.join(f'{v:6.1f}' for v in st.values()) + f' | {tot:6.1f} | {host.ndof * 304 / tot * 1e-3:7.0f} | {11 * M * host.ndof * 8 / 1e6:7.0f}', flush=True)

# complete solves: 512 members in batches of B (one handle per batch, sequentially on one stream)
M = 512
ks, qs, ics = syn.ensemble_sweeps(M, host.tables.n_rxn)
for B in (1, 4, 16, 512):
    torch.cuda.synchronize(); t0 = time.time()
    nfev = 0
    for lo in range(0, M if B >= 16 else 64, B):            # small batches: 64 members are enough for a rate
        mem = MemberParams(B, torch.as_tensor(ks[lo:lo + B], device='cuda'), torch.as_tensor(qs[lo:lo + B], device='cuda'), None)
        y = (torch.as_tensor(ics[lo:lo + B], device='cuda')[:, None] * y0[None, :]).contiguous()
        r = dev.rk45(y, (0.0, 10.0), rtol=1e-6, atol=1e-10, first_step=1e-4, t_eval=te, save_idx=save, members=mem)
        nfev += int(r.nfev.sum())
    torch.cuda.synchronize(); dt = time.time() - t0
    print(f'batches of {B:3d}: {nfev * host.ndof / dt / 1e9:6.1f} G evals/s', flush=True)
