"""Sweep of the warps / CTAs-per-SM configuration of the mechanism-specialised stage kernels (csrc/occm_spec.h) on the C4
workload: live CUDA-event times of the plain RHS and of Dormand-Prince stages 2 and 7 per configuration.  Development aid.

    python scripts/sweep_spec.py [members] "<flags of config 1>" "<flags of config 2>" ...
"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import torch
from openccm_b200 import synthetic as syn
from openccm_b200.device import DeviceSystem, MemberParams
from openccm_b200.system_solvers.export import export_system
from openccm_b200.workloads import cstr_ensemble_workload

M = int(sys.argv[1]) if len(sys.argv) > 1 else 512
CONFIGS = sys.argv[2:] or ['']
N = 100_000
PEAK = 6551.7
model, net, cp, cmesh = cstr_ensemble_workload(n_cstr=N)
host = export_system(model, net, cp, cmesh)
ks, qs, ics = syn.ensemble_sweeps(M, host.tables.n_rxn)
mem = MemberParams(M, torch.as_tensor(ks, device='cuda'), torch.as_tensor(qs, device='cuda'), None)
y0 = torch.as_tensor(host.to_device_layout(host.c0.reshape(-1)), device='cuda')
y = (torch.as_tensor(ics, device='cuda')[:, None] * y0[None, :]).contiguous()
te = np.linspace(0, 10, 101)
save = torch.arange(160, dtype=torch.int32, device='cuda')
csr = 12 * len(host.row_src) + 4 * (N + 1)
BYTES = {0: 16, 2: 32, 7: 32}
ref = None
BURN = int(os.environ.get('SWEEP_BURN', '0'))      # full ensemble solves before the first timing: clocks settle at the power cap
if BURN:
    dev = DeviceSystem(host)
    for _ in range(BURN):
        dev.rk45(y, (0.0, 10.0), rtol=1e-6, atol=1e-10, first_step=1e-4, t_eval=te, save_idx=save, members=mem)
    del dev
for flags in CONFIGS:
    env = {}
    for tok in flags.split():                     # KEY=VALUE tokens set environment switches, the rest are nvcc flags
        if '=' in tok and not tok.startswith('-'):
            k, v = tok.split('=', 1); env[k] = v
    os.environ['OCCM_B200_SPEC_FLAGS'] = ' '.join(t for t in flags.split() if t.startswith('-'))
    for k, v in env.items():
        os.environ[k] = v
    dev = DeviceSystem(host)
    dy = torch.empty_like(y)
    for _ in range(3):
        dev.rhs(y, 0.0, mem, out=dy)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        dev.rhs(y, 0.0, mem, out=dy)
    e1.record(); torch.cuda.synchronize()
    t = {0: e0.elapsed_time(e1) / 10}
    if BURN:
        dev.rk45(y, (0.0, 10.0), rtol=1e-6, atol=1e-10, first_step=1e-4, t_eval=te, save_idx=save, members=mem)
    with dev.open_rk45(y, (0.0, 10.0), rtol=1e-6, atol=1e-10, first_step=1e-4, t_eval=te, save_idx=save, members=mem) as run:
        run.run(max_launch_steps=2)
        for s in (2, 7):
            run.time_stage(s, 2 if not BURN else 20)
            t[s] = run.time_stage(s, 6 if not BURN else 20)
    solve_ms = None
    if BURN:                                      # mean of three complete ensemble solves (CUDA events), clocks at the power cap
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(3):
            dev.rk45(y, (0.0, 10.0), rtol=1e-6, atol=1e-10, first_step=1e-4, t_eval=te, save_idx=save, members=mem)
        s1.record(); torch.cuda.synchronize()
        solve_ms = s0.elapsed_time(s1) / 3
    chk = float(dy.sum())
    ref = chk if ref is None else ref
    frac = {s: (M * host.ndof * BYTES[s] + csr) / v * 1e-6 / PEAK for s, v in t.items()}
    print(f'spec={dev.specialized} [{flags}] ' + '  '.join(f'stage {s}: {t[s]:.3f} ms ({frac[s]:.3f})' for s in (0, 2, 7)) +
          (f'  solve {solve_ms:.1f} ms' if solve_ms else '') + f'  checksum {"same" if chk == ref else "DIFFERENT"}', flush=True)
    for k in env:
        os.environ.pop(k, None)
    del dev
