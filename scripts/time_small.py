"""Reference-sized networks (BASELINE configs C1-C3): persistent one-CTA-per-member RK45 (occm_small.cuh) against the batched
kernels (OCCM_B200_NO_SMALL=1), single solves through solve_system and 4096-member ensembles through EnsembleSolver.

    python scripts/time_small.py
"""
import contextlib
import io
import os
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, 'tests'), os.path.join(REPO, 'tests', 'golden')]
import numpy as np
import torch
import cases as C
from _util import setup_case
from openccm_b200.ensemble import EnsembleSolver
from openccm_b200.system_solvers import solve_system

SINGLE = (('cstr_irreversible', 'B200-RK45', {}), ('openfoam_2d_cstr_nacl', 'B200-RK45', dict(rtol=1e-6, atol=1e-10)),
          ('openfoam_2d_pfr', 'B200-RK45', dict(rtol=1e-6, atol=1e-10)), ('openfoam_2d_pfr', 'B200-BDF', dict(rtol=1e-6, atol=1e-10)))


def single(name, solver, kw, small):
    os.environ.pop('OCCM_B200_NO_SMALL', None)
    if not small:
        os.environ['OCCM_B200_NO_SMALL'] = '1'
    tmp = tempfile.mkdtemp()
    cp, cmesh, net, mn = setup_case(name, solver, tmp, **kw)
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        with contextlib.redirect_stdout(io.StringIO()):
            c, t, im, om, info = solve_system(C.CASES[name]['model'], mn, cp, cmesh, return_info=True)
        torch.cuda.synchronize(); dt = time.time() - t0
    steps = info['n_accepted'] + info['n_rejected']
    return dt, steps, info['nfev'], c


for name, solver, kw in SINGLE:
    variants = (True, False) if solver == 'B200-RK45' else (True,)
    res = {}
    for small in variants:
        dt, steps, nfev, c = single(name, solver, kw, small)
        res[small] = c
        print(f'{name:22s} {solver:10s} {"persistent" if small else "batched   "} {dt:8.3f} s  steps {steps:7d}  nfev {nfev:8d}  {dt / steps * 1e6:8.1f} us/step', flush=True)
    if len(res) == 2:
        print(f'   max |persistent - batched| = {np.abs(res[True] - res[False]).max():.3e}', flush=True)

# ---- ensembles: 4096 members of the two real networks (inflow / rate-constant sweeps), RK45, outlets recorded
os.environ.pop('OCCM_B200_NO_SMALL', None)
for name, model, t_end in (('openfoam_2d_cstr_nacl', 'cstr', 300.0), ('openfoam_2d_pfr', 'pfr', 30.0)):
    tmp = tempfile.mkdtemp()
    cp, cmesh, net, mn = setup_case(name, 'B200-RK45', tmp, rtol=1e-6, atol=1e-10, t_span=f'0, {t_end}', t_eval=f'{t_end / 60}, linear')
    solver = EnsembleSolver(model, mn, cp, cmesh)
    M = 4096
    rng = np.random.default_rng(0)
    qs = rng.uniform(0.5, 2.0, M)
    ks = 10.0 ** rng.uniform(-0.5, 0.5, (M, solver.host.tables.n_rxn)) if solver.host.tables.n_rxn else None
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        r = solver.solve(M, q_scale=qs, k_scale=ks, save='outlets', rtd=False)
        torch.cuda.synchronize(); dt = time.time() - t0
    evals = float(r.nfev.sum()) * solver.host.ndof
    print(f'{name:22s} ensemble M={M} t_end={t_end:g}: {dt:7.3f} s  {M / dt:9.1f} solves/s  {evals / dt / 1e9:7.2f} G evals/s  '
          f'steps/member {int((r.n_accepted + r.n_rejected).mean())}  status {np.unique(r.status)}', flush=True)

# ---- the stiff real network as an ensemble through the batched BDF (all members advance in lock step, per-member order / step)
tmp = tempfile.mkdtemp()
cp, cmesh, net, mn = setup_case('openfoam_2d_pfr', 'B200-BDF', tmp, rtol=1e-6, atol=1e-10)
solver = EnsembleSolver('pfr', mn, cp, cmesh)
for M in (64, 1024):
    qs = np.random.default_rng(0).uniform(0.5, 2.0, M)
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        r = solver.solve(M, q_scale=qs, save='outlets', rtd=True)
        torch.cuda.synchronize(); dt = time.time() - t0
    print(f'openfoam_2d_pfr        BDF ensemble M={M} t in [0, 300], 61 outputs + RTD: {dt:7.3f} s  {M / dt:9.1f} solves/s  '
          f'steps/member {int((r.n_accepted + r.n_rejected).mean())}  status {np.unique(r.status)}', flush=True)
