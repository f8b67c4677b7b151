#!/usr/bin/env python
"""
Tight reference solution of a MID-SIZE C5 instance (stiff 20-reaction mechanism on PFR recycle chains, 1.2e5 DOF) for
tests/test_gpu_fullsize.py::test_c5_mid_bdf_vs_tight_scipy.

The unmodified reference cannot integrate this size (solve_ivp without jac_sparsity builds a dense 1.2e5 x 1.2e5 Jacobian,
pfr_system.py:191), so the yard-stick is the oracle's restated pfr ddt (pinned against the reference's raw ddt on the small golden
cases, tests/test_oracle_vs_golden.py) integrated by scipy's BDF with the exported Jacobian sparsity at
rtol 1e-11 / atol 1e-14.  A run at 1e-10 is the convergence check of that yard-stick, and a run at 1e-8 / 1e-12 — the tolerance
the B200-BDF policy maps the user's 1e-6 / 1e-10 to — records how far scipy's own BDF is from the band there (several bands on
this stiff case: global error / local tolerance grows with stiffness).  Takes ~10 minutes on the build box.

    python tests/golden/make_c5_mid.py
"""
import os
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import numpy as np
import scipy.integrate

from oracle import openccm_oracle as O
from openccm_b200.system_solvers.export import export_system
from openccm_b200.system_solvers.jacobian import jacobian_sparsity
from openccm_b200.workloads import pfr_stiff_workload

PARAMS = dict(n_chains=50, chain_len=20, P=10, t_end=2.0, n_out=11)        # 1000 PFRs x 10 points x 12 species = 120 000 DOF


def main():
    model, net, cp, cmesh = pfr_stiff_workload(workdir=tempfile.mkdtemp(), **PARAMS)
    host = export_system(model, net, cp, cmesh)
    ora = O.build_pfr(net.to_model_network(), cp, cmesh)
    J = jacobian_sparsity(host)
    te = np.linspace(0.0, PARAMS['t_end'], PARAMS['n_out'])
    np.seterr(all='warn', under='ignore')
    sols = {}
    for rtol, atol in ((1e-11, 1e-14), (1e-10, 1e-13), (1e-8, 1e-12)):
        t0 = time.time()
        r = scipy.integrate.solve_ivp(ora.fun, (0.0, PARAMS['t_end']), ora.y0, method='BDF', rtol=rtol, atol=atol, first_step=1e-4,
                                      t_eval=te, jac_sparsity=J)
        assert r.status == 0
        sols[rtol] = r.y
        print(f'rtol {rtol:g}: {time.time() - t0:.0f} s, nfev {r.nfev}, njev {r.njev}, nlu {r.nlu}', flush=True)
    ref = sols[1e-11]                                        # [ndof (reference layout s * n_points + p), T]
    band = 1e-10 + 1e-6 * np.abs(ref)
    agree = float((np.abs(sols[1e-10] - ref) / band).max())
    user = float((np.abs(sols[1e-8] - ref) / band).max())
    print(f'rtol 1e-10 vs 1e-11: {agree:.3f} band units; scipy BDF at 1e-8 / 1e-12: {user:.2f} band units')
    # keep every DOF of a sample of points (all species, all times): 600 points x 12 species x 11 times
    rng = np.random.default_rng(0)
    pts = np.sort(rng.choice(host.n_points, size=600, replace=False))
    idx = (np.arange(host.S)[:, None] * host.n_points + pts[None, :]).reshape(-1)
    pick = lambda y: y[idx].reshape(host.S, len(pts), len(te))
    np.savez_compressed(os.path.join(HERE, 'c5_mid.npz'), t=te, points=pts, ref=pick(ref), scipy_1e8=pick(sols[1e-8]),
                        agreement_band_units=agree, scipy_1e8_band_units=user, checksum=float(np.abs(ref).sum()),
                        **{k: np.int64(v) if isinstance(v, int) else np.float64(v) for k, v in PARAMS.items()})


if __name__ == '__main__':
    main()
