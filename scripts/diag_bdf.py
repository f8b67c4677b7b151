import sys, os, tempfile, io, contextlib, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, 'tests'), os.path.join(REPO, 'tests', 'golden')]
import numpy as np
import cases as C
from _util import setup_case, load_golden
from openccm_b200.system_solvers import solve_system
names = sys.argv[1:] or ['cstr_irreversible', 'cstr_net12_nacl', 'pfr_single', 'pfr_net8', 'openfoam_2d_pfr']
for name in names:
    g = load_golden(name); tmp = tempfile.mkdtemp()
    for rtol, atol in ((1e-6, 1e-10), (1e-9, 1e-12)):
        cp, cmesh, net, mn = setup_case(name, 'B200-BDF', tmp, rtol=rtol, atol=atol)
        t0 = time.time()
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                c, t, im, om, info = solve_system(C.CASES[name]['model'], mn, cp, cmesh, return_info=True)
        except AssertionError as e:
            print(f'{name:22s} rtol {rtol:g} FAILED {e}'); continue
        ref = g['c_LSODA']
        band = np.abs(c - ref) / (1e-10 + 1e-6 * np.abs(ref))
        print(f'{name:22s} rtol {rtol:g} status {info["status"]} max|dc| {np.abs(c-ref).max():.2e} band {band.max():.2e} nfev {info["nfev"]} '
              f'acc {info["n_accepted"]} rej {info["n_rejected"]} time {time.time()-t0:.2f}s  (ref LSODA nfev {int(g["nfev_LSODA"])})')
