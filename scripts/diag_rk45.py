import sys, os, tempfile
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, 'tests'), os.path.join(REPO, 'tests', 'golden')]
import numpy as np
import cases as C
from _util import setup_case, load_golden
from openccm_b200.system_solvers import solve_system
import io, contextlib
for name in sorted(C.CASES):
    if 'RK45' not in C.CASES[name].get('solvers', ('RK45', 'LSODA')): continue
    g = load_golden(name)
    tmp = tempfile.mkdtemp()
    for mode in ('teval', 'all'):
        if mode == 'all' and 't_RK45_all' not in g: continue
        cp, cmesh, net, mn = setup_case(name, 'B200-RK45', tmp, **({'t_eval': 'all'} if mode == 'all' else {}))
        with contextlib.redirect_stdout(io.StringIO()):
            c, t, im, om, info = solve_system(C.CASES[name]['model'], mn, cp, cmesh, return_info=True)
        cr = g['c_RK45_all' if mode == 'all' else 'c_RK45']; tr = g['t_RK45_all' if mode == 'all' else 't_RK45']
        if c.shape != cr.shape:
            print(f'{name:24s} {mode:5s} SHAPE {c.shape} vs {cr.shape} nfev {info["nfev"]}'); continue
        band = np.abs(c - cr) / (1e-10 + 1e-6 * np.abs(cr))
        print(f'{name:24s} {mode:5s} max|dc| {np.abs(c-cr).max():.2e} band(1e-6/1e-10) {band.max():.2e} '
              f'max rel dt {np.max(np.abs(t-tr)/np.maximum(np.abs(tr),1e-300)):.1e} nfev {info["nfev"]} ref {int(g["nfev_RK45_all" if mode=="all" else "nfev_RK45"])}')
