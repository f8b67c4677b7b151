"""Wall time of reference-sized cases through solve_system (second call: tables/JIT-free steady state)."""
import sys, os, tempfile, io, contextlib, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, 'tests'), os.path.join(REPO, 'tests', 'golden')]
import numpy as np, torch
import cases as C
from _util import setup_case
from openccm_b200.system_solvers import solve_system
for name, solver in (('cstr_irreversible', 'B200-RK45'), ('cstr_net12_nacl', 'B200-RK45'), ('pfr_net8', 'B200-RK45'), ('cstr_net12_nacl', 'B200-BDF'),
                     ('openfoam_2d_pfr', 'B200-BDF'), ('openfoam_2d_pfr', 'B200-RK45')):
    tmp = tempfile.mkdtemp()
    kw = dict(rtol=1e-6, atol=1e-10) if name == 'openfoam_2d_pfr' else {}
    cp, cmesh, net, mn = setup_case(name, solver, tmp, **kw)
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        with contextlib.redirect_stdout(io.StringIO()):
            c, t, im, om, info = solve_system(C.CASES[name]['model'], mn, cp, cmesh, return_info=True)
        torch.cuda.synchronize(); dt = time.time() - t0
    steps = info['n_accepted'] + info['n_rejected']
    print(f'{name:20s} {solver:10s} {dt:7.3f}s  steps {steps:7d}  nfev {info["nfev"]:8d}  {dt/steps*1e6:8.1f} us/step')
