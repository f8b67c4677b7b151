"""One RK45 solve of the 444-PFR network (BASELINE C2) through the persistent one-CTA-per-member kernel, bounded to a few thousand
steps, for an ncu capture of occm_rk45_small:
    ncu --set full --import-source on -k regex:occm_rk45_small -c 1 -o /tmp/prof python scripts/ncu_small.py
"""
import contextlib
import io
import os
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, 'tests'), os.path.join(REPO, 'tests', 'golden')]
import torch
import cases as C
from _util import setup_case
from openccm_b200.system_solvers import solve_system

STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
cp, cmesh, net, mn = setup_case('openfoam_2d_pfr', 'B200-RK45', tempfile.mkdtemp(), rtol=1e-6, atol=1e-10)
t0 = time.time()
with contextlib.redirect_stdout(io.StringIO()):
    try:
        c, t, im, om, info = solve_system('pfr', mn, cp, cmesh, return_info=True, max_steps=STEPS)
    except AssertionError as e:            # the PFR path asserts success like the reference; a step budget ends the run early
        info = {'message': str(e)}
torch.cuda.synchronize()
print(f'{time.time() - t0:.3f} s', {k: info[k] for k in info if k in ('status', 'message', 'n_accepted', 'n_rejected', 'nfev')})
