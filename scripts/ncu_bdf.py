"""A short slice of the C5 BDF solve (stiff 2.4e7-DOF PFR network) for an ncu launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/bdf_launches.csv python scripts/ncu_bdf.py [attempts]
Without ncu it prints the wall time per attempt of the slice (after an equal warm-up slice)."""
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import torch
from openccm_b200.bdf import BdfRun
from openccm_b200.device import DeviceSystem
from openccm_b200.system_solvers.export import export_system
from openccm_b200.workloads import pfr_stiff_workload

ATTEMPTS = int(sys.argv[1]) if len(sys.argv) > 1 else 40
SKIP = int(sys.argv[2]) if len(sys.argv) > 2 else 400          # attempts integrated first: the slice sits in the bulk of the solve
t_end = 20.0
model, net, cp, cmesh = pfr_stiff_workload(n_chains=500, chain_len=20, P=200, t_end=t_end)
host = export_system(model, net, cp, cmesh)
dev = DeviceSystem(host)
y0 = torch.as_tensor(host.to_device_layout(host.c0.reshape(-1)), device='cuda')[None].contiguous()
te = np.linspace(0.0, t_end, 41)
save = torch.arange(0, host.ndof, max(1, host.ndof // 1000), dtype=torch.int32, device='cuda')
kw = dict(rtol=1e-6, atol=1e-10, first_step=1e-4, t_eval=te, save_idx=save)
with BdfRun(dev, y0, (0.0, t_end), **kw) as run:
    run.run(max_attempts=SKIP)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    t0 = time.perf_counter()
    run.run(max_attempts=ATTEMPTS)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    torch.cuda.profiler.stop()
    run.stats()
    print(f'{ATTEMPTS} attempts after {SKIP}: {1e3 * dt / ATTEMPTS:.2f} ms per attempt; nfev {int(run.nfev[0])} gmres {int(run.n_gmres[0])} '
          f'accepted {int(run.n_acc[0])} rejected {int(run.n_rej[0])}', flush=True)
    print('kernel ms: rhs %.3f  jv %.3f  precond_apply %.3f' % tuple(run.time_kernel(w, 10) for w in (0, 1, 2)), flush=True)
