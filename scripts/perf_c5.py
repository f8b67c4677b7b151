"""C5 (stiff PFR network, BDF + GMRES) timing aid."""
import sys, os, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np, torch
from openccm_b200.workloads import pfr_stiff_workload
from openccm_b200.system_solvers.export import export_system
from openccm_b200.device import DeviceSystem
from openccm_b200.bdf import solve_bdf

n_chains = int(sys.argv[1]) if len(sys.argv) > 1 else 50
P = int(sys.argv[2]) if len(sys.argv) > 2 else 200
M = int(sys.argv[3]) if len(sys.argv) > 3 else 1
t_end = float(sys.argv[4]) if len(sys.argv) > 4 else 20.0
t0 = time.time()
model, net, cp, cmesh = pfr_stiff_workload(n_chains=n_chains, P=P, t_end=t_end)
host = export_system(model, net, cp, cmesh)
print(f'export {time.time()-t0:.2f}s  Np {host.n_models} P {host.P} S {host.S} ndof {host.ndof} terms {len(host.tables.terms)} n_rxn {host.tables.n_rxn}')
dev = DeviceSystem(host)
y0 = torch.as_tensor(host.to_device_layout(host.c0.reshape(-1)), device='cuda')[None].repeat(M, 1).contiguous()
te = np.linspace(0, t_end, 41)
save = torch.arange(0, host.ndof, max(1, host.ndof // 1000), dtype=torch.int32, device='cuda')
torch.cuda.synchronize(); t1 = time.time()
res = solve_bdf(dev, y0, (0.0, t_end), rtol=1e-6, atol=1e-10, first_step=1e-4, t_eval=te, save_idx=save, max_steps=int(os.environ.get('MAXSTEPS', '0')), lin_tol_factor=float(os.environ.get('LINTOL', '0')))   # OCCM_B200_BDF_PRECOND=1: on-the-fly blocks
torch.cuda.synchronize(); dt = time.time() - t1
print(f'BDF M={M}: {dt:.2f}s status {res.status} acc {res.n_accepted} rej {res.n_rejected} nfev {res.nfev} gmres its {res.n_gmres} newton fails {res.n_newton_fail}')
print(f'  preconditioner applications {res.n_precond_apply} rebuilds {res.n_precond_factor}')
print(f'  RHS evals/s {res.nfev.sum()*host.ndof/dt*1e-9:.2f} G   max y {float(res.y.abs().max()):.3g}')
ck = res.y[:, -1].double().sum().item()
print(f'  checksum(final saved rows) {ck:.12e}')
