"""Quick kernel timing on the C4 workload (development aid; bench.py is the contract)."""
import sys, os, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np, torch
from openccm_b200.workloads import cstr_ensemble_workload
from openccm_b200.system_solvers.export import export_system
from openccm_b200.device import DeviceSystem, MemberParams
from openccm_b200 import synthetic as syn

M = int(sys.argv[1]) if len(sys.argv) > 1 else 128
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
t0 = time.time()
model, net, cp, cmesh = cstr_ensemble_workload(n_cstr=N)
host = export_system(model, net, cp, cmesh)
print(f'export {time.time()-t0:.2f}s  ndof {host.ndof} nnz {len(host.row_src)} terms {len(host.tables.terms)} blob {host.tables.pack().nbytes}B')
dev = DeviceSystem(host)
ks, qs, ics = syn.ensemble_sweeps(M, host.tables.n_rxn)
mem = MemberParams(M, torch.as_tensor(ks, device='cuda'), torch.as_tensor(qs, device='cuda'), None)
y0 = torch.as_tensor(host.to_device_layout(host.c0.reshape(-1)), device='cuda')
y = (torch.as_tensor(ics, device='cuda')[:, None] * y0[None, :]).contiguous()
dy = torch.empty_like(y)
for _ in range(3): dev.rhs(y, 0.0, mem, out=dy)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
reps = 20
ev[0].record()
for _ in range(reps): dev.rhs(y, 0.0, mem, out=dy)
ev[1].record(); torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / reps
evals = M * host.ndof
nnz = len(host.row_src)
bytes_alg = 16.0 * evals + 12 * nnz + 4 * (N + 1)
print(f'RHS: {ms:.3f} ms  {evals/ms*1e-6:.1f} G evals/s  alg {bytes_alg/ms*1e-6:.0f} GB/s')
# full solve
te = np.linspace(0, 10, 101)
torch.cuda.synchronize(); t1 = time.time()
res = dev.rk45(y, (0.0, 10.0), rtol=1e-6, atol=1e-10, first_step=1e-4, t_eval=te,
               save_idx=torch.arange(160, dtype=torch.int32, device='cuda'), members=mem)
torch.cuda.synchronize(); dt = time.time() - t1
print(f'RK45 solve M={M}: {dt:.2f}s status {np.unique(res.status)} steps acc {res.n_accepted.min()}..{res.n_accepted.max()} rej {res.n_rejected.max()} nfev sum {res.nfev.sum()}')
print(f'  evals/s in solve {res.nfev.sum()*host.ndof/dt*1e-9:.1f} G   solves/s {M/dt:.1f}')
