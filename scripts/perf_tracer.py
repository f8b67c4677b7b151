"""Tracer (S = 1, no reactions: the RTD use case) on the C4 network: the generic kernels' throughput (development aid)."""
import sys, os, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np, torch
from openccm_b200 import synthetic as syn
from openccm_b200.workloads import _cmesh, _config
from openccm_b200.system_solvers.export import export_system
from openccm_b200.device import DeviceSystem, MemberParams

M = int(sys.argv[1]) if len(sys.argv) > 1 else 512
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
S = int(sys.argv[3]) if len(sys.argv) > 3 else 1
net = syn.random_network(N, n_io=16, seed=0)
species = syn.species_names(S)
cp = _config('cstr', species, 'None', t_span=(0, 10.0), t_eval='0.1, linear', ic_value=0.0)
cmesh = _cmesh(cp, N, 100)
host = export_system('cstr', net, cp, cmesh)
dev = DeviceSystem(host)
qs = np.random.default_rng(1).uniform(0.5, 2.0, M)
mem = MemberParams(M, None, torch.as_tensor(qs, device='cuda'), None)
y = torch.rand((M, host.ndof), dtype=torch.float64, device='cuda')
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
print(f'S={S} N={N} M={M}: plain RHS {ms:.3f} ms  {evals/ms*1e-6:.1f} G evals/s  algorithmic {16.0*evals/ms*1e-6:.0f} GB/s')
te = np.linspace(0, 10, 101)
torch.cuda.synchronize(); t1 = time.time()
res = dev.rk45(y * 0, (0.0, 10.0), rtol=1e-6, atol=1e-10, first_step=1e-4, t_eval=te,
               save_idx=torch.arange(min(160, host.ndof), dtype=torch.int32, device='cuda'), members=mem)
torch.cuda.synchronize(); dt = time.time() - t1
print(f'RK45 solve: {dt:.2f}s status {np.unique(res.status)} nfev {res.nfev.min()}..{res.nfev.max()}  evals/s {res.nfev.sum()*host.ndof/dt*1e-9:.1f} G  '
      f'(304 B/element/step -> {res.nfev.sum()/6*host.ndof*304/dt*1e-9:.0f} GB/s algorithmic)')
