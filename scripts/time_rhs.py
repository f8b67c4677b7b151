"""Plain fused RHS on the C4 workload, CUDA-event time only (development aid for build variants: OCCM_B200_LIB=... / OCCM_B200_SPEC_FLAGS='-D...' python scripts/time_rhs.py)."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from openccm_b200 import synthetic as syn
from openccm_b200.device import DeviceSystem, MemberParams
from openccm_b200.system_solvers.export import export_system
from openccm_b200.workloads import cstr_ensemble_workload

M = int(sys.argv[1]) if len(sys.argv) > 1 else 512
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
model, net, cp, cmesh = cstr_ensemble_workload(n_cstr=N)
host = export_system(model, net, cp, cmesh)
ks, qs, ics = syn.ensemble_sweeps(M, host.tables.n_rxn)
mem = MemberParams(M, torch.as_tensor(ks, device='cuda'), torch.as_tensor(qs, device='cuda'), None)
y0 = torch.as_tensor(host.to_device_layout(host.c0.reshape(-1)), device='cuda')
y = (torch.as_tensor(ics, device='cuda')[:, None] * y0[None, :]).contiguous()
dev = DeviceSystem(host)
dy = torch.empty_like(y)
for _ in range(3):
    dev.rhs(y, 0.0, mem, out=dy)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    dev.rhs(y, 0.0, mem, out=dy)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
alg = M * host.ndof * 16 + 12 * len(host.row_src) + 4 * (N + 1)
print(f'spec={dev.specialized} flags="{os.environ.get("OCCM_B200_SPEC_FLAGS", "")}" {os.environ.get("OCCM_B200_LIB", "")}: rhs {ms:.3f} ms  {alg / ms * 1e-6:.0f} GB/s  checksum {float(dy.double().sum()):.6e}', flush=True)
