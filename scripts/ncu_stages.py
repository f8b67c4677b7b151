"""One launch of the plain fused RHS and of each Dormand-Prince stage kernel on the C4 workload, for an ncu capture:
    ncu --set full --clock-control none --import-source on -k regex:occm_pipe -c 8 -o gpurun_out/prof python scripts/ncu_stages.py [members]
"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import torch
from openccm_b200 import synthetic as syn
from openccm_b200.device import DeviceSystem, MemberParams
from openccm_b200.system_solvers.export import export_system
from openccm_b200.workloads import cstr_ensemble_workload

M = int(sys.argv[1]) if len(sys.argv) > 1 else 128
model, net, cp, cmesh = cstr_ensemble_workload(n_cstr=100_000)
host = export_system(model, net, cp, cmesh)
dev = DeviceSystem(host)
ks, qs, ics = syn.ensemble_sweeps(M, host.tables.n_rxn)
mem = MemberParams(M, torch.as_tensor(ks, device='cuda'), torch.as_tensor(qs, device='cuda'), None)
y0 = torch.as_tensor(host.to_device_layout(host.c0.reshape(-1)), device='cuda')
y = (torch.as_tensor(ics, device='cuda')[:, None] * y0[None, :]).contiguous()
dy = dev.rhs(y, 0.0, mem)
with dev.open_rk45(y, (0.0, 10.0), rtol=1e-6, atol=1e-10, first_step=1e-4, t_eval=np.linspace(0, 10, 101),
                   save_idx=torch.arange(160, dtype=torch.int32, device='cuda'), members=mem) as run:
    run.run(max_launch_steps=1)
torch.cuda.synchronize()
print('ok', float(dy.abs().sum()))
