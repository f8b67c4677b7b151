"""A/B of the TMA-pipelined kernels (occm_pipe.cuh) against the register-prefetching fast kernels (occm_fast.cuh) on the C4
workload: bit-level agreement of the RHS and of a short RK45 solve, then live CUDA-event times of the plain RHS, of the six
Dormand-Prince stage kernels and of a complete ensemble solve.  Development aid; bench.py is the contract.

    python scripts/perf_pipe.py [members] [n_cstr] [depth caps, comma separated]
"""
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import torch
from openccm_b200 import synthetic as syn
from openccm_b200.device import DeviceSystem, MemberParams
from openccm_b200.system_solvers.export import export_system
from openccm_b200.workloads import cstr_ensemble_workload

M = int(sys.argv[1]) if len(sys.argv) > 1 else 128
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
DEPTHS = [int(x) for x in sys.argv[3].split(',')] if len(sys.argv) > 3 else [0]
PEAK = 6551.7

model, net, cp, cmesh = cstr_ensemble_workload(n_cstr=N)
host = export_system(model, net, cp, cmesh)
ks, qs, ics = syn.ensemble_sweeps(M, host.tables.n_rxn)
mem = MemberParams(M, torch.as_tensor(ks, device='cuda'), torch.as_tensor(qs, device='cuda'), None)
y0 = torch.as_tensor(host.to_device_layout(host.c0.reshape(-1)), device='cuda')
y = (torch.as_tensor(ics, device='cuda')[:, None] * y0[None, :]).contiguous()
te = np.linspace(0, 10, 101)
save = torch.arange(160, dtype=torch.int32, device='cuda')
nnz = len(host.row_src)
csr = 12 * nnz + 4 * (N + 1)
BYTES = {2: 32, 3: 48, 4: 56, 5: 64, 6: 72, 7: 32}


ROW_AB = len(sys.argv) > 4 and sys.argv[4] == 'row'      # A/B of the mechanism-specialised light-stage kernels against the pipe kernels


def make(pipe: bool, depth: int = 0, row: bool = True) -> DeviceSystem:
    os.environ.pop('OCCM_B200_NO_PIPE', None)
    os.environ.pop('OCCM_B200_SPECIALIZE', None)
    if not row:
        os.environ['OCCM_B200_SPECIALIZE'] = '0'
    os.environ.pop('OCCM_B200_PIPE_DEPTH', None)
    if not pipe:
        os.environ['OCCM_B200_NO_PIPE'] = '1'
    if depth:
        os.environ['OCCM_B200_PIPE_DEPTH'] = str(depth)
    return DeviceSystem(host)


def measure(dev: DeviceSystem, label: str):
    dy = torch.empty_like(y)
    for _ in range(3):
        dev.rhs(y, 0.0, mem, out=dy)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        dev.rhs(y, 0.0, mem, out=dy)
    e1.record(); torch.cuda.synchronize()
    ms_r = e0.elapsed_time(e1) / 10
    alg = M * host.ndof * 16 + csr
    out = {'rhs_ms': ms_r, 'rhs_frac': alg / ms_r * 1e-6 / PEAK}
    with dev.open_rk45(y, (0.0, 10.0), rtol=1e-6, atol=1e-10, first_step=1e-4, t_eval=te, save_idx=save, members=mem) as run:
        run.run(max_launch_steps=2)
        st = {}
        for s in (2, 3, 4, 5, 6, 7):
            run.time_stage(s, 2)
            st[s] = run.time_stage(s, 6)
    tot = sum(st.values())
    out['stage_ms'] = {k: round(v, 3) for k, v in st.items()}
    out['stage_frac'] = {k: round((M * host.ndof * BYTES[k] + csr) / v * 1e-6 / PEAK, 3) for k, v in st.items()}
    out['all_stages_frac'] = (M * host.ndof * 304.0 + 6 * csr) / tot * 1e-6 / PEAK
    torch.cuda.synchronize(); t1 = time.time()
    res = dev.rk45(y, (0.0, 10.0), rtol=1e-6, atol=1e-10, first_step=1e-4, t_eval=te, save_idx=save, members=mem)
    torch.cuda.synchronize(); dt = time.time() - t1
    out['solve_s'] = dt
    out['evals_per_s'] = float(res.nfev.sum()) * host.ndof / dt
    print(f'[{label}] rhs {ms_r:.3f} ms ({out["rhs_frac"]:.3f}) stages {out["stage_ms"]} frac {out["stage_frac"]} all {out["all_stages_frac"]:.3f} '
          f'solve {dt:.2f}s {out["evals_per_s"] / 1e9:.1f} G evals/s status {np.unique(res.status)}', flush=True)
    return dy, res


dev_f = make(True, 0, False) if ROW_AB else make(False)
dy_f, res_f = measure(dev_f, 'pipe, table-driven' if ROW_AB else 'fast')
for d in DEPTHS:
    dev_p = make(True, d)
    dy_p, res_p = measure(dev_p, ('specialised' if dev_p.specialized else 'pipe') + f' depth<={d or "max"}')
    print('   RHS bitwise equal:', bool(torch.equal(dy_f, dy_p)), ' max |diff|', float((dy_f - dy_p).abs().max()),
          ' solve outputs max |diff|:', float((res_f.y - res_p.y).abs().max()), ' nfev equal:', bool(np.array_equal(res_f.nfev, res_p.nfev)), flush=True)
    del dev_p, dy_p, res_p
