#!/usr/bin/env python
"""
Benchmark of the reactive-transport hot path (BASELINE.json metric:
"compartment-species RHS evals/s & ensemble ODE solves/s").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--members-per-gpu M] [--impl reference]

Workload (config.workload): C4 of SURVEY.md §8(d) — synthetic 1e5-CSTR network, 10 species, second-order mechanism,
rate-constant / inflow / initial-condition ensemble, t in [0, 10], 101 outputs, rtol 1e-6 / atol 1e-10, RK45.
Weak scaling: every GPU integrates `members_per_gpu` members (512 by default: 4096 members at 8 GPUs, the north-star
configuration); a "step" is one complete ensemble solve of the rank's members.

  value  : compartment-species RHS evaluations per second achieved inside the solves, whole job
           (sum over members of nfev * N * S, divided by the max-over-ranks device time), inputs resident in HBM.
  e2e    : the same metric through the public host API (EnsembleSolver.solve with HOST numpy arrays: upload of the
           network tables, parameters and initial state, solve, download of the recorded outlet concentrations).
  roofline: the dominant kernel = the fused Dormand-Prince stage-6 kernel (largest share of the step: reads the stage state,
           y, K1, K3..K5, writes K6, y_new and the partial error combination), timed live with CUDA events on the launching
           stream; every stage and the plain fused RHS are reported beside it.
  strong : the north-star configuration on THIS many GPUs: 4096 members in total, batched by free HBM on a rank.
  bdf    : (N = 1) C5 — 1e4 PFRs x 200 points x 12 species, stiff 20-reaction mechanism — solved by the device BDF + GMRES;
           live kernel times of the fused RHS / J*v / preconditioner inside Newton; scipy BDF + jac_sparsity on a reduced
           instance of the same workload as its CPU baseline.
  cpu_baseline / --impl reference: the oracle port of the reference's CPU path (numpy/scipy.sparse restatement of
           cstr_system.ddt + scipy.integrate.solve_ivp RK45) on a bounded member sample, all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

import numpy as np

METRIC = "compartment-species RHS evals/s (inside ensemble RK45 solves)"
UNIT = "evals/s"
WORKLOAD = ("C4: synthetic 1e5-CSTR network (ring + random permutation edges), 10 species, 5 second-order reactions, "
            "k/inflow/IC ensemble, t in [0,10], 101 outputs, rtol 1e-6 atol 1e-10, RK45")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=4)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--members-per-gpu', type=int, default=512)
    ap.add_argument('--n-cstr', type=int, default=100_000)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--cpu-sample-members', type=int, default=0, help='members of the CPU baseline sample (0 = auto)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-strong', action='store_true', help='skip the 4096-member strong-scaling solve')
    ap.add_argument('--no-bdf', action='store_true', help='skip the C5 BDF block (N = 1)')
    ap.add_argument('--no-small', action='store_true', help='skip the block of the reference-sized networks C1-C3 (N = 1)')
    ap.add_argument('--strong-members', type=int, default=4096)
    return ap.parse_args()


def build_workload(args):
    from openccm_b200.workloads import cstr_ensemble_workload
    return cstr_ensemble_workload(n_cstr=args.n_cstr)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock, power and clock-event reasons sampled DURING the timed region (B200_PROFILING.md clocks line).

    NVML is polled in-process (nvidia_ml_py, 4 Hz): a separate `nvidia-smi -lms` process takes seconds to attach and
    perturbed the solves it overlapped.  Falls back to one-shot nvidia-smi queries if the binding is missing."""
    REASONS = {0x8: 'hw_slowdown', 0x40: 'hw_thermal_slowdown', 0x20: 'sw_thermal_slowdown', 0x4: 'sw_power_cap'}

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []          # (sm_mhz, power_w, reasons bitmask)
        self.first = 0
        self.max_mhz = None
        self.thread = None
        self._stop = threading.Event()

    def _handle(self):
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get('CUDA_VISIBLE_DEVICES')
        idx = int(vis.split(',')[self.gpu]) if vis and all(v.strip().isdigit() for v in vis.split(',')) else self.gpu
        return pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)

    def start(self):
        if os.environ.get('BENCH_NO_SAMPLER'):
            return
        try:
            nv, h = self._handle()
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            get_reasons = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons

            def poll():
                while not self._stop.is_set():
                    try:
                        self.rows.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)),
                                          nv.nvmlDeviceGetPowerUsage(h) / 1000.0, int(get_reasons(h)), time.perf_counter(),
                                          float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_MEM)),
                                          float(nv.nvmlDeviceGetTemperature(h, nv.NVML_TEMPERATURE_GPU))))
                    except Exception:
                        pass
                    self._stop.wait(0.25)
        except Exception:
            def poll():
                q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active'
                while not self._stop.is_set():
                    try:
                        out = subprocess.run(['nvidia-smi', f'--query-gpu={q}', '--format=csv,noheader,nounits', '-i', str(self.gpu)],
                                             capture_output=True, text=True, timeout=10).stdout.strip().split(',')
                        self.max_mhz = float(out[1])
                        self.rows.append((float(out[0]), float(out[2]), int(out[3].strip(), 16)))
                    except Exception:
                        pass
                    self._stop.wait(0.5)
        self.thread = threading.Thread(target=poll, daemon=True)
        self.thread.start()

    def mark(self):
        """Start of the timed region: earlier samples are discarded."""
        self.first = len(self.rows)
        self.t_mark = time.perf_counter()

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["sampler unavailable"]}
        self._stop.set()
        self.thread.join(timeout=5)
        rows = self.rows[self.first:] or self.rows
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples"]}
        mask = 0
        for r in rows:
            mask |= r[2]
        if os.environ.get('BENCH_TRACE'):
            print('[trace] t_s, sm_mhz, power_w, reasons, mem_mhz, temp_c', file=sys.stderr)
            for r in self.rows:
                if len(r) >= 6:
                    print(f'[trace] {r[3] - getattr(self, "t_mark", r[3]):8.2f} {r[0]:6.0f} {r[1]:7.1f} {r[2]:#x} {r[4]:6.0f} {r[5]:4.0f}', file=sys.stderr)
        return {"sm_mhz": float(np.median([r[0] for r in rows])), "sm_min_mhz": float(min(r[0] for r in rows)),
                "sm_max_mhz": self.max_mhz, "power_w": float(np.median([r[1] for r in rows])),
                "reasons": sorted(n for b, n in self.REASONS.items() if mask & b), "samples": len(rows)}


# ------------------------------------------------------------------------------------------------ CPU / reference arm
def _cpu_member_job(payload):
    """One ensemble member through the oracle port (scipy RK45 around the sparse restated ddt)."""
    (n_cstr, m, k_row, q, ic) = payload
    import scipy.integrate
    try:                                   # one BLAS thread per worker process: the pool already uses every core
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    from oracle import openccm_oracle as O
    model, net, cp, cmesh, sys_o = _cpu_system(n_cstr)
    # apply the member's scales the way the device does: k_scale per reaction, q_scale on transport, ic on y0
    fun = _scaled_fun(sys_o, cp, k_row, q)
    t_eval = O.generate_t_eval(cp)
    t0 = time.perf_counter()
    res = scipy.integrate.solve_ivp(fun, cp.get_list(['SIMULATION', 't_span'], float), sys_o.y0 * ic, method='RK45',
                                    rtol=cp.get_item(['SIMULATION', 'rtol'], float), atol=cp.get_item(['SIMULATION', 'atol'], float),
                                    first_step=cp.get_item(['SIMULATION', 'first_timestep'], float), t_eval=t_eval)
    return res.nfev, time.perf_counter() - t0, res.status


_CPU_CACHE = {}


def _cpu_system(n_cstr):
    if n_cstr not in _CPU_CACHE:
        from oracle import openccm_oracle as O
        from openccm_b200.workloads import cstr_ensemble_workload
        model, net, cp, cmesh = cstr_ensemble_workload(n_cstr=n_cstr, solver='RK45')
        sys_o = O.build_cstr(net.to_model_network(), cp, cmesh, sparse=True)
        _CPU_CACHE[n_cstr] = (model, net, cp, cmesh, sys_o)
    return _CPU_CACHE[n_cstr]


def _scaled_fun(sys_o, cp, k_row, q):
    """ddt of one member: q * (A c + b) + r(c; k * k_row), built from the oracle's pieces."""
    from oracle import openccm_oracle as O
    S, N = sys_o.c_shape
    A = sys_o.extra['A']
    species = cp.get_list(['SIMULATION', 'specie_names'], str)
    rx = O.make_reactions_scaled(cp, species, k_row)
    bcs = sys_o.extra['bcs']

    def fun(t, y):
        c = y.reshape(S, N)
        d = np.ascontiguousarray((A @ c.T).T)
        bcs(t, d)
        d *= q
        rx(c, d)
        return d.ravel()
    return fun


def cpu_baseline(args, n_members_sample, k_scale, q_scale, ic_scale, budget_s=25.0):
    """Oracle port on the host cores: members in a process pool. Returns (evals/s, cores, description)."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    N, S = args.n_cstr, 10
    n = n_members_sample or max(cores, 8)
    jobs = [(args.n_cstr, m, k_scale[m], float(q_scale[m]), float(ic_scale[m])) for m in range(n)]
    t0 = time.perf_counter()
    # spawn: the device arm calls this after CUDA is initialised in the parent (fork would inherit that context)
    with mp.get_context('spawn').Pool(min(cores, n)) as pool:
        out = pool.map(_cpu_member_job, jobs)
    wall = time.perf_counter() - t0
    nfev = sum(o[0] for o in out)
    procs = min(cores, n)
    # members run concurrently, n / procs rounds: parallel solve time = rounds * slowest member (setup excluded)
    rounds = -(-n // procs)
    solve_wall = rounds * max(o[1] for o in out)
    return nfev * N * S / solve_wall, procs, (f"{n} members of the same ensemble, scipy RK45 + scipy.sparse restated ddt, "
                                              f"{procs} processes; solve time {solve_wall:.1f}s (wall incl. setup {wall:.1f}s)")


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the reference itself is pure Python
    with a dense N x N operator that cannot hold N = 1e5) on all host cores, bounded sample per step."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from openccm_b200 import synthetic as syn
    cores = os.cpu_count() or 1
    n = args.cpu_sample_members or cores
    from oracle import openccm_oracle as O
    _, _, cp_ref, _, _ = _cpu_system(args.n_cstr)
    n_rxn = len(O.read_reactions_file(cp_ref.get_item(['SIMULATION', 'reactions_file_path'], str))[0])
    ks, qs, ics = syn.ensemble_sweeps(max(n, 1), n_rxn)
    vals, walls = [], []
    desc = ''
    for i in range(args.warmup + args.steps):
        v, used, desc = cpu_baseline(args, n, ks, qs, ics)
        if i >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "members_per_step": n, "note": "bounded sample of the ensemble per step"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)




# ------------------------------------------------------------------------------------------------ reference-sized networks (C1-C3)
def small_block(args):
    """BASELINE configs[0..2] — one CSTR (A -> B -> C), the 444-PFR and the 76-CSTR networks of examples/OpenFOAM/pipe_with_recirc
    (inputs: tests/golden fixtures made from the unmodified reference) — single solves through the drop-in `solve_system` and
    4096-member ensembles.  These networks fit in shared memory: RK45 runs as one persistent CTA per member (occm_small.cuh)."""
    import contextlib
    import io
    import tempfile
    import torch
    for p in (os.path.join(REPO, 'tests'), os.path.join(REPO, 'tests', 'golden')):
        if p not in sys.path:
            sys.path.insert(0, p)
    import cases as C
    from _util import setup_case
    from openccm_b200.ensemble import EnsembleSolver
    from openccm_b200.system_solvers import solve_system
    out = {}

    def single(name, solver, **kw):
        cp, cmesh, net, mn = setup_case(name, solver, tempfile.mkdtemp(), **kw)
        for _ in range(2):                                   # second call: tables uploaded, kernels loaded
            torch.cuda.synchronize(); t0 = time.perf_counter()
            with contextlib.redirect_stdout(io.StringIO()):
                c, t, im, om, info = solve_system(C.CASES[name]['model'], mn, cp, cmesh, return_info=True)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
        steps = info['n_accepted'] + info['n_rejected']
        return {"seconds": dt, "steps": steps, "nfev": info['nfev'], "us_per_step": 1e6 * dt / max(steps, 1), "status": info['status']}

    def ensemble(name, model, solver, M, rtd=False, **kw):
        cp, cmesh, net, mn = setup_case(name, solver, tempfile.mkdtemp(), **kw)
        es = EnsembleSolver(model, mn, cp, cmesh)
        rng = np.random.default_rng(0)
        qs = rng.uniform(0.5, 2.0, M)
        ks = 10.0 ** rng.uniform(-0.5, 0.5, (M, es.host.tables.n_rxn)) if es.host.tables.n_rxn else None
        for _ in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            r = es.solve(M, q_scale=qs, k_scale=ks, save='outlets', rtd=rtd)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
        return {"members": M, "seconds": dt, "ode_solves_per_s": M / dt, "evals_per_s": float(r.nfev.sum()) * es.host.ndof / dt,
                "steps_per_member": float((r.n_accepted + r.n_rejected).mean()), "ok": bool(np.all(r.status == 1))}

    user = dict(rtol=1e-6, atol=1e-10)
    out["C1_single_cstr_ensemble_rk45"] = ensemble('cstr_irreversible', 'cstr', 'B200-RK45', 4096, **user)
    out["C2_444pfr_rk45"] = single('openfoam_2d_pfr', 'B200-RK45', **user)
    out["C2_444pfr_bdf"] = single('openfoam_2d_pfr', 'B200-BDF', **user)
    out["C2_444pfr_lsoda"] = single('openfoam_2d_pfr', 'B200-LSODA', **user)
    out["C2_444pfr_bdf_ensemble_rtd"] = ensemble('openfoam_2d_pfr', 'pfr', 'B200-BDF', 256, rtd=True, **user)
    out["C3_76cstr_nacl_rk45"] = single('openfoam_2d_cstr_nacl', 'B200-RK45', **user)
    out["C3_76cstr_nacl_ensemble_rk45"] = ensemble('openfoam_2d_cstr_nacl', 'cstr', 'B200-RK45', 4096, **user)
    out["reference"] = ("measured with the unmodified reference in the build container (BASELINE.md, SURVEY 3.5): C2 LSODA 12.9 s "
                        "(5.9 s on the GPU box's host, tests/test_gpu_dropin.py), RK45 96.6 s; C3 LSODA 12.9 s incl. numba JIT")
    return out

# ------------------------------------------------------------------------------------------------ BDF block (C5)
def _bdf_cpu_job(payload):
    """scipy BDF + jac_sparsity (openccm_b200.system_solvers.jacobian) around the oracle's restated pfr ddt, reduced C5."""
    n_chains, chain_len, P, t_end = payload
    import tempfile
    import scipy.integrate
    from oracle import openccm_oracle as O
    from openccm_b200.system_solvers.export import export_system
    from openccm_b200.system_solvers.jacobian import jacobian_sparsity
    from openccm_b200.workloads import pfr_stiff_workload
    model, net, cp, cmesh = pfr_stiff_workload(n_chains=n_chains, chain_len=chain_len, P=P, t_end=t_end, n_out=11, workdir=tempfile.mkdtemp())
    host = export_system(model, net, cp, cmesh)
    ora = O.build_pfr(net.to_model_network(), cp, cmesh)
    J = jacobian_sparsity(host)
    np.seterr(all='warn', under='ignore')
    t0 = time.perf_counter()
    r = scipy.integrate.solve_ivp(ora.fun, (0.0, t_end), ora.y0, method='BDF', rtol=1e-6, atol=1e-10, first_step=1e-4,
                                  t_eval=np.linspace(0.0, t_end, 11), jac_sparsity=J)
    return host.ndof, time.perf_counter() - t0, int(r.nfev), int(r.njev), int(r.nlu), int(r.status)


def bdf_block(args, peak):
    """C5: Np = 1e4 PFRs x 200 points x 12 species = 2.4e7 DOF, 20-reaction mechanism with k from 1e-3 to 1e4, t in [0, 20]."""
    import torch
    from openccm_b200.bdf import BdfRun, device_tolerances
    from openccm_b200.device import DeviceSystem
    from openccm_b200.system_solvers.export import export_system
    from openccm_b200.workloads import pfr_stiff_workload
    t_end = 20.0
    model, net, cp, cmesh = pfr_stiff_workload(n_chains=500, chain_len=20, P=200, t_end=t_end)
    host = export_system(model, net, cp, cmesh)
    dev = DeviceSystem(host)
    y0 = torch.as_tensor(host.to_device_layout(host.c0.reshape(-1)), device='cuda')[None].contiguous()
    te = np.linspace(0.0, t_end, 41)
    save = torch.arange(0, host.ndof, max(1, host.ndof // 1000), dtype=torch.int32, device='cuda')
    rtol, atol = 1e-6, 1e-10                     # the integrator's own tolerances (C5 as measured in round 1); the B200-BDF solver
    kw = dict(rtol=rtol, atol=atol, first_step=1e-4, t_eval=te, save_idx=save)     # name would map them through device_tolerances
    with BdfRun(dev, y0, (0.0, t_end), **kw) as run:                                # warm-up slice
        run.run(max_attempts=30)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with BdfRun(dev, y0, (0.0, t_end), **kw) as run:
        ev0.record()
        run.run()
        ev1.record(); torch.cuda.synchronize()
        run.stats()
        solve_s = ev0.elapsed_time(ev1) * 1e-3
        attempts = int(run.n_acc[0] + run.n_rej[0])
        kern = {name: run.time_kernel(which, 10) for which, name in ((0, 'rhs'), (1, 'jv'), (2, 'precond_apply'))}
        status, nfev, n_gmres = int(run.status[0]), int(run.nfev[0]), int(run.n_gmres[0])
        n_factor, n_apply, n_fail = int(run.n_precond_factor[0]), int(run.n_precond_apply[0]), int(run.n_newton_fail[0])
        mode = run.precond_mode
    ndof, S = host.ndof, host.S
    newton_its = nfev - n_gmres - 1                 # RHS calls that are Newton residuals (the others are Jacobian actions)
    alg = {'rhs': 16.0 * ndof, 'jv': 24.0 * ndof, 'precond_apply': (24.0 + 4.0 * S) * ndof}      # bytes per launch (block inverses stored as float)
    roof = {k: {"ms_per_launch": v, "achieved_gbs": alg[k] / (v * 1e-3) / 1e9, "frac": alg[k] / (v * 1e-3) / 1e9 / peak} for k, v in kern.items()}
    out = {"workload": "C5: 1e4 PFRs (500 recycle chains of 20) x 200 points x 12 species = 2.4e7 DOF, 20 reactions with k in [1e-3, 1e4], "
                       "t in [0, 20], rtol 1e-6 atol 1e-10, BDF + block-preconditioned GMRES(30), one member",
           "status": status, "solve_s": solve_s, "solves_per_s": 1.0 / solve_s, "attempts": attempts, "ms_per_attempt": 1e3 * solve_s / max(attempts, 1),
           "accepted": int(run.n_acc[0]), "newton_failures": n_fail, "rhs_calls": nfev, "gmres_iterations": n_gmres,
           "gmres_per_newton": n_gmres / max(newton_its, 1), "newton_per_attempt": newton_its / max(attempts, 1),
           "precond": {"mode": mode, "applications": n_apply, "rebuilds": n_factor},
           "dof_steps_per_s": ndof * float(run.n_acc[0]) / solve_s,
           "roofline": {"bound": "hbm", "peak": peak, "unit": "GB/s", "kernels": roof}}
    del dev, y0
    torch.cuda.empty_cache()
    if not args.no_cpu_baseline:
        try:
            import multiprocessing as mp
            with mp.get_context('spawn').Pool(1) as pool:
                nd, sec, nf, nj, nl, st = pool.map(_bdf_cpu_job, [(5, 10, 20, 5.0)])[0]
            out["cpu_baseline"] = {"kind": "port", "cores": 1, "ndof": nd, "seconds": sec, "nfev": nf, "njev": nj, "nlu": nl, "status": st,
                                   "sample": "reduced C5 (5 chains x 10 PFRs x 20 points x 12 species, t in [0, 5]): scipy BDF with "
                                             "jac_sparsity around the oracle's restated pfr ddt, rtol 1e-6 atol 1e-10, one core"}
        except Exception as exc:
            out["cpu_baseline"] = {"failed": repr(exc)}
    return out


# ------------------------------------------------------------------------------------------------ device arm
_REAL_STDOUT = None


def _claim_stdout():
    """Everything any library writes to fd 1 from here on (NCCL prints its version banner there) goes to stderr; the one JSON
    line of the contract is written to the original stdout by `_emit`."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def _emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    args = parse_args()
    _claim_stdout()
    if args.impl == 'reference':
        run_reference_arm(args)
        return
    import torch
    import torch.distributed as dist
    from openccm_b200 import _lib, synthetic as syn
    from openccm_b200.device import MemberParams
    from openccm_b200.ensemble import EnsembleSolver

    world_size = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    if world_size > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    lib = _lib.load()
    # nvidia-smi takes seconds to attach (NVML init) and perturbs the GPU while it does (measured: +15..40 % on the solves
    # that overlap it): start it before anything else and only keep the samples taken inside the timed region
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    model, net, cp, cmesh = build_workload(args)
    solver = EnsembleSolver(model, net, cp, cmesh)
    host = solver.host
    N, S, ndof = host.n_models, host.S, host.ndof
    M_loc = args.members_per_gpu
    M_glob = M_loc * world_size
    ks_all, qs_all, ics_all = syn.ensemble_sweeps(M_glob, host.tables.n_rxn)
    # members are dealt to the ranks by a stiffness estimate (boustrophedon over the sorted list): every rank gets the same mix
    from openccm_b200.ensemble import balanced_order, stiffness_estimate
    mine = balanced_order(stiffness_estimate(host, M_glob, ks_all, qs_all), world_size)[rank * M_loc:(rank + 1) * M_loc]
    ks, qs, ics = ks_all[mine], qs_all[mine], ics_all[mine]
    save_idx = solver.outlet_save_idx()
    te = np.asarray(solver.t_eval)
    dev = solver.upload()
    device = dev.dev
    dev_specialized = bool(dev.specialized)

    # ---- device-resident inputs for `value`
    mem = MemberParams(M_loc, torch.as_tensor(ks, device=device), torch.as_tensor(qs, device=device), None)
    c0_dev = torch.as_tensor(host.to_device_layout(host.c0.reshape(-1)), device=device)
    y0 = (torch.as_tensor(ics, device=device)[:, None] * c0_dev[None, :]).contiguous()
    save_dev = torch.as_tensor(save_idx, dtype=torch.int32, device=device)

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_solve():
        t_a = time.perf_counter()
        with dev.open_rk45(y0, solver.t_span, rtol=solver.rtol, atol=solver.atol, first_step=solver.first_step,
                           t_eval=te, save_idx=save_dev, members=mem) as run:
            t_b = time.perf_counter()
            run.run()
            t_c = time.perf_counter()
            run.stats()
            t_d = time.perf_counter()
            if os.environ.get('BENCH_TRACE'):
                print(f'[trace] solve: open {1e3 * (t_b - t_a):.1f} ms, run {1e3 * (t_c - t_b):.1f} ms, stats {1e3 * (t_d - t_c):.1f} ms', file=sys.stderr)
            return int(run.nfev.sum()), int(run.n_acc.sum() + run.n_rej.sum()), np.array(run.status)

    # The solves run at the 1000 W power cap; the power controller needs ~5 s of load to settle (measured: the solve that
    # overlaps the transient is 15-20 % slower, SM clock dips to 1700 MHz).  At least 5 untimed solves, more if asked for.
    warmup_eff = max(args.warmup, 5)
    for _ in range(warmup_eff):
        one_solve()
    barrier()
    sampler.mark()
    launches0 = lib.occm_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    nfev_total = 0
    step_events = [ev0]
    for _ in range(args.steps):
        nfev, attempts, status = one_solve()
        nfev_total += nfev
        step_events.append(torch.cuda.Event(enable_timing=True))
        step_events[-1].record()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    step_ms = [round(a.elapsed_time(b), 2) for a, b in zip(step_events[:-1], step_events[1:])]
    launches = lib.occm_launch_count() - launches0
    assert np.all(status == 1), f'members failed: {np.unique(status)}'
    t_ms = torch.tensor([ms], dtype=torch.float64, device=device)
    work = torch.tensor([float(nfev_total)], dtype=torch.float64, device=device)
    if world_size > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
    ms_max = float(t_ms.item())
    evals = float(work.item()) * N * S
    value = evals / (ms_max * 1e-3)
    solves_per_s = M_glob * args.steps / (ms_max * 1e-3)

    # ---- e2e through the host API (host numpy in, host numpy out), copies inside the timed region
    solver_e2e = EnsembleSolver(model, net, cp, cmesh, host=host)
    def e2e_step():
        solver_e2e.dev = None                      # re-upload tables every step
        # global host arrays in; the solver takes this rank's contiguous member block (no collective on the data path)
        # (the NCCL all_gather of the recorded outlet concentrations + RTD arrays and the all_reduce of the RTD moment statistics
        #  — the only collectives of the path — are inside the timed region)
        r = solver_e2e.solve(M_glob, k_scale=ks_all, q_scale=qs_all, ic_scale=ics_all, save=save_idx, gather='root', rtd=True)
        return r
    for _ in range(min(args.warmup, 1)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    nfev_e2e = 0
    gather_ms, reduce_ms = [], []
    for _ in range(args.steps):
        r = e2e_step()
        nfev_e2e += int(r.nfev.sum())            # rank 0: the counters of ALL members (gathered); other ranks: their own
        gather_ms.append(1e3 * r.timing.get('gather_s', 0.0)); reduce_ms.append(1e3 * r.timing.get('reduce_s', 0.0))
    barrier()
    e2e_s = time.perf_counter() - t0
    t_e = torch.tensor([e2e_s], dtype=torch.float64, device=device)
    if world_size > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = float(nfev_e2e) * N * S / float(t_e.item())      # used on rank 0 only, where nfev_e2e covers every member
    blob_bytes = host.tables.pack().nbytes
    h2d = (host.row_ptr.nbytes + host.row_src.nbytes + host.row_w.nbytes) * 2 + blob_bytes + host.c0.nbytes + \
        ks.nbytes + qs.nbytes + ics.nbytes + save_idx.nbytes + te.nbytes
    # rank 0 downloads the gathered members: recorded outlets [M, T, n_save], RTD [M, S, T, 3], moments, counters (the other ranks
    # their own 1 / N of it)
    d2h = M_glob * (len(te) * len(save_idx) * 8 + S * len(te) * 3 * 8 + S * 3 * 8 + 4 + 3 * 8)
    clocks = sampler.stop() if rank == 0 else None

    # ---- roofline of the dominant kernel (rank 0): fused stage-6 kernel (largest share of the step), live CUDA-event timing
    roof = None
    plain = None
    STAGE_BYTES = {2: 32, 3: 48, 4: 56, 5: 64, 6: 72, 7: 32}       # algorithmic bytes per element
    if rank == 0:
        with dev.open_rk45(y0, solver.t_span, rtol=solver.rtol, atol=solver.atol, first_step=solver.first_step,
                           t_eval=te, save_idx=save_dev, members=mem) as run:
            run.run(max_launch_steps=2)            # buffers hold a genuine mid-solve state
            run.time_stage(6, 3)
            ms6 = run.time_stage(6, 10)
            stage_ms = {}
            for s_ in (2, 3, 4, 5, 6, 7):
                run.time_stage(s_, 2)              # two untimed launches of the stage, then five timed ones
                stage_ms[s_] = run.time_stage(s_, 5)
        nnz = len(host.row_src)
        csr_bytes = 12 * nnz + 4 * (N + 1)
        alg6 = M_loc * ndof * STAGE_BYTES[6] + csr_bytes          # 6 vector reads + 3 vector writes per element, CSR once
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        peak = float(peaks.get('hbm_gbs', 6650.0))
        ach = alg6 / (ms6 * 1e-3) / 1e9
        traffic = None          # dram__bytes_read + dram__bytes_write of this kernel from the committed ncu capture (same workload)
        try:
            tr = json.load(open(os.path.join(REPO, 'profiles', 'traffic_r02.json')))
            if tr.get('members_per_gpu') == M_loc and tr.get('n_cstr') == N:
                traffic = tr['traffic_bytes_per_launch']
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": "occm_pipe<CSTR, stage 6, K> (fused Dormand-Prince stage: RHS of the stage state + K6, y_new and the "
                                          "partial error combination; TMA bulk-copy ring)",
                "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": traffic, "peak_source": "measured" if peaks else "fallback",
                "ms_per_launch": ms6, "alg_bytes_per_launch": alg6,
                "stage_ms": stage_ms,
                "stage_frac": {s: (M_loc * ndof * STAGE_BYTES[s] + csr_bytes) / (v * 1e-3) / 1e9 / peak for s, v in stage_ms.items()},
                # all six stage kernels of a step attempt together: 304 B per element over their summed time
                "all_stages_frac": (M_loc * ndof * 304.0 + 6 * csr_bytes) / (sum(stage_ms.values()) * 1e-3) / 1e9 / peak}
        # plain fused RHS (16 B per compartment-species eval)
        dy = torch.empty_like(y0)
        for _ in range(3):
            dev.rhs(y0, 0.0, mem, out=dy)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            dev.rhs(y0, 0.0, mem, out=dy)
        e1.record(); torch.cuda.synchronize()
        ms_r = e0.elapsed_time(e1) / 10
        alg_r = M_loc * ndof * 16 + csr_bytes
        plain = {"kernel": "occm_spec_stage0 (mechanism-specialised, thread per point)" if dev.specialized else "occm_pipe<CSTR, stage 0, K>",
                 "ms_per_launch": ms_r, "evals_per_s": M_loc * ndof / (ms_r * 1e-3), "achieved_gbs": alg_r / (ms_r * 1e-3) / 1e9,
                 "frac": alg_r / (ms_r * 1e-3) / 1e9 / peak}
        del dy

    # ---- strong scaling: the north-star ensemble (4096 members in total) on this many GPUs, batched by free HBM per rank
    strong = None
    del y0, mem
    torch.cuda.empty_cache()
    if not args.no_strong:
        M_s = args.strong_members
        ks_s, qs_s, ics_s = syn.ensemble_sweeps(M_s, host.tables.n_rxn)
        solver_s = EnsembleSolver(model, net, cp, cmesh, host=host)
        solver_s.dev = dev
        bpr = -(-(M_s // world_size) // solver_s.members_per_batch(len(save_idx), len(te)))
        barrier()
        t0 = time.perf_counter()
        r_s = solver_s.solve(M_s, k_scale=ks_s, q_scale=qs_s, ic_scale=ics_s, save=save_idx, gather='root', rtd=True)
        barrier()
        t_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
        if world_size > 1:
            dist.all_reduce(t_s, op=dist.ReduceOp.MAX)
        assert np.all(r_s.status == 1)
        lo_s, hi_s = r_s.local_range
        strong = {"scaling": "strong", "members_total": M_s, "members_per_rank": hi_s - lo_s, "seconds": float(t_s.item()),
                  "ode_solves_per_s": M_s / float(t_s.item()), "evals_per_s": float(r_s.nfev.sum()) * N * S / float(t_s.item()),
                  "batches_per_rank": bpr,
                  "gather_ms": 1e3 * r_s.timing.get('gather_s', 0.0), "reduce_ms": 1e3 * r_s.timing.get('reduce_s', 0.0),
                  "note": "host API (EnsembleSolver.solve): host arrays in, gathered outlets + RTD out, one timed solve"}
        del r_s

    # ---- BDF + GMRES on C5 (N = 1 only: a single network does not shard)
    bdf = None
    if rank == 0 and world_size == 1 and not args.no_bdf:
        del solver_e2e, solver, dev
        torch.cuda.empty_cache()
        try:
            bdf = bdf_block(args, peak if rank == 0 else 6650.0)
        except Exception as exc:           # never take the headline down
            bdf = {"failed": repr(exc)}

    small = None
    if rank == 0 and world_size == 1 and not args.no_small:
        try:
            small = small_block(args)
        except Exception as exc:
            small = {"failed": repr(exc)}

    cpu = None
    if rank == 0 and world_size == 1 and not args.no_cpu_baseline:
        try:
            v, used, desc = cpu_baseline(args, args.cpu_sample_members, ks_all, qs_all, ics_all)
            cpu = {"value": v, "unit": UNIT, "cores": used, "kind": "port", "sample": desc}
        except Exception as exc:           # the baseline must never take the bench line down
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {exc!r}"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": args.steps, "warmup": warmup_eff,
                "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": WORKLOAD, "members_per_gpu": M_loc, "members_total": M_glob, "n_cstr": N, "n_species": S,
                           "l2": f"inputs larger than L2: {11 * M_loc * ndof * 8 / 2**30:.1f} GiB of state vectors per GPU",
                           "parallelism": f"members dealt to {world_size} GPU(s) by a stiffness estimate, no data-path collective",
                           "kernels": ("mechanism-specialised image for the plain RHS and stages 2, 7 (occm_spec_stage*), TMA-pipelined "
                                       "table-driven kernels for stages 3-6 (occm_pipe)") if dev_specialized else
                                      "TMA-pipelined table-driven kernels (occm_pipe)"},
                "ode_solves_per_s": solves_per_s, "rhs_evals_total": evals, "step_ms": step_ms,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "solves_per_s": M_glob * args.steps / float(t_e.item()),
                        "collectives": "gather to rank 0 (outlets, RTD, counters) + all_reduce(RTD moment statistics) inside the timed region",
                        "gather_ms": float(np.mean(gather_ms)), "reduce_ms": float(np.mean(reduce_ms))},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "plain_rhs": plain, "cpu_baseline": cpu,
                "strong": strong, "bdf": bdf, "small_networks": small}
        _emit(line)
    if world_size > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
