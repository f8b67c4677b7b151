"""
Batched ensembles (rate-constant, inflow, boundary-value and initial-condition sweeps) sharded member-wise over GPUs.

The reference has no ensemble API: sweeps are Python loops that call ``run`` again
(/root/reference/examples/simple_reactors/cstr/irreversible/cstr_analysis.py:60-65).  Here all members of a rank
advance together in the batched kernels, each with its own time and step size; ranks never exchange data during
the integration.  ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) is used only to gather the recorded
observables and to reduce RTD statistics.
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from typing import Dict, Optional, Sequence, Tuple, Union

import numpy as np
import torch
import torch.distributed as dist

from .system_solvers.export import HostSystem, export_system
from .system_solvers.helper_functions import generate_t_eval


# ------------------------------------------------------------------------------------------------ sharding helpers
def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n_members: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block of members owned by ``rank`` (sizes differ by at most one)."""
    return n_members * rank // world_size, n_members * (rank + 1) // world_size


def stiffness_estimate(host: HostSystem, n_members: int, k_scale: Optional[np.ndarray] = None,
                       q_scale: Optional[np.ndarray] = None) -> np.ndarray:
    """Fastest rate of each member, up to a constant: |transport diagonal|_max * q_scale + max_r |k_r| * k_scale[m, r]
    (adaptive step counts differ per member; members are ordered by this before they are dealt to ranks)."""
    rows = np.repeat(np.arange(len(host.row_ptr) - 1), np.diff(host.row_ptr))
    if host.model == 'pfr':
        tmax = float(np.max(np.abs(host.a))) if host.a is not None and len(host.a) else 0.0
    else:
        on_diag = host.row_src == rows
        tmax = float(np.max(np.abs(host.row_w[on_diag]))) if on_diag.any() else 0.0
    est = tmax * (np.ones(n_members) if q_scale is None else np.asarray(q_scale, dtype=np.float64)[:n_members])
    terms = host.tables.terms
    if terms:
        coeff = np.array([abs(t.coeff) for t in terms]); rxn = np.array([t.rxn for t in terms])
        if k_scale is None:
            est = est + coeff.max()
        else:
            est = est + (np.asarray(k_scale, dtype=np.float64)[:n_members][:, rxn] * coeff[None, :]).max(axis=1)
    return est


def balanced_order(estimate: np.ndarray, world_size: int) -> np.ndarray:
    """Permutation of the members such that rank r owns ``perm[shard_bounds(n, r, world_size)]``: members sorted by decreasing
    stiffness estimate and dealt to the ranks in boustrophedon order (0..w-1, w-1..0, ...), so every rank gets the same mix
    and, inside a rank, neighbours in the batch have similar step counts.  Deterministic: every rank computes the same order."""
    n = len(estimate)
    order = np.argsort(-np.asarray(estimate, dtype=np.float64), kind='stable')
    cap = [shard_bounds(n, r, world_size)[1] - shard_bounds(n, r, world_size)[0] for r in range(world_size)]
    buckets = [[] for _ in range(world_size)]
    r, step = 0, 1
    for mbr in order.tolist():
        while len(buckets[r]) >= cap[r]:                      # (sizes differ by at most one)
            r += step
            if r == world_size or r < 0:
                step = -step; r += step
        buckets[r].append(mbr)
        r += step
        if r == world_size or r < 0:
            step = -step; r += step
    return np.concatenate([np.asarray(b, dtype=np.int64) for b in buckets]) if n else np.zeros(0, dtype=np.int64)


def gather_members(local: torch.Tensor, n_members: int, root: Optional[int] = None) -> Optional[torch.Tensor]:
    """Gather per-member rows (axis 0) from every rank, in rank order. ``root=None``: all_gather, every rank gets all rows;
    ``root=r``: gather to rank r only (the others return None).  Works on CUDA (NCCL) and CPU (gloo)."""
    rank, ws = world()
    if ws == 1:
        return local
    per_rank = [shard_bounds(n_members, r, ws) for r in range(ws)]
    n_max = max(hi - lo for lo, hi in per_rank)
    pad = torch.zeros((n_max,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    if root is None:
        parts = [torch.empty_like(pad) for _ in range(ws)]
        dist.all_gather(parts, pad)
    else:
        parts = [torch.empty_like(pad) for _ in range(ws)] if rank == root else None
        dist.gather(pad, parts, dst=root)
        if rank != root:
            return None
    return torch.cat([parts[r][:hi - lo] for r, (lo, hi) in enumerate(per_rank)], dim=0)


def to_host_pinned(x: Optional[torch.Tensor]) -> Optional[np.ndarray]:
    """Device -> host through a pinned staging buffer (a pageable .cpu() of the gathered outputs ran at a quarter of the PCIe rate)."""
    if x is None:
        return None
    if not x.is_cuda or x.numel() == 0:
        return x.cpu().numpy()
    buf = torch.empty(x.shape, dtype=x.dtype, pin_memory=True)
    buf.copy_(x, non_blocking=True)
    torch.cuda.current_stream(x.device).synchronize()
    return buf.numpy()


def reduce_statistics(local: torch.Tensor, n_members: int) -> Dict[str, torch.Tensor]:
    """Ensemble mean / min / max / variance over the member axis (axis 0) across all ranks via all_reduce."""
    rank, ws = world()
    s1 = local.sum(0); s2 = (local * local).sum(0)
    if local.shape[0]:
        mn = local.min(0).values; mx = local.max(0).values
    else:
        mn = torch.full_like(s1, float('inf')); mx = torch.full_like(s1, float('-inf'))
    if ws > 1:
        dist.all_reduce(s1, op=dist.ReduceOp.SUM); dist.all_reduce(s2, op=dist.ReduceOp.SUM)
        dist.all_reduce(mn, op=dist.ReduceOp.MIN); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    mean = s1 / n_members
    return dict(mean=mean, min=mn, max=mx, var=(s2 / n_members - mean * mean).clamp_min(0.0))


# ------------------------------------------------------------------------------------------------ results
@dataclass
class EnsembleResult:
    t: np.ndarray                       # [T]
    y: Optional[np.ndarray]             # [M (or local M), T, n_save], device layout of the saved elements
    save_idx: np.ndarray                # flat device-layout element index (p * S + s) of each saved column
    status: np.ndarray
    nfev: np.ndarray
    n_accepted: np.ndarray
    n_rejected: np.ndarray
    rtd: Optional[np.ndarray] = None    # [M, S, T, 3]
    moments: Optional[np.ndarray] = None        # [M, S, 3]  (m0, mean, variance)
    moment_stats: Optional[Dict[str, np.ndarray]] = None
    timing: Dict[str, float] = field(default_factory=dict)
    local_range: Tuple[int, int] = (0, 0)
    member_ids: Optional[np.ndarray] = None     # global member id of every row of y / status / ... (identity after a gather)


class EnsembleSolver:
    """Export once (host), upload once per rank, solve any number of parameter batches."""

    def __init__(self, model: str, model_network, config_parser, cmesh, host: Optional[HostSystem] = None):
        t0 = time.perf_counter()
        self.cp = config_parser
        self.model = model.lower()
        self.host = host if host is not None else export_system(self.model, model_network, config_parser, cmesh)
        self.cmesh = cmesh
        self.setup_s = time.perf_counter() - t0
        self.dev = None
        self.t_span = config_parser.get_list(['SIMULATION', 't_span'], float)
        self.rtol = config_parser.get_item(['SIMULATION', 'rtol'], float)
        self.atol = config_parser.get_item(['SIMULATION', 'atol'], float)
        self.first_step = config_parser.get_item(['SIMULATION', 'first_timestep'], float)
        self.solver = config_parser.get_item(['SIMULATION', 'solver'], str).upper()
        self.t_eval = generate_t_eval(config_parser)
        if self.t_eval is None:
            raise ValueError("ensembles need explicit output times (t_eval = 'dt, linear' | 'log, n' | list)")

    # -- what to record
    def outlet_save_idx(self) -> np.ndarray:
        """All species at the outlet node of every model linked to the RTD outlet boundary."""
        from .postprocessing.analysis import outlet_selection
        out_id = self.cmesh.grouped_bcs.id(self.cp.get_item(['POST-PROCESSING', 'outlet_bc_name'], str))
        idx, _, _ = outlet_selection(self.host.outlet_map.get(out_id, []), self.host.flows, self.host.P, self.host.S)
        return idx

    def upload(self):
        from .device import DeviceSystem
        if self.dev is None:
            self.dev = DeviceSystem(self.host)
        return self.dev

    def members_per_batch(self, n_save: int, T: int, reserve: float = 0.85) -> int:
        free, _ = torch.cuda.mem_get_info()
        per_member = 11 * self.host.ndof * 8 + 2 * T * max(n_save, 1) * 8 + 4096
        return max(1, int(free * reserve // per_member))

    def solve(self, n_members: int, *, k_scale: Optional[np.ndarray] = None, q_scale: Optional[np.ndarray] = None,
              bc_scale: Optional[np.ndarray] = None, ic_scale: Optional[np.ndarray] = None,
              y0: Optional[Union[np.ndarray, torch.Tensor]] = None, save: Union[str, np.ndarray] = 'outlets',
              rtd: bool = False, batch: Optional[int] = None, gather: Union[bool, str] = True, to_host: bool = True,
              balance: bool = True) -> EnsembleResult:
        """Solve ``n_members`` systems that share the network but differ in the given per-member scales.

        All array arguments are HOST arrays indexed by the global member id (each rank reads its own members).
        ``y0`` ([M, ndof], reference layout per member) overrides ``ic_scale * c0``.
        ``gather``: True / 'all' = every rank receives all members (all_gather); 'root' = rank 0 only (the other ranks keep their
        own members: at 8 GPUs every rank downloading all 4096 members cost a third of the end-to-end time); False = no collective.
        ``balance``: deal the members to ranks (and to batches inside a rank) by a stiffness estimate instead of in contiguous
        blocks; gathered results are returned in member order either way."""
        from .device import MemberParams
        from .postprocessing.analysis import outlet_selection, rtd_on_device
        tm: Dict[str, float] = {}
        t_start = time.perf_counter()
        sysd = self.upload()
        device = sysd.dev
        host = self.host
        rank, ws = world()
        lo, hi = shard_bounds(n_members, rank, ws)
        M_loc = hi - lo
        perm = balanced_order(stiffness_estimate(host, n_members, k_scale, q_scale), ws) if balance else np.arange(n_members, dtype=np.int64)
        ids = perm[lo:hi]                                   # global ids of this rank's members, stiffest first
        if isinstance(save, str):
            save_idx = None if save == 'all' else self.outlet_save_idx()
        else:
            save_idx = np.asarray(save, dtype=np.int32)
        n_save = host.ndof if save_idx is None else len(save_idx)
        save_dev = None if save_idx is None else torch.as_tensor(save_idx, dtype=torch.int32, device=device)
        te = np.asarray(self.t_eval, dtype=np.float64)
        T = len(te)
        batch = batch or self.members_per_batch(n_save, T)
        c0_dev = torch.as_tensor(host.to_device_layout(host.c0.reshape(-1)), device=device)
        # PFR: the inlet-node part of c0 is the folded boundary value (scales with bc_scale, not with ic_scale)
        c0_bc_dev = None
        if host.c0_bc is not None and np.any(host.c0_bc):
            c0_bc_dev = torch.as_tensor(host.to_device_layout(host.c0_bc.reshape(-1)), device=device)

        def dev_slice(a, b_lo, b_hi, cols=None):
            if a is None:
                return None
            a = np.asarray(a, dtype=np.float64)
            return torch.as_tensor(np.ascontiguousarray(a[ids[b_lo - lo:b_hi - lo]]), device=device)

        y_parts, status, nfev, nacc, nrej, rtd_parts, mom_parts = [], [], [], [], [], [], []
        if self.solver != 'B200-RK45' and self.solver != 'B200-BDF':
            raise ValueError(f'solver {self.solver!r} is not a device solver')
        for b_lo in range(lo, hi, batch):
            b_hi = min(hi, b_lo + batch)
            mb = b_hi - b_lo
            mem = MemberParams(mb, dev_slice(k_scale, b_lo, b_hi), dev_slice(q_scale, b_lo, b_hi), dev_slice(bc_scale, b_lo, b_hi))
            if y0 is not None:
                sel = ids[b_lo - lo:b_hi - lo]
                if isinstance(y0, torch.Tensor):
                    yb = y0[torch.as_tensor(sel, device=y0.device)].to(device)
                else:
                    yb = torch.as_tensor(np.ascontiguousarray(np.asarray(y0)[sel]), device=device)
                # reference layout [S, n_points] -> device layout [n_points, S]
                yb = yb.reshape(mb, host.S, host.n_points).transpose(1, 2).contiguous().reshape(mb, -1)
            else:
                ones = torch.ones(mb, dtype=torch.float64, device=device)
                ic_b = dev_slice(ic_scale, b_lo, b_hi) if ic_scale is not None else ones
                if c0_bc_dev is None:
                    yb = (ic_b[:, None] * c0_dev[None, :]).contiguous()
                else:
                    bc_b = dev_slice(bc_scale, b_lo, b_hi) if bc_scale is not None else ones
                    yb = (ic_b[:, None] * (c0_dev - c0_bc_dev)[None, :] + bc_b[:, None] * c0_bc_dev[None, :]).contiguous()
            if self.solver == 'B200-RK45':
                res = sysd.rk45(yb, self.t_span, rtol=self.rtol, atol=self.atol, first_step=self.first_step, t_eval=te,
                                save_idx=save_dev, members=mem)
            else:
                from .bdf import device_tolerances, solve_bdf
                rtol_d, atol_d = device_tolerances(self.rtol, self.atol)
                res = solve_bdf(sysd, yb, self.t_span, rtol=rtol_d, atol=atol_d, first_step=self.first_step,
                                t_eval=te, save_idx=save_dev, members=mem)
            y_parts.append(res.y)
            status.append(res.status); nfev.append(res.nfev); nacc.append(res.n_accepted); nrej.append(res.n_rejected)
            if rtd:
                in_id = self.cmesh.grouped_bcs.id(self.cp.get_item(['POST-PROCESSING', 'inlet_bc_name'], str))
                out_id = self.cmesh.grouped_bcs.id(self.cp.get_item(['POST-PROCESSING', 'outlet_bc_name'], str))
                q_in = float(sum(host.flows[c] for _, c in host.inlet_map[in_id]))
                idx, w, sp = outlet_selection(host.outlet_map[out_id], host.flows, host.P, host.S)
                if save_idx is None:
                    pos = idx
                else:
                    lookup = {int(e): i for i, e in enumerate(save_idx)}
                    pos = np.array([lookup[int(e)] for e in idx], dtype=np.int32)
                r, mo = rtd_on_device(res.y, torch.as_tensor(te, device=device), torch.as_tensor(pos, dtype=torch.int32, device=device),
                                      torch.as_tensor(w, device=device), torch.as_tensor(sp, device=device), host.S, q_in)
                rtd_parts.append(r); mom_parts.append(mo)
        tm['solve_s'] = time.perf_counter() - t_start

        cat = lambda parts, shape: torch.cat(parts, 0) if parts else torch.zeros(shape, dtype=torch.float64, device=device)
        y_loc = cat(y_parts, (0, T, n_save))
        st_loc = torch.as_tensor(np.concatenate(status) if status else np.zeros(0, np.int32), device=device)
        cnt_loc = torch.as_tensor(np.stack([np.concatenate(nfev), np.concatenate(nacc), np.concatenate(nrej)], 1)
                                  if nfev else np.zeros((0, 3), np.int64), device=device)
        rtd_loc = cat(rtd_parts, (0, host.S, T, 3)) if rtd else None
        mom_loc = cat(mom_parts, (0, host.S, 3)) if rtd else None
        stats = None
        member_ids = ids
        if rtd:
            torch.cuda.synchronize(device); t_r = time.perf_counter()
            stats = {k: v.cpu().numpy() for k, v in reduce_statistics(mom_loc, n_members).items()}
            tm['reduce_s'] = time.perf_counter() - t_r            # all_reduce(SUM, SUM, MIN, MAX) of the RTD moments + download
        root = 0 if gather == 'root' else None
        if gather and (root is None or ws > 1):
            torch.cuda.synchronize(device); t_g = time.perf_counter()
            inv = torch.as_tensor(np.argsort(perm), device=device)        # gathered rows arrive in dealing order
            have_all = root is None or rank == root
            def collect(x):
                if not (ws > 1 or balance):
                    return x
                g = gather_members(x, n_members, root)
                return g[inv] if g is not None else x               # non-root ranks of a rooted gather keep their own rows
            y_loc, st_loc, cnt_loc = collect(y_loc), collect(st_loc), collect(cnt_loc)
            if rtd:
                rtd_loc, mom_loc = collect(rtd_loc), collect(mom_loc)
            torch.cuda.synchronize(device)
            tm['gather_s'] = time.perf_counter() - t_g              # (all_)gather of the recorded observables (NCCL on GPUs)
            if have_all and (ws > 1 or balance):
                member_ids = np.arange(n_members, dtype=np.int64)
        elif gather and balance:                                    # rooted gather on one rank: only the permutation is undone
            inv = torch.as_tensor(np.argsort(perm), device=device)
            y_loc, st_loc, cnt_loc = y_loc[inv], st_loc[inv], cnt_loc[inv]
            if rtd:
                rtd_loc, mom_loc = rtd_loc[inv], mom_loc[inv]
            member_ids = np.arange(n_members, dtype=np.int64)
        tm['total_s'] = time.perf_counter() - t_start
        cnt = cnt_loc.cpu().numpy()
        to_np = to_host_pinned if to_host else (lambda x: x)
        return EnsembleResult(te, to_np(y_loc), np.arange(host.ndof, dtype=np.int32) if save_idx is None else save_idx,
                              st_loc.cpu().numpy(), cnt[:, 0], cnt[:, 1], cnt[:, 2], to_np(rtd_loc), to_np(mom_loc), stats, tm,
                              (lo, hi), member_ids)


def solve_ensemble(model: str, model_network, config_parser, cmesh, n_members: int, **kw) -> EnsembleResult:
    """One-call form: export, upload, solve."""
    return EnsembleSolver(model, model_network, config_parser, cmesh).solve(n_members, **kw)
