"""
Host wrapper of the batched BDF integrator (occm_bdf_* in include/occm_b200.h).

Replaces ``solve_ivp(..., method='BDF'/'LSODA')`` on stiff networks (the reference's default solver is LSODA,
config_functions/expanded_config_parser.py:55, which spends >97 % of its RHS calls on dense finite-difference Jacobians,
SURVEY.md §3.5).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import os

import numpy as np

import torch

from . import _lib
from .device import DeviceSystem, MemberParams, _ptr, _stream


@dataclass
class BdfResult:
    t: object
    y: object
    status: np.ndarray
    n_accepted: np.ndarray
    n_rejected: np.ndarray
    nfev: np.ndarray
    n_gmres: np.ndarray
    n_newton_fail: np.ndarray
    y_final: Optional[torch.Tensor] = None
    n_rows: Optional[np.ndarray] = None
    n_precond_apply: Optional[np.ndarray] = None
    n_precond_factor: Optional[np.ndarray] = None


def device_tolerances(rtol: float, atol: float):
    """Tolerance-mapping policy of the ``B200-BDF`` solver name: the device integrates at ``f * (rtol, atol)``, f = 0.01
    (``OCCM_B200_BDF_TOL_FACTOR`` overrides).

    Why: measured against tight solutions of the golden cases (tests/golden/tight_*.npz, LSODA at rtol 1e-13), scipy's BDF at
    rtol 1e-6 / atol 1e-10 is up to 22 acceptance bands (1e-10 + 1e-6 |c|) away, at 1e-7 up to 3.4, at 1e-8 inside the band on
    every case; the reference's default LSODA at 1e-6 is within 0.4 ... 4 bands (19 on the point-mass case).  A BDF that mirrors
    scipy's controller at the user's own numbers would therefore miss the band where the reference meets it.  Two decades
    cost a variable-order method (error ~ h^(q+1), q <= 5) about 2x the steps.  `solve_bdf` itself applies no mapping."""
    f = float(os.environ.get('OCCM_B200_BDF_TOL_FACTOR', '0.01'))
    return rtol * f, atol * f


def transport_diagonal(system: DeviceSystem) -> torch.Tensor:
    """Diagonal of the transport operator per model: CSTR Q_div_v[j, j] (cstr_system.py:133-135), PFR -Q_pfr/dV."""
    host = system.host
    if host.model == 'pfr':
        d = -np.asarray(host.a, dtype=np.float64)
    else:
        rows = np.repeat(np.arange(host.n_models), np.diff(host.row_ptr))
        on_diag = host.row_src == rows
        d = np.bincount(rows[on_diag], weights=host.row_w[on_diag], minlength=host.n_models)
    return torch.as_tensor(np.ascontiguousarray(d), device=system.dev)


class BdfRun:
    """One batched BDF integration: owns the workspace / output tensors and the C handle (context manager)."""

    def __init__(self, system: DeviceSystem, y0: torch.Tensor, t_span: Sequence[float], *, rtol: float, atol: float, first_step: float,
                 t_eval: Optional[Sequence[float]] = None, save_idx: Optional[torch.Tensor] = None,
                 members: Optional[MemberParams] = None, max_steps: int = 0, gmres_restart: int = 30, lin_tol_factor: float = 0.0,
                 all_capacity: int = 4096, precond_mode: Optional[int] = None):
        auto_mode = precond_mode is None and 'OCCM_B200_BDF_PRECOND' not in os.environ
        if precond_mode is None:
            precond_mode = int(os.environ.get('OCCM_B200_BDF_PRECOND', '0'))
        assert y0.is_cuda and y0.dtype == torch.float64 and y0.is_contiguous()
        self.sys = system
        lib = self.lib = system.lib
        M = self.M = 1 if y0.dim() == 1 else y0.shape[0]
        assert y0.numel() == M * system.ndof
        self.members = members if members is not None else MemberParams(M)
        self.n_save = system.ndof if save_idx is None else int(save_idx.numel())
        self.all_mode = t_eval is None
        if self.all_mode:
            T, self.te, self.te_np = int(all_capacity), None, None
        else:
            te_np = np.ascontiguousarray(np.asarray(t_eval, dtype=np.float64))
            if np.any(te_np < min(t_span)) or np.any(te_np > max(t_span)):
                raise ValueError("Values in `t_eval` are not within `t_span`.")
            if np.any(np.diff(te_np) <= 0):
                raise ValueError("Values in `t_eval` are not properly sorted.")
            T = te_np.size
            self.te_np = te_np
            self.te = torch.as_tensor(te_np, device=system.dev)
        self.out = torch.empty((M, T, self.n_save), dtype=torch.float64, device=system.dev)
        self.out_t = torch.empty((M, T), dtype=torch.float64, device=system.dev)
        self.tdiag = transport_diagonal(system)
        ws_bytes = lib.occm_bdf_workspace_bytes(system.handle, M, gmres_restart, int(precond_mode))
        if auto_mode and precond_mode == 0:
            # the stored block inverses are optional: drop them when they would not fit next to the Krylov vectors
            free_bytes = torch.cuda.mem_get_info(system.dev)[0]
            if ws_bytes > 0.9 * free_bytes:
                precond_mode = 1
                ws_bytes = lib.occm_bdf_workspace_bytes(system.handle, M, gmres_restart, 1)
        self.precond_mode = int(precond_mode)
        self.ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=system.dev)
        ws_ptr = (self.ws.data_ptr() + 255) & ~255
        self.save_idx = save_idx
        prob = _lib.BdfProblem(float(t_span[0]), float(t_span[1]), float(first_step), float(rtol), float(atol), y0.data_ptr(),
                               _ptr(self.te), T, _ptr(save_idx), self.n_save, self.out.data_ptr(), self.out_t.data_ptr(), int(max_steps),
                               self.tdiag.data_ptr(), int(gmres_restart), float(lin_tol_factor), int(precond_mode))
        self._ms = self.members.c_struct()
        self.handle = C.c_void_p()
        _lib.check(lib.occm_bdf_init(C.byref(self.handle), system.handle, C.byref(self._ms), C.byref(prob), ws_ptr, ws_bytes, _stream()))
        i64 = lambda: np.zeros(M, dtype=np.int64)
        self.status = np.zeros(M, dtype=np.int32); self.order = np.zeros(M, dtype=np.int32); self.n_rows = np.zeros(M, dtype=np.int32)
        self.n_acc, self.n_rej, self.nfev, self.n_gmres, self.n_newton_fail = i64(), i64(), i64(), i64(), i64()
        self.n_precond_apply, self.n_precond_factor = i64(), i64()
        self.t_now = np.zeros(M)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def close(self):
        if self.handle:
            self.lib.occm_bdf_free(self.handle)
            self.handle = C.c_void_p()
        self.ws = None

    def run(self, max_attempts: int = 0) -> int:
        """Advance until every member is finished / failed / paused, or for at most ``max_attempts`` step attempts."""
        running = C.c_int32(0)
        _lib.check(self.lib.occm_bdf_run(self.handle, int(max_attempts), C.byref(running), _stream()))
        return running.value

    def stats(self):
        _lib.check(self.lib.occm_bdf_stats(self.handle, self.status.ctypes.data, self.n_acc.ctypes.data, self.n_rej.ctypes.data,
                                           self.nfev.ctypes.data, self.n_gmres.ctypes.data, self.n_newton_fail.ctypes.data,
                                           self.t_now.ctypes.data, self.n_rows.ctypes.data, self.order.ctypes.data, _stream()))
        _lib.check(self.lib.occm_bdf_precond_stats(self.handle, self.n_precond_apply.ctypes.data, self.n_precond_factor.ctypes.data, _stream()))

    def resume(self):
        _lib.check(self.lib.occm_bdf_resume(self.handle, _stream()))

    def time_kernel(self, which: int, reps: int = 10) -> float:
        """Mean duration (ms, CUDA events) of one kernel of the Newton / GMRES loop on the handle's current buffers:
        0 = fused RHS at the predictor, 1 = matrix-free J*v (RHS of y + eps z), 2 = preconditioner application."""
        ms = C.c_float(0.0)
        _lib.check(self.lib.occm_bdf_time_kernel(self.handle, int(which), int(reps), C.byref(ms), _stream()))
        return float(ms.value)

    def final_state(self) -> torch.Tensor:
        y = torch.empty((self.M, self.sys.ndof), dtype=torch.float64, device=self.sys.dev)
        _lib.check(self.lib.occm_bdf_get_state(self.handle, y.data_ptr(), _stream()))
        torch.cuda.current_stream().synchronize()
        return y


def solve_bdf(system: DeviceSystem, y0: torch.Tensor, t_span: Sequence[float], *, return_final: bool = False, **kw) -> BdfResult:
    """``precond_mode``: 0 keeps the block inverses of the preconditioner in HBM (ndof * S doubles per member, rebuilt by
    CVODE's setup policy), 1 rebuilds and solves the blocks on every application; None = ``OCCM_B200_BDF_PRECOND`` or 0."""
    with BdfRun(system, y0, t_span, **kw) as run:
        M = run.M
        ts_all = [[] for _ in range(M)]; ys_all = [[] for _ in range(M)]
        while True:
            run.run()
            run.stats()
            if not run.all_mode:
                break
            for m in range(M):
                r = int(run.n_rows[m])
                if r:
                    ts_all[m].append(run.out_t[m, :r].cpu().numpy()); ys_all[m].append(run.out[m, :r].clone())
            if not np.any(run.status == 2):
                break
            run.resume()
        y_final = run.final_state() if return_final else None
        st = (run.status.copy(), run.n_acc.copy(), run.n_rej.copy(), run.nfev.copy(), run.n_gmres.copy(), run.n_newton_fail.copy())
        pa, pf = run.n_precond_apply.copy(), run.n_precond_factor.copy()
        if run.all_mode:
            t_res = [np.concatenate(x) if x else np.zeros(0) for x in ts_all]
            y_res = [torch.cat(x) if x else run.out[m, :0] for m, x in enumerate(ys_all)]
            return BdfResult(t_res, y_res, *st, y_final, None, pa, pf)
        return BdfResult(run.te_np, run.out, *st, y_final, run.n_rows.copy(), pa, pf)
