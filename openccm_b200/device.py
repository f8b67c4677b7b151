"""
Device-resident system: uploads a ``HostSystem`` once (torch CUDA tensors own the memory) and drives the
C ABI of liboccm_b200.so.  PyTorch is plumbing here (allocation, streams, distributed); all arithmetic on the
solver path runs in the hand-written CUDA kernels.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from .system_solvers.export import HostSystem

STATUS_MESSAGES = {
    1: "The solver successfully reached the end of the integration interval.",
    2: "Paused: output buffer full.",
    0: "Still running.",
    -1: "Required step size is less than spacing between numbers.",
    -2: "Non-finite error estimate.",
    -3: "Step budget exhausted.",
}


def _require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.OccmError('no CUDA device: the B200 solver path has no CPU fallback')
    return torch.device('cuda', torch.cuda.current_device())


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def ell_from_csr(row_ptr: np.ndarray, row_src: np.ndarray, row_w: np.ndarray, pulse_ptr: Optional[np.ndarray] = None,
                 coverage: float = 0.98, max_width: int = 4):
    """ELL copy of the CSR rows (column-major [K][N]) plus the per-row flag of the rows it cannot represent.

    K = the smallest width that holds ``coverage`` of the rows (capped at ``max_width``).  Rows that are longer,
    contain a boundary entry (src < 0) or carry a pulse source are flagged: the kernel evaluates them from the CSR
    arrays.  Short rows are padded with (own index, weight 0) so the fast path needs no branch."""
    row_ptr = np.asarray(row_ptr, dtype=np.int64)
    n = len(row_ptr) - 1
    deg = np.diff(row_ptr)
    flags = np.zeros(n, dtype=np.uint8)
    has_bc = np.zeros(n, dtype=bool)
    if len(row_src):
        rows = np.repeat(np.arange(n), deg)
        has_bc[np.unique(rows[np.asarray(row_src) < 0])] = True
    special = has_bc.copy()
    if pulse_ptr is not None:
        special |= np.diff(np.asarray(pulse_ptr, dtype=np.int64)) > 0
    plain_deg = deg[~special] if (~special).any() else deg          # rows the ELL path can serve at all
    k = int(min(max_width, max(1, np.quantile(plain_deg, coverage, method='higher')))) if n else 1
    if n and plain_deg.max() <= max_width and plain_deg.max() - k <= 1:
        k = int(max(1, plain_deg.max()))
    flags[(deg > k) | special] = 1
    src = np.tile(np.arange(n, dtype=np.int64), (k, 1))
    w = np.zeros((k, n), dtype=np.float64)
    good = flags == 0
    for col in range(k):
        take = good & (deg > col)
        idx = row_ptr[:-1][take] + col
        src[col, take] = row_src[idx]
        w[col, take] = row_w[idx]
    # columns ordered nearest source first: the diagonal and the numbering's neighbours land in the leading columns, whose rows the
    # pipelined kernels find in the staged chunk; far (out-of-chunk) sources share the last columns, which the specialised
    # kernels gather warp-cooperatively (csrc/occm_row.cuh)
    if k > 1 and n:
        order = np.argsort(np.abs(src - np.arange(n, dtype=np.int64)[None, :]), axis=0, kind='stable')
        src = np.take_along_axis(src, order, axis=0)
        w = np.take_along_axis(w, order, axis=0)
    return k, src.astype(np.int32), w, flags


@dataclass
class MemberParams:
    """Per-member ensemble parameters (device tensors or None)."""
    n_members: int
    k_scale: Optional[torch.Tensor] = None   # [M, n_rxn]
    q_scale: Optional[torch.Tensor] = None   # [M]
    bc_scale: Optional[torch.Tensor] = None  # [M]

    def c_struct(self) -> _lib.Members:
        return _lib.Members(self.n_members, _ptr(self.k_scale), _ptr(self.q_scale), _ptr(self.bc_scale))


@dataclass
class Rk45Result:
    t: np.ndarray              # [T] (t_eval mode) or list of per-member arrays ("all" mode)
    y: torch.Tensor            # device [M, T, n_save]  (t_eval mode)
    status: np.ndarray
    n_accepted: np.ndarray
    n_rejected: np.ndarray
    nfev: np.ndarray
    y_final: Optional[torch.Tensor] = None
    n_rows: Optional[np.ndarray] = None      # t_eval mode: output rows actually written per member (< T after a failure)


class DeviceSystem:
    def __init__(self, host: HostSystem, specialize: Optional[str] = None):
        """specialize: 'auto' | 'on' | 'off' — mechanism-specialised kernel image (openccm_b200/specialize.py; default: the
        OCCM_B200_SPECIALIZE environment variable, else 'auto')."""
        self.lib = _lib.load()
        self.dev = _require_cuda()
        self.host = host
        i32 = lambda a: None if a is None else torch.as_tensor(np.ascontiguousarray(a, dtype=np.int32), device=self.dev)
        f64 = lambda a: None if a is None else torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device=self.dev)
        self.row_ptr, self.row_src, self.row_w = i32(host.row_ptr), i32(host.row_src), f64(host.row_w)
        if self.row_src.numel() == 0:       # keep valid pointers for empty operators
            self.row_src = torch.zeros(1, dtype=torch.int32, device=self.dev)
            self.row_w = torch.zeros(1, dtype=torch.float64, device=self.dev)
        self.a = f64(host.a)
        self.pulse_ptr, self.pulse_fn, self.pulse_w = i32(host.pulse_ptr), i32(host.pulse_fn), f64(host.pulse_w)
        blob = host.tables.pack()
        self.tables_blob = blob
        self.tables = torch.as_tensor(blob, device=self.dev)
        self.fields = f64(host.fields)
        self.n_rxn = host.tables.n_rxn
        self.ell_k, self.ell_src, self.ell_w, self.ell_packed = 0, None, None, None
        if host.model != 'pfr':
            k, es, ew, flags = ell_from_csr(host.row_ptr, host.row_src, host.row_w, host.pulse_ptr)
            self.ell_k, self.ell_src, self.ell_w = k, i32(es), f64(ew)
            packed = np.zeros((k, host.n_models), dtype=np.dtype([('src', '<i4'), ('flag', '<i4'), ('w', '<f8')]))
            packed['src'] = es; packed['w'] = ew; packed['flag'][0] = flags
            self.ell_packed = torch.as_tensor(packed.view(np.uint8).reshape(-1), device=self.dev)
        else:
            flags = np.zeros(host.n_points, dtype=np.uint8)
            if host.pulse_ptr is not None:       # every node of a PFR with a pulse source takes the generic path
                flags[np.repeat(np.diff(host.pulse_ptr) > 0, host.P)] = 1
            flags[::host.P] = 1                  # inlet nodes: differentiated algebraic coupling to other PFRs (pfr_system.py:313-314)
            # one-entry row per point for the fast kernels: {upwind point, flag, Q/dV}
            pts = np.arange(host.n_points, dtype=np.int64)
            packed = np.zeros((1, host.n_points), dtype=np.dtype([('src', '<i4'), ('flag', '<i4'), ('w', '<f8')]))
            packed['src'][0] = np.where(pts % host.P == 0, pts, pts - 1)
            packed['w'][0] = np.repeat(np.asarray(host.a, dtype=np.float64), host.P)
            packed['flag'][0] = flags
            self.ell_k = 1
            self.ell_packed = torch.as_tensor(packed.view(np.uint8).reshape(-1), device=self.dev)
        self.point_flags = torch.as_tensor(flags, device=self.dev)
        flagged = np.nonzero(flags)[0].astype(np.int32)
        self.flagged_points = torch.as_tensor(flagged if len(flagged) else np.zeros(1, np.int32), device=self.dev)
        desc = _lib.SystemDesc(1 if host.model == 'pfr' else 0, host.S, host.n_models, host.P,
                               _ptr(self.row_ptr), _ptr(self.row_src), _ptr(self.row_w), _ptr(self.a),
                               _ptr(self.pulse_ptr), _ptr(self.pulse_fn), _ptr(self.pulse_w),
                               _ptr(self.tables), blob.nbytes, _ptr(self.fields),
                               0 if host.fields is None else host.fields.shape[0],
                               self.ell_k, _ptr(self.ell_src), _ptr(self.ell_w), _ptr(self.point_flags), _ptr(self.ell_packed),
                               _ptr(self.flagged_points), int(len(flagged)))
        h = C.c_void_p()
        _lib.check(self.lib.occm_system_create(C.byref(h), C.byref(desc)))
        self.handle = h
        self.ndof = host.ndof
        self.spec_image = None
        from .specialize import specialize as _specialize
        self.specialized = _specialize(self, specialize)

    def __del__(self):
        try:
            if getattr(self, 'handle', None):
                self.lib.occm_system_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ------------------------------------------------------------------ fused RHS
    def rhs(self, y: torch.Tensor, t, members: Optional[MemberParams] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """dy[m] = ddt(t[m], y[m]); y: device [M, ndof] in device layout (point-major, species innermost)."""
        assert y.is_cuda and y.dtype == torch.float64 and y.is_contiguous()
        M = 1 if y.dim() == 1 else y.shape[0]
        assert y.numel() == M * self.ndof
        mem = members if members is not None else MemberParams(M)
        assert mem.n_members == M
        if out is None:
            out = torch.empty_like(y)
        ms = mem.c_struct()
        if isinstance(t, torch.Tensor):
            assert t.is_cuda and t.dtype == torch.float64 and t.numel() == M
            _lib.check(self.lib.occm_rhs(self.handle, C.byref(ms), t.data_ptr(), 0.0, y.data_ptr(), out.data_ptr(), _stream()))
        else:
            _lib.check(self.lib.occm_rhs(self.handle, C.byref(ms), None, float(t), y.data_ptr(), out.data_ptr(), _stream()))
        return out

    # ------------------------------------------------------------------ RK45
    def open_rk45(self, y0: torch.Tensor, t_span: Sequence[float], *, rtol: float, atol: float, first_step: float,
                  t_eval: Optional[Sequence[float]] = None, save_idx: Optional[torch.Tensor] = None,
                  members: Optional[MemberParams] = None, max_steps: int = 0, all_capacity: int = 4096) -> "Rk45Run":
        return Rk45Run(self, y0, t_span, rtol=rtol, atol=atol, first_step=first_step, t_eval=t_eval, save_idx=save_idx,
                       members=members, max_steps=max_steps, all_capacity=all_capacity)

    def rk45(self, y0: torch.Tensor, t_span: Sequence[float], *, return_final: bool = False, **kw) -> Rk45Result:
        """Batched scipy-compatible RK45. y0: device [M, ndof] (device layout)."""
        with self.open_rk45(y0, t_span, **kw) as run:
            return run.solve(return_final=return_final)


class Rk45Run:
    """One batched RK45 integration: owns the workspace / output tensors and the C handle."""

    def __init__(self, system: DeviceSystem, y0: torch.Tensor, t_span, *, rtol, atol, first_step, t_eval, save_idx,
                 members, max_steps, all_capacity):
        assert y0.is_cuda and y0.dtype == torch.float64 and y0.is_contiguous()
        self.sys = system
        lib = self.lib = system.lib
        M = self.M = 1 if y0.dim() == 1 else y0.shape[0]
        assert y0.numel() == M * system.ndof
        self.members = members if members is not None else MemberParams(M)
        assert self.members.n_members == M
        self.n_save = system.ndof if save_idx is None else int(save_idx.numel())
        if save_idx is not None:
            assert save_idx.is_cuda and save_idx.dtype == torch.int32
        self.save_idx = save_idx
        ws_bytes = lib.occm_rk45_workspace_bytes(system.handle, M)
        self.ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=system.dev)
        ws_ptr = (self.ws.data_ptr() + 255) & ~255
        self.all_mode = t_eval is None
        if self.all_mode:
            T = int(all_capacity)
            self.te = None; self.te_np = None
        else:
            te_np = np.ascontiguousarray(np.asarray(t_eval, dtype=np.float64))
            if te_np.ndim != 1 or te_np.size == 0:
                raise ValueError("`t_eval` must be 1-dimensional.")
            if np.any(te_np < min(t_span)) or np.any(te_np > max(t_span)):
                raise ValueError("Values in `t_eval` are not within `t_span`.")
            if np.any(np.diff(te_np) <= 0):
                raise ValueError("Values in `t_eval` are not properly sorted.")
            T = te_np.size
            self.te_np = te_np
            self.te = torch.as_tensor(te_np, device=system.dev)
        self.T = T
        self.out = torch.empty((M, T, self.n_save), dtype=torch.float64, device=system.dev)
        self.out_t = torch.empty((M, T), dtype=torch.float64, device=system.dev)
        prob = _lib.Rk45Problem(float(t_span[0]), float(t_span[1]), float(first_step), float(rtol), float(atol),
                                y0.data_ptr(), _ptr(self.te), T, _ptr(save_idx), self.n_save, self.out.data_ptr(),
                                self.out_t.data_ptr(), int(max_steps))
        self._ms = self.members.c_struct()
        self.handle = C.c_void_p()
        _lib.check(lib.occm_rk45_init(C.byref(self.handle), system.handle, C.byref(self._ms), C.byref(prob), ws_ptr,
                                      ws_bytes, _stream()))
        self.status = np.zeros(M, dtype=np.int32); self.n_acc = np.zeros(M, dtype=np.int64)
        self.n_rej = np.zeros(M, dtype=np.int64); self.nfev = np.zeros(M, dtype=np.int64)
        self.t_now = np.zeros(M); self.n_rows = np.zeros(M, dtype=np.int32)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def close(self):
        if self.handle:
            self.lib.occm_rk45_free(self.handle)
            self.handle = C.c_void_p()
        self.ws = None            # the workspace (11 state-sized vectors) goes back to the allocator right away

    def run(self, max_launch_steps: int = 0) -> int:
        running = C.c_int32(0)
        _lib.check(self.lib.occm_rk45_run(self.handle, int(max_launch_steps), C.byref(running), _stream()))
        return running.value

    def stats(self):
        _lib.check(self.lib.occm_rk45_stats(self.handle, self.status.ctypes.data, self.n_acc.ctypes.data,
                                            self.n_rej.ctypes.data, self.nfev.ctypes.data, self.t_now.ctypes.data,
                                            self.n_rows.ctypes.data, _stream()))

    def time_stage(self, stage: int, reps: int = 10) -> float:
        ms = C.c_float(0.0)
        _lib.check(self.lib.occm_rk45_time_stage(self.handle, int(stage), int(reps), C.byref(ms), _stream()))
        return float(ms.value)

    def final_state(self) -> torch.Tensor:
        y = torch.empty((self.M, self.sys.ndof), dtype=torch.float64, device=self.sys.dev)
        _lib.check(self.lib.occm_rk45_get_state(self.handle, y.data_ptr(), _stream()))
        torch.cuda.current_stream().synchronize()
        return y

    def solve(self, return_final: bool = False) -> Rk45Result:
        M = self.M
        ts_all = [[] for _ in range(M)]; ys_all = [[] for _ in range(M)]
        while True:
            self.run()
            self.stats()
            if not self.all_mode:
                break
            for m in range(M):                       # harvest the rows written so far, un-pause and continue
                r = int(self.n_rows[m])
                if r:
                    ts_all[m].append(self.out_t[m, :r].cpu().numpy())
                    ys_all[m].append(self.out[m, :r].clone())
            if not np.any(self.status == 2):
                break
            _lib.check(self.lib.occm_rk45_resume(self.handle, _stream()))
        y_final = self.final_state() if return_final else None
        st = (self.status.copy(), self.n_acc.copy(), self.n_rej.copy(), self.nfev.copy())
        if self.all_mode:
            t_res = [np.concatenate(x) if x else np.zeros(0) for x in ts_all]
            y_res = [torch.cat(x) if x else self.out[m, :0] for m, x in enumerate(ys_all)]
            return Rk45Result(t_res, y_res, *st, y_final)
        return Rk45Result(self.te_np, self.out, *st, y_final, self.n_rows.copy())
