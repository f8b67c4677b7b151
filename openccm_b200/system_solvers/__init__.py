"""
Device replacement for ``openccm.system_solvers`` (/root/reference/openccm/system_solvers/__init__.py:38-81).

``solve_system(model, model_network, config_parser, cmesh)`` has the reference's signature and return value
``(c[S, n_points, T], t[T], inlet_map, outlet_map)``; the time integration runs on the GPU through
liboccm_b200.so.  Selection is by ``[SIMULATION] solver``:

    B200-RK45   adaptive Dormand-Prince, step-for-step mirror of scipy's RK45
    B200-LSODA  automatic choice: an RK45 probe, then BDF from the probe's state if the step size is stability limited
    B200-BDF    variable-order BDF with Newton + preconditioned GMRES (stiff networks); integrates at 0.01 x the
                user's (rtol, atol) so that it meets the 1e-10 + 1e-6 |c| band wherever the reference's LSODA does
                (openccm_b200.bdf.device_tolerances)

Any other solver name raises here (this package has no CPU path); ``openccm_b200.dropin.install()`` forwards
those names to the reference's own scipy path when the reference is importable.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np

from .export import HostSystem, export_system
from .helper_functions import generate_t_eval

DEVICE_SOLVERS = ('B200-RK45', 'B200-BDF', 'B200-LSODA')

# B200-LSODA: automatic method selection in the spirit of LSODA (the reference's default solver, expanded_config_parser.py:55;
# scipy/integrate/_ivp/lsoda.py:120-197 wraps ODEPACK, which starts non-stiff and switches when the problem turns stiff).
# Here: a probe of explicit RK45 steps; when it has not finished and the step size it settled at projects to more than
# LSODA_SWITCH_STEPS further steps — on these networks that means the step is limited by the fastest compartment
# (h <~ 3.3 tau_min), not by accuracy — the integration continues from the probe's state with the device BDF.
# One-way (non-stiff -> stiff), decided once.
LSODA_PROBE_STEPS = 400
LSODA_SWITCH_STEPS = 4000


def is_device_solver(name: str) -> bool:
    return name.upper() in DEVICE_SOLVERS


def solve_system(model: str, model_network, config_parser, cmesh, *, return_info: bool = False, fields=None, max_steps: int = 0):
    """Drop-in for openccm.system_solvers.solve_system (system_solvers/__init__.py:38-81)."""
    import torch
    from ..device import DeviceSystem, STATUS_MESSAGES

    print("Solving simulation")
    cp = config_parser
    t_span = cp.get_list(['SIMULATION', 't_span'], float)
    first_timestep = cp.get_item(['SIMULATION', 'first_timestep'], float)
    atol = cp.get_item(['SIMULATION', 'atol'], float)
    rtol = cp.get_item(['SIMULATION', 'rtol'], float)
    solver = cp.get_item(['SIMULATION', 'solver'], str).upper()
    assert first_timestep > 0                                  # cstr_system.py:95 / pfr_system.py:95
    if solver not in DEVICE_SOLVERS:
        raise ValueError(f"solver = {solver!r} is not a device solver; use one of {DEVICE_SOLVERS} "
                         f"(this package has no CPU integration path)")
    t_eval = generate_t_eval(cp)
    model = model.lower()
    host = export_system(model, model_network, cp, cmesh, fields=fields)
    dev = DeviceSystem(host)
    y0 = torch.as_tensor(host.to_device_layout(host.c0.reshape(-1)), device=dev.dev).reshape(1, -1)

    switched = None
    if solver == 'B200-RK45':
        res = dev.rk45(y0, t_span, rtol=rtol, atol=atol, first_step=first_timestep, t_eval=t_eval, max_steps=max_steps)
    elif solver == 'B200-LSODA' and t_eval is not None:
        res, switched = _solve_auto(dev, y0, t_span, rtol, atol, first_timestep, t_eval)
    else:
        from ..bdf import device_tolerances, solve_bdf
        rtol_d, atol_d = device_tolerances(rtol, atol)         # tolerance-mapping policy of B200-BDF (see its docstring)
        res = solve_bdf(dev, y0, t_span, rtol=rtol_d, atol=atol_d, first_step=first_timestep, t_eval=t_eval, max_steps=max_steps)
    status = int(res.status[0])
    message = STATUS_MESSAGES.get(status, str(status))
    if model == 'pfr':
        assert status == 1, message                            # pfr_system.py:192 (the CSTR path does not check)
    if t_eval is None:
        ts = np.asarray(res.t[0]); y = res.y[0].cpu().numpy()
    else:
        # like solve_ivp, a run that stops early (CSTR path: no success check, cstr_system.py:151-154) returns the rows reached
        n_rows = len(res.t) if status == 1 or res.n_rows is None else int(res.n_rows[0])
        ts = np.asarray(res.t)[:n_rows]; y = res.y[0, :n_rows].cpu().numpy()
    # [T, n_points, S] -> [S, n_points, T]   (cstr_system.py:152-154)
    c = np.ascontiguousarray(y.reshape(len(ts), host.n_points, host.S).transpose(2, 1, 0))

    if model == 'pfr' and host.coupling is not None and len(host.coupling) and host.coupling[0, 0] > 0:
        _dae_residual_warning(c, host, atol)                   # pfr_system.py:198-212
    print("Done solving simulation")
    out = (c, ts, dict(host.inlet_map), dict(host.outlet_map))
    if return_info:
        return out + (dict(status=status, message=message, nfev=int(res.nfev[0]), n_accepted=int(res.n_accepted[0]),
                           n_rejected=int(res.n_rejected[0]), switched_at=switched, specialized=bool(dev.specialized)),)
    return out


def _solve_auto(dev, y0, t_span, rtol, atol, first_step, t_eval):
    """B200-LSODA: RK45 probe, then (if the projected number of steps says the step is stability limited) BDF from the probe's
    state.  Returns (result with the rows of both phases, switch time or None)."""
    import torch
    from ..bdf import device_tolerances, solve_bdf
    from ..device import Rk45Result
    te = np.asarray(t_eval, dtype=np.float64)
    with dev.open_rk45(y0, t_span, rtol=rtol, atol=atol, first_step=first_step, t_eval=te, save_idx=None, members=None,
                       max_steps=LSODA_PROBE_STEPS, all_capacity=0) as run:
        run.run()
        run.stats()
        probe = Rk45Result(te, run.out, run.status.copy(), run.n_acc.copy(), run.n_rej.copy(), run.nfev.copy(), None, run.n_rows.copy())
        if int(run.status[0]) != -3:                           # finished (or failed) inside the probe: a plain RK45 solve
            return probe, None
        t_now, rows, n_acc = float(run.t_now[0]), int(run.n_rows[0]), int(run.n_acc[0])
        if rows >= len(te):                                    # every requested output is already written
            probe.status[:] = 1
            return probe, None
        y_now = run.final_state()
    h_avg = (t_now - float(t_span[0])) / max(n_acc, 1)
    if (float(t_span[1]) - t_now) / h_avg <= LSODA_SWITCH_STEPS:            # accuracy limited: stay explicit
        return dev.rk45(y0, t_span, rtol=rtol, atol=atol, first_step=first_step, t_eval=te), None
    rtol_d, atol_d = device_tolerances(rtol, atol)
    rest = solve_bdf(dev, y_now, (t_now, float(t_span[1])), rtol=rtol_d, atol=atol_d, first_step=min(h_avg, float(t_span[1]) - t_now),
                     t_eval=te[rows:])
    y = torch.cat([probe.y[:, :rows], rest.y], dim=1)
    n_rows = np.array([rows + (len(te) - rows if int(rest.status[0]) == 1 else int(rest.n_rows[0]))], dtype=np.int32)
    return Rk45Result(te, y, rest.status, probe.n_accepted + rest.n_accepted, probe.n_rejected + rest.n_rejected,
                      probe.nfev + rest.nfev, None, n_rows), t_now


def _dae_residual_warning(c: np.ndarray, host: HostSystem, atol: float) -> None:
    """How well the inlet-node algebraic constraint c_in = sum w c_out was kept (pfr_system.py:198-212)."""
    cpl = host.coupling
    S, _, T = c.shape
    acc = np.zeros((S, host.n_points, T))
    for s in range(S):
        np.add.at(acc[s], cpl[:, 0], host.Q_weight[cpl[:, 1]][:, None] * c[s, cpl[:, 2], :])
    nodes = np.unique(cpl[:, 0])
    max_error = float(np.max(np.abs(acc[:, nodes, :] - c[:, nodes, :])))
    if max_error > atol:
        print(f"WARNING: Maximum absolute error in solving inlet BC for PFRs is {max_error:.3e} but absolute tolerance is {atol:.3e}.")
