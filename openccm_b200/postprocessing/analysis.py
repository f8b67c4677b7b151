"""
Residence-time distribution on the device — replacement for ``openccm.postprocessing.analysis.network_to_rtd``
(/root/reference/openccm/postprocessing/analysis.py:34-103).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .. import _lib


def outlet_selection(outlets: Sequence[Tuple[int, int]], q_connections: np.ndarray, P: int, S: int):
    """Flat state elements (device layout p * S + s) of every (outlet link, species) with its flow weight."""
    idx, w, sp = [], [], []
    for model, conn in outlets:
        node = P * (model + 1) - 1                    # analysis.py:85
        for s in range(S):
            idx.append(node * S + s); w.append(float(q_connections[conn])); sp.append(s)
    return np.asarray(idx, dtype=np.int32), np.asarray(w, dtype=np.float64), np.asarray(sp, dtype=np.int32)


def rtd_on_device(out: torch.Tensor, t: torch.Tensor, sel_pos: torch.Tensor, sel_w: torch.Tensor, sel_species: torch.Tensor,
                  S: int, q_in_total: float):
    """out: device [M, T, n_save]; sel_pos indexes the last axis. Returns (rtd [M, S, T, 3], moments [M, S, 3])."""
    lib = _lib.load()
    M, T, n_save = out.shape
    rtd = torch.empty((M, S, T, 3), dtype=torch.float64, device=out.device)
    mom = torch.empty((M, S, 3), dtype=torch.float64, device=out.device)
    per_member = 1 if t.dim() == 2 else 0
    _lib.check(lib.occm_rtd(out.data_ptr(), M, T, n_save, t.data_ptr(), per_member, sel_pos.data_ptr(), sel_w.data_ptr(),
                            sel_species.data_ptr(), int(sel_pos.numel()), S, float(q_in_total), rtd.data_ptr(),
                            mom.data_ptr(), torch.cuda.current_stream().cuda_stream))
    return rtd, mom


def pulse_moments(u, t0: float, t1: float, n: int = 200_001) -> Tuple[float, float, float]:
    """(area, mean, variance) of an inlet pulse u(t) supported in [t0, t1] (composite Simpson on n points)."""
    from scipy.integrate import simpson
    t = np.linspace(t0, t1, n | 1)
    v = np.asarray([u(x) for x in t], dtype=np.float64) if not isinstance(u, np.ndarray) else u
    area = simpson(v, x=t)
    mean = simpson(t * v, x=t) / area
    return float(area), float(mean), float(simpson((t - mean) ** 2 * v, x=t) / area)


def pulse_to_rtd(system_results, c_mesh, config_parser, network, pulse: Tuple[float, float, float], *,
                 source_flow: Optional[float] = None):
    """RTD from a PULSE experiment on the same solver kernels (north star item 4): ``system_results`` is the solve of the tracer
    injected as a narrow pulse — an inlet boundary condition such as ``a: inlet -> (H(t) - H(t - w)) / w`` or a point-mass
    initial condition (boundary_and_initial_conditions.py:174-180) — and ``pulse`` its (area, mean, variance) in time.
    Returns (rtd[S, T, 3] = (t, E, F), moments[S, 3] = (m0, mean, variance) of E with the pulse's own moments removed).
    ``source_flow``: the flow the pulse area refers to (default: the total inlet flow, i.e. an inlet-concentration pulse)."""
    id_inlet = c_mesh.grouped_bcs.id(config_parser.get_item(['POST-PROCESSING', 'inlet_bc_name'], str))
    id_outlet = c_mesh.grouped_bcs.id(config_parser.get_item(['POST-PROCESSING', 'outlet_bc_name'], str))
    P = config_parser.get_item(['SIMULATION', 'points_per_pfr'], int)
    species = config_parser.get_list(['SIMULATION', 'specie_names'], str)
    y, t, inlets_map, outlets_map = system_results
    q = np.asarray(network[2] if not hasattr(network, 'flows') else network.flows)
    S = len(species)
    q_in = float(sum(q[conn] for _, conn in inlets_map[id_inlet])) if source_flow is None else float(source_flow)
    idx, w, sp = outlet_selection(outlets_map[id_outlet], q, P, S)
    dev = torch.device('cuda', torch.cuda.current_device())
    sel = np.ascontiguousarray(np.asarray(y)[sp, idx // S, :].T)[None]
    out = torch.as_tensor(sel, device=dev)
    M, T, n_save = out.shape
    rtd = torch.empty((M, S, T, 3), dtype=torch.float64, device=dev)
    mom = torch.empty((M, S, 3), dtype=torch.float64, device=dev)
    tt = torch.as_tensor(np.asarray(t, dtype=np.float64), device=dev)
    pos = torch.arange(len(idx), dtype=torch.int32, device=dev)
    wv, spv = torch.as_tensor(w, device=dev), torch.as_tensor(sp, device=dev)
    lib = _lib.load()
    _lib.check(lib.occm_rtd_pulse(out.data_ptr(), M, T, n_save, tt.data_ptr(), 0, pos.data_ptr(), wv.data_ptr(), spv.data_ptr(),
                                  int(pos.numel()), S, q_in, float(pulse[0]), float(pulse[1]), float(pulse[2]), rtd.data_ptr(),
                                  mom.data_ptr(), torch.cuda.current_stream().cuda_stream))
    return rtd[0].cpu().numpy(), mom[0].cpu().numpy()


def network_to_rtd(system_results, c_mesh, config_parser, network, *, return_moments: bool = False):
    """Same signature and return value as the reference: rtd[S, T, 3] = (t, E, F) for a unit-step tracer."""
    print("Start calculating RTD")
    id_inlet = c_mesh.grouped_bcs.id(config_parser.get_item(['POST-PROCESSING', 'inlet_bc_name'], str))
    id_outlet = c_mesh.grouped_bcs.id(config_parser.get_item(['POST-PROCESSING', 'outlet_bc_name'], str))
    P = config_parser.get_item(['SIMULATION', 'points_per_pfr'], int)
    species = config_parser.get_list(['SIMULATION', 'specie_names'], str)
    y, t, inlets_map, outlets_map = system_results
    q = np.asarray(network[2] if not hasattr(network, 'flows') else network.flows)
    S = len(species)
    q_in = float(sum(q[conn] for _, conn in inlets_map[id_inlet]))          # analysis.py:78
    idx, w, sp = outlet_selection(outlets_map[id_outlet], q, P, S)
    dev = torch.device('cuda', torch.cuda.current_device())
    # gather the outlet rows of c[S, n_points, T] into out[1, T, n_sel]
    nodes = idx // S
    sel = np.ascontiguousarray(np.asarray(y)[sp, nodes, :].T)[None]          # [1, T, n_sel]
    out = torch.as_tensor(sel, device=dev)
    pos = torch.arange(len(idx), dtype=torch.int32, device=dev)
    rtd, mom = rtd_on_device(out, torch.as_tensor(np.asarray(t, dtype=np.float64), device=dev), pos,
                             torch.as_tensor(w, device=dev), torch.as_tensor(sp, device=dev), S, q_in)
    print('Done calculating RTD')
    res = rtd[0].cpu().numpy()
    return (res, mom[0].cpu().numpy()) if return_moments else res
