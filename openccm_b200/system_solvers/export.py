"""
Export layer: (model network, CONFIG, reactions file) -> flat host tables for the device.

Mirrors, in vectorised numpy, what the reference does with Python loops before it calls ``solve_ivp``:

* CSTR assembly                 /root/reference/openccm/system_solvers/cstr_system.py:113-135
* PFR assembly                  .../pfr_system.py:105-174
* initial conditions            .../boundary_and_initial_conditions.py:254-389   (480 s at 2e6 DOF in the reference)
* boundary conditions           .../boundary_and_initial_conditions.py:46-247
* DOF <-> element map           /root/reference/openccm/mesh/__init__.py:142-198
* reactions                     .../reactions.py:41-96 (via tables.py)

The result (``HostSystem``) is plain numpy; ``device.py`` uploads it once.  Nothing here touches the GPU.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import sympy as sp

from ..network import NetworkArrays
from .tables import H, MechanismTables, T_SYM, read_reactions_file

_X, _Y, _Z = sp.symbols('x y z')


@dataclass
class HostSystem:
    model: str                       # 'cstr' | 'pfr'
    species: List[str]
    n_models: int
    P: int                           # points per model (1 for CSTR)
    row_ptr: np.ndarray              # int32 [n_models + 1]
    row_src: np.ndarray              # int32 [nnz]
    row_w: np.ndarray                # f64   [nnz]
    a: Optional[np.ndarray]          # f64 [n_models] (PFR: Q_pfr / dV)
    pulse_ptr: Optional[np.ndarray]
    pulse_fn: Optional[np.ndarray]
    pulse_w: Optional[np.ndarray]
    tables: MechanismTables
    c0: np.ndarray                   # f64 [S, n_points]  (reference layout)
    inlet_map: Dict[int, List[Tuple[int, int]]]
    outlet_map: Dict[int, List[Tuple[int, int]]]
    flows: np.ndarray
    # PFR extras kept for the post-solve DAE residual check (pfr_system.py:198-212)
    coupling: Optional[np.ndarray] = None     # [n, 3] (inlet node, connection, outlet node)
    Q_weight: Optional[np.ndarray] = None
    fields: Optional[np.ndarray] = None       # [n_fields, n_points]
    c0_bc: Optional[np.ndarray] = None        # [S, n_points] part of c0 that is the folded PFR boundary value (scales with
                                              # the boundary values, not with the initial condition, in ensembles)

    @property
    def S(self) -> int:
        return len(self.species)

    @property
    def n_points(self) -> int:
        return self.n_models * self.P

    @property
    def ndof(self) -> int:
        return self.n_points * self.S

    def to_device_layout(self, y_ref: np.ndarray) -> np.ndarray:
        """reference flattening y[s * n_points + p]  ->  device flattening y[p * S + s] (last axis)."""
        lead = y_ref.shape[:-1]
        return np.ascontiguousarray(np.swapaxes(y_ref.reshape(lead + (self.S, self.n_points)), -1, -2)).reshape(lead + (-1,))

    def to_reference_layout(self, y_dev: np.ndarray) -> np.ndarray:
        lead = y_dev.shape[:-1]
        return np.ascontiguousarray(np.swapaxes(y_dev.reshape(lead + (self.n_points, self.S)), -1, -2)).reshape(lead + (-1,))


# ------------------------------------------------------------------------------------------- DOF -> element map
def element_dofs(model_to_element_map: Optional[list], n_models: int, P: int) -> Tuple[np.ndarray, np.ndarray]:
    """For every (model, element) pair the DOF the element is averaged into: (dof[], element[]).

    Vectorised form of create_dof_to_element_map (mesh/__init__.py:142-198) for distance-sorted element lists:
    point i < P-1 takes the not-yet-taken elements with dist <= dof_dist[i] + delta, the last point takes the rest."""
    if model_to_element_map is None:
        model_to_element_map = [[(0.0, i)] for i in range(n_models)]
    counts = np.fromiter((len(m) for m in model_to_element_map), dtype=np.int64, count=len(model_to_element_map))
    total = int(counts.sum())
    dist = np.empty(total); elem = np.empty(total, dtype=np.int64)
    k = 0
    for m in model_to_element_map:
        n = len(m)
        if n:
            arr = np.asarray(m, dtype=np.float64)
            dist[k:k + n] = arr[:, 0]; elem[k:k + n] = arr[:, 1].astype(np.int64)
            k += n
    model = np.repeat(np.arange(len(model_to_element_map), dtype=np.int64), counts)
    if P == 1:
        return model, elem
    dof_dist = np.linspace(0.0, 1.0, num=P)
    delta = 1.0 / (P - 1) / 2
    thr = dof_dist[:P - 1] + delta
    thr[0] = delta
    # prefix semantics of the reference: an element is taken by the first point whose threshold is not exceeded
    # by it *or by any earlier element of the same model* (lists are distance sorted, so this is a running max)
    run = dist.copy()
    starts = np.concatenate([[0], np.cumsum(counts)[:-1]])
    for s0, n in zip(starts.tolist(), counts.tolist()):
        if n > 1:
            np.maximum.accumulate(run[s0:s0 + n], out=run[s0:s0 + n])
    idx = np.searchsorted(thr, run, side='left')            # first i with run <= thr[i]; P-1 if none
    return model * P + idx, elem


def element_interpolation(model_to_element_map: Optional[list], n_models: int, P: int, n_elements: int):
    """Per mesh element the two DOFs and the weight its value is interpolated from:
    ``(el_dof, el_other, el_w)`` with ``value = c[el_dof] * el_w + c[el_other] * (1 - el_w)`` and ``el_dof = -1`` for
    elements outside every compartment.

    Vectorised form of create_dof_to_element_map (mesh/__init__.py:142-198) for distance-sorted element lists.  Along a
    PFR the reference walks the list once; point 0 takes the prefix with dist <= delta, a middle point i first the
    prefix with dist <= dof_dist[i] (interpolating towards point i-1), then the prefix with dist <= dof_dist[i] + delta
    (towards point i+1), the last point the rest — "prefix" meaning up to the first element that exceeds the bound."""
    if model_to_element_map is None:
        model_to_element_map = [[(0.0, i)] for i in range(n_models)]
    counts = np.fromiter((len(m) for m in model_to_element_map), dtype=np.int64, count=len(model_to_element_map))
    total = int(counts.sum())
    dist = np.empty(total); elem = np.empty(total, dtype=np.int64)
    k = 0
    for m in model_to_element_map:
        n = len(m)
        if n:
            arr = np.asarray(m, dtype=np.float64)
            dist[k:k + n] = arr[:, 0]; elem[k:k + n] = arr[:, 1].astype(np.int64)
            k += n
    model = np.repeat(np.arange(len(model_to_element_map), dtype=np.int64), counts)
    ndof = len(model_to_element_map) * P
    if P == 1:
        # the "last point" branch with delta = 1e20 (:160,:171-174): weight clips to 1, dof_other = dof - 1 (index -1 wraps)
        dof = model
        other = (dof - 1) % ndof
        w = np.clip(1.0 + (dist - 1.0) / (2 * 1e20), 0.0, 1.0)
    else:
        dof_dist = np.linspace(0.0, 1.0, num=P)
        delta = 1.0 / (P - 1) / 2
        # bounds of the successive prefixes: [delta | d_1, d_1 + delta | d_2, d_2 + delta | ...]; past the end = last point
        bounds = np.empty(1 + 2 * (P - 2))
        bounds[0] = delta
        bounds[1::2] = dof_dist[1:P - 1]
        bounds[2::2] = dof_dist[1:P - 1] + delta
        run = dist.copy()
        starts = np.concatenate([[0], np.cumsum(counts)[:-1]])
        for s0, n in zip(starts.tolist(), counts.tolist()):
            if n > 1:
                np.maximum.accumulate(run[s0:s0 + n], out=run[s0:s0 + n])
        seg = np.searchsorted(bounds, run, side='left')
        last = seg == len(bounds)
        point = np.where(last, P - 1, (seg + 1) // 2)
        before = (~last) & (seg % 2 == 1)                    # middle point, interpolating towards point - 1
        first = seg == 0
        dof = model * P + point
        other = np.where(first | ((~last) & ~before), dof + 1, dof - 1)
        with np.errstate(invalid='ignore'):
            w = np.where(last, 1.0 + (dist - 1.0) / (2 * delta),
                 np.where(first, 1.0 - dist / (2 * delta),
                 np.where(before, (dist - dof_dist[np.maximum(point - 1, 0)]) / (2 * delta),
                          (dof_dist[np.minimum(point + 1, P - 1)] - dist) / (2 * delta))))
        w = np.clip(w, 0.0, 1.0)
    el_dof = -np.ones(n_elements, dtype=np.int32)
    el_other = np.zeros(n_elements, dtype=np.int32)
    el_w = np.zeros(n_elements, dtype=np.float64)
    el_dof[elem] = dof; el_other[elem] = other; el_w[elem] = w      # an element listed twice keeps its last entry (:322 overwrites)
    return el_dof, el_other, el_w


# ------------------------------------------------------------------------------------------- initial conditions
def _lambdify_xyz(text: str):
    return sp.lambdify([_X, _Y, _Z, T_SYM], text, 'numpy')


def initial_conditions(cp, species: List[str], cmesh, dofs: np.ndarray, elems: np.ndarray, n_points: int, P: int,
                       coupling: Optional[np.ndarray], Q_weight: Optional[np.ndarray]) -> Tuple[np.ndarray, str]:
    """c0[S, n_points] and the boundary-condition string extended by the point-mass pseudo-BCs
    (boundary_and_initial_conditions.py:254-389)."""
    t0 = cp.get_list(['SIMULATION', 't_span'], float)[0]
    if len(set(species)) != len(species):
        raise AssertionError('duplicate specie names')
    funcs = {s: ([], [], []) for s in species}
    bc_string = cp.get_item(['SIMULATION', 'boundary_conditions'], str)
    for line in cp.get_item(['SIMULATION', 'initial_conditions'], str).splitlines():
        if not line.strip():
            continue
        specie, f = [v.strip() for v in line.split('->')]
        if specie not in funcs:
            raise ValueError(f'Unknown specie: {specie} when specifying initial conditions')
        funcs[specie][2 if 'point' in f else 1 if '@' in f else 0].append(f)

    S = len(species)
    c0 = np.zeros((S, n_points))
    is_inlet = (np.arange(n_points) % P == 0) if P > 1 else np.zeros(n_points, dtype=bool)
    keep = ~is_inlet[dofs]                                   # nothing is mapped onto PFR inlet DOFs (:355-356)
    d_keep, e_keep = dofs[keep], elems[keep]
    n_el_per_dof = np.bincount(d_keep, minlength=n_points)
    mapped = n_el_per_dof > 0
    to_interp = (~mapped) & (~is_inlet)
    if int(to_interp.sum()) == n_points:                     # :318-319 (can only trigger for CSTRs)
        raise AssertionError("No DOF maps to an mesh element.")

    cent = np.asarray(cmesh.element_centroids, dtype=np.float64)
    xs, ys = cent[:, 0], cent[:, 1]
    zs = cent[:, 2] if cent.shape[1] > 2 else np.zeros(len(cent))
    n_el = len(cmesh.element_sizes)
    for i, specie in enumerate(species):
        full, restricted, point = funcs[specie]
        if not (full or restricted or point):
            raise ValueError(f"Initial conditions were not specified for specie: {specie}")
        if len(full) > 1:
            raise ValueError(f"Only one default IC value can be specified per specie.Specie {specie} had {full} specified.")
        buf = np.zeros(n_el)
        if full:
            buf[:] = np.broadcast_to(_lambdify_xyz(full[0])(xs, ys, zs, t0), (n_el,))
        for r in restricted:
            e_str, p_str = r.split('@')
            mask = np.broadcast_to(_lambdify_xyz(p_str)(xs, ys, zs, t0), (n_el,)).astype(bool)
            vals = np.broadcast_to(_lambdify_xyz(e_str)(xs, ys, zs, t0), (n_el,))
            buf[mask] = vals[mask]
        for pm in point:                                      # -> smoothed pulse sources (:346-350)
            mass, pt = pm.replace(' ', '').split('@')
            bc_string += f"{specie}: {pt} -> {mass or '1.0'}\n"
        sums = np.bincount(d_keep, weights=buf[e_keep], minlength=n_points)
        c0[i, mapped] = sums[mapped] / n_el_per_dof[mapped]

    if to_interp.any() and mapped.any():                      # :361-380, mean of the nearest mapped non-inlet DOFs
        # (with no mapped DOF at all the reference reads c0[:, -1] == 0 for both neighbours: zeros stay zeros)
        idx = np.arange(n_points)
        before = np.maximum.accumulate(np.where(mapped, idx, -1))
        after = np.minimum.accumulate(np.where(mapped, idx, n_points)[::-1])[::-1]
        after = np.where(after == n_points, -1, after)
        b = np.where(before == -1, after, before)
        a = np.where(after == -1, before, after)
        ti = np.nonzero(to_interp)[0]
        c0[:, ti] = (c0[:, b[ti]] + c0[:, a[ti]]) / 2

    if coupling is not None and len(coupling):                # :383-387
        c0[:, coupling[:, 0]] = 0
        contrib = Q_weight[coupling[:, 1]][None, :] * c0[:, coupling[:, 2]]
        for i in range(S):
            np.add.at(c0[i], coupling[:, 0], contrib[i])
    return c0, bc_string


# ------------------------------------------------------------------------------------------- boundary conditions
@dataclass
class BCPlan:
    """Boundary groups in the order the generated module applies them (sorted cleaned names,
    boundary_and_initial_conditions.py:209,236-239)."""
    names: List[str] = field(default_factory=list)
    ids: List[int] = field(default_factory=list)
    exprs: List[Dict[int, sp.Expr]] = field(default_factory=list)   # species index -> g(t) (already differentiated for PFR)
    points: List[np.ndarray] = field(default_factory=list)
    weights: List[np.ndarray] = field(default_factory=list)
    is_pulse: List[bool] = field(default_factory=list)
    c0_bc: Optional[np.ndarray] = None


def boundary_conditions(c0: np.ndarray, cp, species: List[str], bc_string: str, points_for_bc: Dict[int, np.ndarray],
                        weights_for_bc: Dict[int, np.ndarray], t0: float, P: int, cmesh, dofs: np.ndarray,
                        elems: np.ndarray, volumes: np.ndarray) -> BCPlan:
    """Parse ``boundary_conditions`` (+ point-mass pseudo-BCs), fold PFR boundary values into c0.
    boundary_and_initial_conditions.py:46-247."""
    need_deriv = P > 1
    plan = BCPlan()
    by_name: Dict[str, int] = {}
    seen: Dict[int, List[str]] = {}
    for_eval: List[Tuple[int, int, sp.Expr]] = []
    n_points = c0.shape[1]

    nearest = None
    if 'point' in bc_string:                                  # :130-149
        from scipy.spatial import KDTree
        is_inlet = (dofs % P == 0) if P > 1 else np.zeros(len(dofs), dtype=bool)
        sub_dofs = dofs[~is_inlet]
        order = np.argsort(sub_dofs, kind='stable')           # the reference enumerates DOFs in ascending order
        sub_dofs = sub_dofs[order]
        tree = KDTree(np.asarray(cmesh.element_centroids)[elems[~is_inlet][order]])
        nearest = lambda pos: int(sub_dofs[tree.query(pos)[1]])

    no_flux = getattr(cmesh.grouped_bcs, 'no_flux_names', ())
    for line in bc_string.splitlines():
        if not line.strip():
            continue
        specie, info = [v.strip() for v in line.split(':')]
        if specie not in species:
            raise ValueError(f'Unknown specie: {specie} when specifying boundary condition: {line}.')
        i_specie = species.index(specie)
        name, eqn_str = [v.strip() for v in info.split('->')]
        is_point = 'point' in name
        if name in no_flux:
            raise ValueError(f'BC value specified for the no-flux bc {name}.')
        cleaned = name
        for ch in '().,':
            cleaned = cleaned.replace(ch, '_')
        if cleaned not in by_name:
            bc_id = hash(cleaned) if 'point' in cleaned else cmesh.grouped_bcs.id(cleaned)
            by_name[cleaned] = len(plan.names)
            plan.names.append(cleaned); plan.ids.append(bc_id); plan.exprs.append({})
            plan.points.append(np.asarray(points_for_bc.get(bc_id, np.zeros(0, dtype=np.int64)), dtype=np.int64))
            plan.weights.append(np.asarray(weights_for_bc.get(bc_id, np.zeros(0)), dtype=np.float64))
            plan.is_pulse.append(is_point)
        g = by_name[cleaned]
        bc_id = plan.ids[g]
        if not is_point:
            if specie in seen.setdefault(bc_id, []):
                raise ValueError(f"Specie {specie} has multiple BCs specified for boundary {name}.")
            seen[bc_id].append(specie)
        else:
            dof = nearest(eval(name.replace('point', ''), {'__builtins__': {}}))
            model = int(dof / P)
            pts = np.array([model]) if P == 1 else np.arange(model * P + 1, (model + 1) * P)
            plan.points[g] = pts
            plan.weights[g] = np.ones(len(pts))
            eqn_str = f"{float(eqn_str) / (0.01 * volumes[model])} * H(t) * H(-(t - 2*0.01))"
        eqn = sp.parse_expr(eqn_str, local_dict={'H': H})
        free = eqn.free_symbols
        if free & {_X, _Y, _Z}:
            raise ValueError(f"Boundary condition {eqn} is written in terms of a spatial coordinate (x, y, z).")
        if free != {T_SYM} and len(free) != 0:
            raise ValueError(f"Boundary condition {eqn} uses a variable other than t (time).")
        if is_point or not need_deriv:
            plan.exprs[g][i_specie] = eqn
        else:
            plan.exprs[g][i_specie] = eqn.diff(T_SYM)
            for_eval.append((g, i_specie, eqn))
            c0[i_specie, plan.points[g]] = 0                  # :199
    plan.c0_bc = np.zeros_like(c0)
    for g, i_specie, eqn in for_eval:                         # :202-207
        np.add.at(plan.c0_bc[i_specie], plan.points[g], plan.weights[g] * float(eqn.evalf(subs={'t': t0})))
    c0 += plan.c0_bc

    order = sorted(range(len(plan.names)), key=lambda k: plan.names[k])
    for attr in ('names', 'ids', 'exprs', 'points', 'weights', 'is_pulse'):
        setattr(plan, attr, [getattr(plan, attr)[k] for k in order])
    return plan


def _last_write_wins(points: np.ndarray, weights: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """``ddt[s, points] += g * weights`` with repeated points keeps only the LAST weight of each point
    (numpy/numba fancy-index ``+=``, boundary_and_initial_conditions.py:238)."""
    if len(points) == 0:
        return points, weights
    rev_pts = points[::-1]
    uniq, first_in_rev = np.unique(rev_pts, return_index=True)
    return uniq, weights[::-1][first_in_rev]


# ------------------------------------------------------------------------------------------- assembly
def _csr_from_entries(n_rows: int, rows: np.ndarray, src: np.ndarray, w: np.ndarray):
    order = np.argsort(rows, kind='stable')
    rows, src, w = rows[order], src[order], w[order]
    ptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.add.at(ptr, rows + 1, 1)
    return np.cumsum(ptr).astype(np.int32), src.astype(np.int32), w.astype(np.float64)


def _build_tables(cp, species: List[str], plan: BCPlan, fields: Optional[Dict[str, np.ndarray]]):
    tables = MechanismTables(species, list(fields) if fields else [])
    path = cp.get_item(['SIMULATION', 'reactions_file_path'], str)
    if path != 'None':
        rxns, rates, extra = read_reactions_file(path)
        tables.add_reactions(rxns, rates, extra)
    groups = [tables.add_bc_group(e) for e in plan.exprs]
    return tables, groups


def export_system(model: str, network, cp, cmesh, fields: Optional[Dict[str, np.ndarray]] = None) -> HostSystem:
    """Everything the reference computes between entering ``solve_system`` and calling ``solve_ivp``."""
    net = network if isinstance(network, NetworkArrays) else NetworkArrays.from_model_network(network)
    species = cp.get_list(['SIMULATION', 'specie_names'], str)
    t0 = cp.get_list(['SIMULATION', 't_span'], float)[0]
    model = model.lower()
    N = net.n_models
    V, Q = net.volumes, net.flows
    inlet_map, outlet_map = net.inlet_outlet_maps()
    in_bc = net.in_other < 0
    in_int = ~in_bc

    if model != 'pfr':
        P = 1
        dofs, elems = element_dofs(net.model_to_element_map, N, 1)
        c0, bc_string = initial_conditions(cp, species, cmesh, dofs, elems, N, 1, None, None)
        # inlet weights Q/V per boundary id, in iteration order (cstr_system.py:121-126)
        pts_bc: Dict[int, np.ndarray] = {}; w_bc: Dict[int, np.ndarray] = {}
        for bc_id in np.unique(net.in_other[in_bc]).tolist():
            sel = net.in_other == bc_id
            pts_bc[bc_id] = net.in_model[sel]
            w_bc[bc_id] = Q[net.in_conn[sel]] / V[net.in_model[sel]]
        plan = boundary_conditions(c0, cp, species, bc_string, pts_bc, w_bc, t0, 1, cmesh, dofs, elems, V)
        tables, groups = _build_tables(cp, species, plan, fields)
        # transport rows: diagonal -(sum Q_out)/V first, then one entry per inflowing connection (cstr_system.py:128-135)
        diag = -np.bincount(net.out_model, weights=Q[net.out_conn], minlength=N) / V
        rows = [np.arange(N), net.in_model[in_int]]
        srcs = [np.arange(N), net.in_other[in_int]]
        ws = [diag, Q[net.in_conn[in_int]] / V[net.in_model[in_int]]]
        p_rows: List[np.ndarray] = []; p_fn: List[np.ndarray] = []; p_w: List[np.ndarray] = []
        for g, pts, w, pulse in zip(groups, plan.points, plan.weights, plan.is_pulse):
            pts, w = _last_write_wins(pts, w)
            if pulse:                                         # point-mass source: not a flow term (no q_scale / bc_scale)
                p_rows.append(pts); p_fn.append(np.full(len(pts), g)); p_w.append(w)
            else:
                rows.append(pts); srcs.append(np.full(len(pts), -1 - g)); ws.append(w)
        row_ptr, row_src, row_w = _csr_from_entries(N, np.concatenate(rows), np.concatenate(srcs), np.concatenate(ws))
        pulse_ptr = pulse_fn = pulse_w = None
        if p_rows:
            pulse_ptr, pulse_fn, pulse_w = _csr_from_entries(N, np.concatenate(p_rows), np.concatenate(p_fn), np.concatenate(p_w))
        return HostSystem('cstr', species, N, 1, row_ptr, row_src, row_w, None, pulse_ptr, pulse_fn, pulse_w, tables, c0,
                          inlet_map, outlet_map, Q, fields=_stack_fields(fields, N))

    # ---------------------------------------------------------------- PFR (pfr_system.py:100-190)
    P = cp.get_item(['SIMULATION', 'points_per_pfr'], int)
    if P < 2:
        raise AssertionError('points_per_pfr must be >= 2')
    n_points = N * P
    dV = V / (P - 1)
    Q_pfr = np.bincount(net.in_model, weights=Q[net.in_conn], minlength=N)
    Q_weight = Q.copy()
    Q_weight[net.in_conn] = Q[net.in_conn] / Q_pfr[net.in_model]
    tot = np.bincount(net.in_model, weights=Q_weight[net.in_conn], minlength=N)
    if not np.allclose(tot, 1):
        raise AssertionError('inlet flow weights of a PFR do not sum to 1 (pfr_system.py:119-121)')
    coupling = np.stack([P * net.in_model[in_int], net.in_conn[in_int], P * (net.in_other[in_int] + 1) - 1], axis=1) \
        if in_int.any() else np.zeros((0, 3), dtype=np.int64)
    dofs, elems = element_dofs(net.model_to_element_map, N, P)
    c0, bc_string = initial_conditions(cp, species, cmesh, dofs, elems, n_points, P, coupling, Q_weight)
    pts_bc = {}; w_bc = {}
    for bc_id in np.unique(net.in_other[in_bc]).tolist():
        sel = net.in_other == bc_id
        pts_bc[bc_id] = P * net.in_model[sel]
        w_bc[bc_id] = Q_weight[net.in_conn[sel]]
    plan = boundary_conditions(c0, cp, species, bc_string, pts_bc, w_bc, t0, P, cmesh, dofs, elems, V)
    tables, groups = _build_tables(cp, species, plan, fields)
    rows = [net.in_model[in_int]]; srcs = [net.in_other[in_int]]; ws = [Q_weight[net.in_conn[in_int]]]
    p_rows: List[np.ndarray] = []; p_fn: List[np.ndarray] = []; p_w: List[np.ndarray] = []
    for g, pts, w, pulse in zip(groups, plan.points, plan.weights, plan.is_pulse):
        if pulse:
            # applies (weight 1) to every non-inlet node of one PFR (:174-180)
            models = np.unique(pts // P)
            p_rows.append(models); p_fn.append(np.full(len(models), g)); p_w.append(np.ones(len(models)))
        else:
            pts, w = _last_write_wins(pts, w)
            rows.append(pts // P); srcs.append(np.full(len(pts), -1 - g)); ws.append(w)
    row_ptr, row_src, row_w = _csr_from_entries(N, np.concatenate(rows), np.concatenate(srcs), np.concatenate(ws))
    pulse_ptr = pulse_fn = pulse_w = None
    if p_rows:
        pulse_ptr, pulse_fn, pulse_w = _csr_from_entries(N, np.concatenate(p_rows), np.concatenate(p_fn), np.concatenate(p_w))
    return HostSystem('pfr', species, N, P, row_ptr, row_src, row_w, Q_pfr / dV, pulse_ptr, pulse_fn, pulse_w, tables, c0,
                      inlet_map, outlet_map, Q, coupling=coupling, Q_weight=Q_weight, fields=_stack_fields(fields, n_points),
                      c0_bc=plan.c0_bc)


def _stack_fields(fields: Optional[Dict[str, np.ndarray]], n_points: int) -> Optional[np.ndarray]:
    if not fields:
        return None
    arr = np.stack([np.asarray(v, dtype=np.float64).reshape(-1) for v in fields.values()])
    if arr.shape[1] != n_points:
        raise ValueError(f'per-point fields must have {n_points} entries')
    return np.ascontiguousarray(arr)
