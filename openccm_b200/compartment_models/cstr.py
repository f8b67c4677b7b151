"""
CSTR network construction: restatement of ``openccm.compartment_models.cstr.connect_cstr_compartments``
(cstr.py:34-135) and ``create_cstr_network`` (cstr.py:138-251) on flat arrays.

The reference walks ``compartment_network`` (dict of dict of dict) facet by facet in Python, keeps the connections in dicts
keyed by signed ids and renumbers them afterwards.  Here the facets are flattened once; inflow signs, net flows per compartment
pair, the first-encounter rule, the flow threshold and the sequential connection numbering are array operations, and the result
is produced directly as ``NetworkArrays`` (the flat form ``export_system`` turns into CSR) — the reference's 5-tuple is derived
from it for callers that want the drop-in signature.  Floating-point sums are accumulated in the reference's order (facet order
inside a pair, set-iteration order of a compartment's elements), so flows and volumes are bit-identical.
Host code: this is the producer of the hot path's input, not a kernel.
"""
from __future__ import annotations

from typing import Dict, List, Set, Tuple

import numpy as np

from ..network import NetworkArrays
from .helpers import tweak_final_flows


def _sequential_group_sums(values: np.ndarray, starts: np.ndarray, lengths: np.ndarray) -> np.ndarray:
    """Left-to-right sums of consecutive groups (Python's ``acc += v`` order): one vector add per position inside a group."""
    out = np.zeros(len(starts))
    for r in range(int(lengths.max()) if len(lengths) else 0):
        sel = lengths > r
        out[sel] = out[sel] + values[starts[sel] + r]
    return out


def _flatten_compartment_network(compartment_network: Dict[int, Dict[int, Dict[int, int]]]):
    """One row per (compartment, neighbour, facet) in iteration order, plus the (compartment, neighbour) groups."""
    g_comp, g_nb, g_len, facets, elems = [], [], [], [], []
    for id_compartment, neighbours in compartment_network.items():
        for id_neighbour, facet_dict in neighbours.items():
            g_comp.append(id_compartment); g_nb.append(id_neighbour); g_len.append(len(facet_dict))
            facets.extend(facet_dict.keys()); elems.extend(facet_dict.values())
    i64 = lambda x: np.asarray(x, dtype=np.int64)
    return i64(g_comp), i64(g_nb), i64(g_len), i64(facets), i64(elems)


def connect_arrays(compartment_network, mesh, flows_and_upwind: np.ndarray, check_level: int, flow_threshold: float):
    """Array form of connect_cstr_compartments.  Returns, per (compartment, neighbour) group in iteration order:
    (g_comp, g_nb, conn) with conn = signed 1-based connection id as seen from g_comp (+ inlet, - outlet, 0 = no connection),
    and the flows of connections 1..n."""
    g_comp, g_nb, g_len, facets, elems = _flatten_compartment_network(compartment_network)
    n_groups = len(g_comp)
    starts = np.concatenate([[0], np.cumsum(g_len)[:-1]]).astype(np.int64) if n_groups else np.zeros(0, dtype=np.int64)
    flow = flows_and_upwind[facets, 0].astype(np.float64)
    flag = flows_and_upwind[facets, 1].astype(np.int64)
    facet_elements = mesh.facet_elements                               # indexable by facet id: tuple of (element, element) pairs
    # element upwind of the facet (cstr.py:113): mesh.facet_elements[facet][flag]; boundary facets (flag == -1) are inflows
    upwind = np.array([facet_elements[f][k] if k >= 0 else -1 for f, k in zip(facets.tolist(), flag.tolist())], dtype=np.int64)
    inflow = (flag == -1) | (elems != upwind)
    signed = np.where(inflow, -flow, flow)
    net = _sequential_group_sums(signed, starts, g_len)
    # first encounter of a compartment pair decides (cstr.py:96-121); boundary neighbours (negative ids) are always new
    lo, hi = np.minimum(g_comp, g_nb), np.maximum(g_comp, g_nb)
    key = lo * (int(hi.max()) + 2 if n_groups else 1) + hi
    internal = g_nb >= 0
    first = np.ones(n_groups, dtype=bool)
    if internal.any():
        _, first_idx, inverse = np.unique(key[internal], return_index=True, return_inverse=True)
        idx_int = np.nonzero(internal)[0]
        first[idx_int] = first_idx[inverse] == np.arange(internal.sum())
        partner_of = np.full(n_groups, -1, dtype=np.int64)
        partner_of[idx_int] = idx_int[first_idx[inverse]]           # group that decided this pair
    else:
        partner_of = np.full(n_groups, -1, dtype=np.int64)
    keep = first & ~((check_level >= 1) & (np.abs(net) < flow_threshold))
    ids = np.zeros(n_groups, dtype=np.int64)
    ids[keep] = np.arange(1, keep.sum() + 1)                          # sequential numbering in encounter order (cstr.py:123-124)
    conn = np.zeros(n_groups, dtype=np.int64)
    conn[keep] = np.where(net[keep] < 0, ids[keep], -ids[keep])      # outward normals: negative net = flow into the compartment
    second = ~first
    conn[second] = -conn[partner_of[second]]                          # the other side sees the same connection reversed
    return g_comp, g_nb, conn, np.abs(net[keep])


def connect_cstr_compartments(compartment_network, mesh, flows_and_upwind, check_level, config_parser):
    """Same signature and return value as the reference (cstr.py:34-135): (connection_pairing, volumetric_flows)."""
    flow_threshold = config_parser.get_item(['COMPARTMENT MODELLING', 'flow_threshold'], float)
    g_comp, g_nb, conn, flows = connect_arrays(compartment_network, mesh, flows_and_upwind, check_level, flow_threshold)
    pairing: Dict[int, Dict[int, int]] = {int(c): {} for c in compartment_network}
    for c, nb, k in zip(g_comp.tolist(), g_nb.tolist(), conn.tolist()):
        if k:
            pairing[c][k] = nb
    if check_level == 2:
        for c, d in pairing.items():
            assert len(d) > 1                                         # cstr.py:129
    return pairing, {i + 1: float(v) for i, v in enumerate(flows)}


def _check_connected(n_models_ids: np.ndarray, g_comp: np.ndarray, g_nb: np.ndarray, conn: np.ndarray) -> None:
    """helpers.py:44-72 with scipy's connected components; boundary neighbours do not join compartments."""
    import scipy.sparse as sp
    from scipy.sparse.csgraph import connected_components
    ids = {int(c): i for i, c in enumerate(n_models_ids.tolist())}
    sel = (conn != 0) & (g_nb >= 0)
    r = np.array([ids[int(c)] for c in g_comp[sel]], dtype=np.int64); c = np.array([ids[int(x)] for x in g_nb[sel]], dtype=np.int64)
    n = len(ids)
    n_sub, _ = connected_components(sp.coo_matrix((np.ones(len(r)), (r, c)), shape=(n, n)), directed=False)
    if n_sub > 1:
        raise ValueError(f"There are {n_sub} disconnected subgraphs in the compartment network."
                         f"Check min_magnitude_threshold, flow_threshold, and boundary conditions (especially any internal ones).")


def create_cstr_network_arrays(compartments: Dict[int, Set[int]], compartment_network, mesh, flows_and_upwind: np.ndarray,
                               config_parser) -> NetworkArrays:
    """Compartments -> CSTR network as flat arrays: one CSTR per compartment (cstr.py:138-251)."""
    print('Creating CSTR network')
    atol_opt = config_parser.get_item(['COMPARTMENT MODELLING', 'atol_opt'], float)
    flow_threshold = config_parser.get_item(['COMPARTMENT MODELLING', 'flow_threshold'], float)
    g_comp, g_nb, conn, flows = connect_arrays(compartment_network, mesh, flows_and_upwind, 2, flow_threshold)
    comp_ids = np.asarray(list(compartment_network.keys()), dtype=np.int64)
    counts = np.bincount(np.searchsorted(np.sort(comp_ids), g_comp[conn != 0]), minlength=len(comp_ids))
    assert np.all(counts > 1)                                        # cstr.py:129: at least two connections per compartment
    _check_connected(comp_ids, g_comp, g_nb, conn)
    # volumes: element sizes summed in the iteration order of the compartment's element set (cstr.py:203-205)
    sizes = np.asarray(mesh.element_sizes, dtype=np.float64)
    volumes = np.zeros(len(compartments))
    for id_compartment, elements in compartments.items():
        el = np.fromiter(elements, dtype=np.int64, count=len(elements))
        volumes[id_compartment] = np.add.accumulate(sizes[el])[-1] if len(el) else 0.0
    # connection ids 1..n -> 0..n-1 (cstr.py:211-221); inlet (+) and outlet (-) ends, grouped by model in encounter order
    used = conn != 0
    is_in = used & (conn > 0)
    is_out = used & (conn < 0)
    i64 = lambda x: np.ascontiguousarray(x, dtype=np.int64)
    in_order = np.argsort(g_comp[is_in], kind='stable'); out_order = np.argsort(g_comp[is_out], kind='stable')
    assert np.all(np.bincount(g_comp[is_in], minlength=len(volumes)) > 0) and np.all(np.bincount(g_comp[is_out], minlength=len(volumes)) > 0)
    m2e = [[(0.0, element) for element in compartments[c]] for c in sorted(compartments.keys())]      # cstr.py:246-248
    net = NetworkArrays(len(volumes), volumes, flows.copy(),
                        i64(g_comp[is_in][in_order]), i64(conn[is_in][in_order] - 1), i64(g_nb[is_in][in_order]),
                        i64(g_comp[is_out][out_order]), i64(-conn[is_out][out_order] - 1), i64(g_nb[is_out][out_order]), m2e)
    connections = net.to_model_network()[0]
    tweak_final_flows(connections, net.flows, mesh.grouped_bcs, atol_opt)                               # cstr.py:241 (in place)
    print("Done creating CSTR network")
    return net


def create_cstr_network(compartments, compartment_network, mesh, flows_and_upwind, dir_vec, config_parser):
    """Drop-in signature of the reference: (connections, volumes, volumetric_flows, compartment_to_cstr_map, cstr_to_element_map)."""
    net = create_cstr_network_arrays(compartments, compartment_network, mesh, flows_and_upwind, config_parser)
    connections, volumes, flows, _, m2e = net.to_model_network()
    compartment_to_cstr_map: Dict[int, List[int]] = {cstr: [cstr] for cstr in range(len(connections))}
    return connections, volumes, flows, compartment_to_cstr_map, m2e
