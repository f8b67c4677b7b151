"""
Array form of an OpenCCM model network.

The reference hands the solvers a 5-tuple built by ``create_model_network``
(/root/reference/openccm/compartment_models/__init__.py:76-91):

    connections:  Dict[id -> (inlets {conn_id: other}, outlets {conn_id: other})]
                  ``other < 0`` is a domain boundary id (GroupedBCs.id, mesh/cmesh.py:91-115)
    volumes:      f64[n_models]
    volumetric_flows: f64[n_connections]
    compartment_to_model_map, model_to_element_map

Walking Python dicts is what makes the reference's assembly O(N) interpreter time
(cstr_system.py:118-135, pfr_system.py:108-174).  ``NetworkArrays`` is the same
information as flat numpy arrays, one entry per *(connection, side)* in the exact
order the reference iterates them, so that every order-dependent quirk (e.g. the
last-write-wins boundary weights of boundary_and_initial_conditions.py:238) can be
reproduced by vectorised code.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

ModelNetwork = Tuple[Dict[int, Tuple[Dict[int, int], Dict[int, int]]], np.ndarray, np.ndarray, dict, list]


@dataclass
class NetworkArrays:
    """
    Flat description of a model network.

    ``in_model[k], in_conn[k], in_other[k]``: the k-th inlet entry when iterating
    ``for id_model in sorted(connections): for conn, other in connections[id_model][0].items()``.
    ``out_*`` likewise for the outlet dictionaries.
    """
    n_models: int
    volumes: np.ndarray          # f64[n_models]
    flows: np.ndarray            # f64[n_connections]
    in_model: np.ndarray         # i64[n_in]
    in_conn: np.ndarray          # i64[n_in]
    in_other: np.ndarray         # i64[n_in]   (<0: domain inlet BC id)
    out_model: np.ndarray        # i64[n_out]
    out_conn: np.ndarray         # i64[n_out]
    out_other: np.ndarray        # i64[n_out]  (<0: domain outlet BC id)
    model_to_element_map: Optional[list] = None  # List[List[(dist, element)]] or None (synthetic: identity)

    # ------------------------------------------------------------------ conversion
    @staticmethod
    def from_model_network(model_network: Sequence) -> "NetworkArrays":
        """Flatten the reference's 5-tuple (order preserving)."""
        connections, volumes, flows = model_network[0], model_network[1], model_network[2]
        m2e = model_network[4] if len(model_network) > 4 else None
        ids = sorted(connections.keys())
        assert ids == list(range(len(ids))), "model ids must be 0..n-1"
        in_model: List[int] = []; in_conn: List[int] = []; in_other: List[int] = []
        out_model: List[int] = []; out_conn: List[int] = []; out_other: List[int] = []
        for i in ids:
            inlets, outlets = connections[i]
            in_model.extend([i] * len(inlets));   in_conn.extend(inlets.keys());   in_other.extend(inlets.values())
            out_model.extend([i] * len(outlets)); out_conn.extend(outlets.keys()); out_other.extend(outlets.values())
        a = lambda x: np.asarray(x, dtype=np.int64)
        return NetworkArrays(len(ids), np.asarray(volumes, dtype=np.float64), np.asarray(flows, dtype=np.float64),
                             a(in_model), a(in_conn), a(in_other), a(out_model), a(out_conn), a(out_other), m2e)

    def to_model_network(self) -> ModelNetwork:
        """Rebuild the reference's dict-of-dicts 5-tuple (used to drive the oracle / the reference itself)."""
        connections: Dict[int, Tuple[Dict[int, int], Dict[int, int]]] = {i: ({}, {}) for i in range(self.n_models)}
        for m, c, o in zip(self.in_model.tolist(), self.in_conn.tolist(), self.in_other.tolist()):
            connections[m][0][c] = o
        for m, c, o in zip(self.out_model.tolist(), self.out_conn.tolist(), self.out_other.tolist()):
            connections[m][1][c] = o
        m2e = self.model_to_element_map
        if m2e is None:
            m2e = [[(0.0, i)] for i in range(self.n_models)]
        comp_to_model = {i: [i] for i in range(self.n_models)}
        return connections, self.volumes.copy(), self.flows.copy(), comp_to_model, m2e

    # ------------------------------------------------------------------ helpers
    def inlet_outlet_maps(self) -> Tuple[Dict[int, List[Tuple[int, int]]], Dict[int, List[Tuple[int, int]]]]:
        """inlet_map / outlet_map as returned by the reference (cstr_system.py:126,132; pfr_system.py:162-174)."""
        inlet_map: Dict[int, List[Tuple[int, int]]] = {}
        outlet_map: Dict[int, List[Tuple[int, int]]] = {}
        for k in np.nonzero(self.in_other < 0)[0].tolist():
            inlet_map.setdefault(int(self.in_other[k]), []).append((int(self.in_model[k]), int(self.in_conn[k])))
        for k in np.nonzero(self.out_other < 0)[0].tolist():
            outlet_map.setdefault(int(self.out_other[k]), []).append((int(self.out_model[k]), int(self.out_conn[k])))
        return inlet_map, outlet_map
