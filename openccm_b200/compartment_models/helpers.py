"""
Flow balancing of the model network: sparse restatement of the LP of
``openccm.compartment_models.helpers.tweak_final_flows`` (helpers.py:268-438).

The reference assembles the mass-balance constraints as a DENSE int8 ``models x connections`` matrix (helpers.py:313 —
20 GB at 1e5 x 2e5) with Python loops, and checks rows / columns one at a time.  The matrix has at most two non-zeros
per column, so it is assembled here in COO form from flat arrays and handed to ``scipy.optimize.linprog`` (HiGHS) as CSR:
the LP — objective, bounds, ``integrality=2`` — is the reference's, so the adjustments are the same.  Host code: this is
the producer side of the hot path's input, not a kernel.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import scipy.sparse as sp
from scipy.optimize import linprog


def _flatten(connections: Dict[int, Tuple[Dict[int, int], Dict[int, int]]]):
    """(model row, connection, other side, sign) of every connection end; +1 = inlet of the model, -1 = outlet."""
    ids = list(connections.keys())
    rows, cols, other, sign = [], [], [], []
    for i, id_model in enumerate(ids):
        inlets, outlets = connections[id_model]
        n_in, n_out = len(inlets), len(outlets)
        rows.extend([i] * (n_in + n_out))
        cols.extend(inlets.keys()); cols.extend(outlets.keys())
        other.extend(inlets.values()); other.extend(outlets.values())
        sign.extend([1] * n_in); sign.extend([-1] * n_out)
    return ids, (np.asarray(rows, dtype=np.int64), np.asarray(cols, dtype=np.int64), np.asarray(other, dtype=np.int64),
                 np.asarray(sign, dtype=np.int8))


def tweak_final_flows(connections: Dict[int, Tuple[Dict[int, int], Dict[int, int]]], volumetric_flows: np.ndarray,
                      grouped_bcs, atol_opt: float, *, verbose: bool = True) -> None:
    """Same signature and in-place semantics as the reference: ``volumetric_flows`` is adjusted so that every model
    conserves mass; raises like the reference when the network is inconsistent or the LP fails."""
    say = print if verbose else (lambda *a, **k: None)
    v_min = min(volumetric_flows)
    volumetric_flows /= v_min                                     # helpers.py:303-304
    n_conn = volumetric_flows.shape[0]
    c = np.ones(volumetric_flows.shape)
    ids, (rows, cols, other, sign) = _flatten(connections)
    n_models = len(ids)
    # the reference indexes rows of `a` by model id (helpers.py:323) and rows of `b` by position (:365): identical for the
    # contiguous ids 0..n-1 the producers emit; anything else was an IndexError there
    if ids != list(range(n_models)):
        raise IndexError('model ids must be 0..n-1 in order (helpers.py:323 indexes the constraint matrix by id)')
    a = sp.coo_matrix((sign.astype(np.float64), (rows, cols)), shape=(n_models, n_conn)).tocsr()   # duplicates cannot occur: dict keys
    domain = np.zeros(n_conn, dtype=bool)
    domain[cols[other < 0]] = True
    # every connection is connected to something, interior connections pair an inlet with an outlet (helpers.py:330-351)
    touched = np.bincount(cols, minlength=n_conn)
    col_sum = np.bincount(cols, weights=sign, minlength=n_conn)
    bad = np.flatnonzero(~domain & (col_sum != 0))
    if bad.size:
        k = int(bad[0])
        ends = [f"component {ids[r]} {'inlet from' if s > 0 else 'outlet to'} {o}" for r, cc, o, s in zip(rows, cols, other, sign) if cc == k]
        raise ValueError(f"Connection {k} is not balanced in the assembled network: column_sum={int(col_sum[k])}, endpoints={ends}")
    assert np.all(touched > 0)
    # every model has an inlet and an outlet (helpers.py:354-357)
    assert np.all(np.bincount(rows[sign > 0], minlength=n_models) > 0)
    assert np.all(np.bincount(rows[sign < 0], minlength=n_models) > 0)
    # b = -(sum of inlets) + (sum of outlets), accumulated in the reference's order: inlets first, then outlets (:361-372)
    b = np.zeros(n_models)
    np.subtract.at(b, rows[sign > 0], volumetric_flows[cols[sign > 0]])
    np.add.at(b, rows[sign < 0], volumetric_flows[cols[sign < 0]])

    say(f"Flow optimization problem size: {n_models} components, {n_conn} connections")
    say("BEFORE: MAX abs error: {:.4e}".format(np.max(abs(b)) * v_min))
    say("BEFORE: AVG abs error: {:.4e}".format(np.mean(abs(b)) * v_min))
    results = linprog(c, A_eq=a, b_eq=b, bounds=(0, None), integrality=2)          # helpers.py:385
    if not results["success"]:
        raise Exception("Flow optimization did not converge. \n" + "Message:" + results["message"])
    adjustments = results["x"]
    assert np.all(adjustments >= -1e-6 / v_min)
    b_new = np.abs(a @ adjustments - b)
    say("AFTER:  MAX abs error: {:.4e}".format(np.max(b_new) * v_min))
    say("AFTER:  AVG abs error: {:.4e}".format(np.mean(b_new) * v_min))
    volumetric_flows += adjustments
    volumetric_flows *= v_min

    # conservation of mass of the whole system and of every model (helpers.py:407-438)
    is_bc = other < 0
    inlet_bc = np.isin(other, np.asarray(grouped_bcs.domain_inlets, dtype=np.int64))
    outlet_bc = np.isin(other, np.asarray(grouped_bcs.domain_outlets, dtype=np.int64))
    q = volumetric_flows[cols]
    net_flow = q[is_bc & (sign > 0) & inlet_bc].sum() - q[is_bc & (sign < 0) & outlet_bc].sum()
    missing_flow = q[is_bc & (sign > 0) & ~inlet_bc].sum() - q[is_bc & (sign < 0) & ~outlet_bc].sum()
    assert np.isclose(net_flow, 0, rtol=0, atol=atol_opt)
    assert np.isclose(missing_flow, 0, rtol=0, atol=atol_opt)
    total = np.bincount(rows, weights=sign * q, minlength=n_models)
    worst = int(np.argmax(np.abs(total)))
    if not np.isclose(total[worst], 0, rtol=0, atol=atol_opt):
        print("Total flowrate around component {} is not balanced. Error {}".format(ids[worst], total[worst]))
        assert False
