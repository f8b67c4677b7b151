"""sparse restatement of the flow-balancing LP (`tweak_final_flows`, helpers.py:268-438).

`tests/golden/flow_balance.npz` holds the flows the UNMODIFIED reference produced (`make_golden.py --flows`); the sparse
assembly hands HiGHS the same LP, so the results are identical."""
import os
from types import SimpleNamespace

import numpy as np
import pytest

from openccm_b200 import synthetic as syn
from openccm_b200.compartment_models import tweak_final_flows

HERE = os.path.dirname(os.path.abspath(__file__))
FLOW_CASES = ((8, 1), (40, 2), (300, 3), (1500, 4))


def flow_case(n, seed):
    net = syn.random_network(n, n_io=2, seed=seed)
    conn = net.to_model_network()[0]
    rng = np.random.default_rng(seed)
    flows = net.flows * (1 + 1e-3 * rng.standard_normal(net.flows.shape))
    ids_in = sorted({o for (i, _) in conn.values() for o in i.values() if o < 0})
    ids_out = sorted({o for (_, oo) in conn.values() for o in oo.values() if o < 0})
    return conn, flows, SimpleNamespace(domain_inlets=tuple(ids_in), domain_outlets=tuple(ids_out))


@pytest.mark.parametrize('n,seed', FLOW_CASES)
def test_flows_identical_to_reference(n, seed):
    g = np.load(os.path.join(HERE, 'golden', 'flow_balance.npz'))
    conn, flows, gb = flow_case(n, seed)
    before = flows.copy()
    tweak_final_flows(conn, flows, gb, 1e-9, verbose=False)
    np.testing.assert_array_equal(flows, g[f'flows_{n}_{seed}'])
    assert np.all(flows >= before * (1 - 1e-12))           # adjustments are non-negative


def test_large_network_stays_sparse():
    """2e4 models x 4e4 connections: the dense int8 matrix of the reference would be 0.8 GB; here the LP sees ~8e4 non-zeros."""
    import time
    conn, flows, gb = flow_case(20_000, 7)
    t0 = time.perf_counter()
    tweak_final_flows(conn, flows, gb, 1e-8, verbose=False)
    assert time.perf_counter() - t0 < 60
    total = np.zeros(len(conn))
    for m, (inl, outl) in conn.items():
        total[m] = sum(flows[c] for c in inl) - sum(flows[c] for c in outl)
    assert np.abs(total).max() < 1e-8


def test_unpaired_connection_is_reported_like_the_reference():
    conn, flows, gb = flow_case(8, 1)
    # connection 0 enters model 1 but no longer leaves model 0
    k = next(c for c, o in conn[0][1].items() if o >= 0)
    other = conn[0][1].pop(k)
    conn[0][1][len(flows)] = -9                             # keep an outlet on model 0 (a new domain outlet)
    flows = np.append(flows, 1.0)
    with pytest.raises(ValueError, match=f'Connection {k} is not balanced'):
        tweak_final_flows(conn, flows, gb, 1e-9, verbose=False)


def test_model_without_outlet_is_rejected():
    conn, flows, gb = flow_case(8, 1)
    for c in list(conn[3][1]):
        o = conn[3][1].pop(c)
        if o >= 0:
            conn[o][0].pop(c)
    with pytest.raises((AssertionError, ValueError)):
        tweak_final_flows(conn, flows, gb, 1e-9, verbose=False)
