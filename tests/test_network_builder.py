"""f-2: vectorised CSTR network construction (openccm_b200/compartment_models/cstr.py) against the UNMODIFIED reference's
`create_cstr_network` on the OpenFOAM 2-D example (tests/golden/openfoam_2d_cstr_builder.npz, made by
tests/golden/make_golden.py --cstr-builder: flattened inputs + the reference's network).  CPU only."""
import os
from types import SimpleNamespace

import numpy as np
import pytest

from _util import load_golden


class _Cfg:
    def __init__(self, g):
        self.v = {('COMPARTMENT MODELLING', 'flow_threshold'): float(g['flow_threshold']), ('COMPARTMENT MODELLING', 'atol_opt'): float(g['atol_opt'])}

    def get_item(self, path, typ):
        return typ(self.v[tuple(path)])


def _inputs(g):
    """Rebuild the reference-shaped inputs (dict of dict of dict, dict of sets, mesh attributes) from the flattened fixture."""
    cn, pos = {}, 0
    for c, nb, n in zip(g['g_comp'].tolist(), g['g_nb'].tolist(), g['g_len'].tolist()):
        cn.setdefault(c, {})[nb] = {int(f): int(e) for f, e in zip(g['facets'][pos:pos + n], g['elems'][pos:pos + n])}
        pos += n
    comp = {c: set(g['comp_el'][a:b].tolist()) for c, (a, b) in enumerate(zip(g['comp_ptr'][:-1], g['comp_ptr'][1:]))}
    n_facets = int(g['facets'].max()) + 1
    fu = np.zeros((n_facets, 2), dtype=object)
    fu[:, 1] = 0
    fu[g['facets'], 0] = g['facet_flow']; fu[g['facets'], 1] = g['facet_flag']
    facet_elements = {int(f): tuple(int(x) for x in pair if x >= 0) for f, pair in zip(g['facets'], g['facet_el'])}
    mesh = SimpleNamespace(facet_elements=facet_elements, element_sizes=g['element_sizes'],
                           grouped_bcs=SimpleNamespace(domain_inlets=tuple(g['domain_inlets'].tolist()), domain_outlets=tuple(g['domain_outlets'].tolist())))
    return comp, cn, mesh, fu


def test_cstr_network_matches_reference():
    from openccm_b200.compartment_models import create_cstr_network, create_cstr_network_arrays
    g = load_golden('openfoam_2d_cstr_builder')
    comp, cn, mesh, fu = _inputs(g)
    net = create_cstr_network_arrays(comp, cn, mesh, fu, _Cfg(g))
    # the 5-tuple's first three members: identical ids, connection order, flows; volumes to summation order of a rebuilt set
    for key in ('in_model', 'in_conn', 'in_other', 'out_model', 'out_conn', 'out_other'):
        np.testing.assert_array_equal(getattr(net, key), g[key], err_msg=key)
    np.testing.assert_array_equal(net.flows, g['flows'])                    # after the flow-balancing LP, bit-identical
    np.testing.assert_allclose(net.volumes, g['volumes'], rtol=1e-14, atol=0)
    connections, volumes, flows, c2m, m2e = create_cstr_network(comp, cn, mesh, fu, None, _Cfg(g))
    assert c2m == {i: [i] for i in range(len(volumes))}
    assert [sorted(e for _, e in row) for row in m2e] == [sorted(comp[c]) for c in sorted(comp)]
    # every CSTR conserves mass after balancing (helpers.py:426-438)
    bal = np.bincount(net.in_model, weights=net.flows[net.in_conn], minlength=net.n_models) - \
        np.bincount(net.out_model, weights=net.flows[net.out_conn], minlength=net.n_models)
    assert np.abs(bal).max() < 1e-15


def test_connect_cstr_compartments_threshold_and_signs():
    """Three compartments in a row with a boundary inlet and outlet; the middle interface is below the flow threshold from one
    side's point of view only if its net flow is — checks first-encounter numbering, sign convention and rejection."""
    from openccm_b200.compartment_models import connect_cstr_compartments
    # facets: 0 inlet (boundary, flag -1), 1 between comp 0 / 1 (upwind element 0), 2 between 1 / 2 (upwind element 1),
    # 3 outlet (boundary, upwind = the single element), 4 a second facet between 1 / 2 whose tiny back-flow is kept in the net sum
    facet_elements = {0: (0,), 1: (0, 1), 2: (1, 2), 3: (2,), 4: (1, 2)}
    fu = np.array([[1.0, -1], [1.0, 0], [1.25, 0], [1.0, 0], [0.25, 1]], dtype=object)
    cn = {0: {-1: {0: 0}, 1: {1: 0}}, 1: {0: {1: 1}, 2: {2: 1, 4: 1}}, 2: {1: {2: 2, 4: 2}, -2: {3: 2}}}
    mesh = SimpleNamespace(facet_elements=facet_elements)
    cfg = SimpleNamespace(get_item=lambda path, typ: typ(1e-15))
    pairing, flows = connect_cstr_compartments(cn, mesh, fu, 2, cfg)
    assert pairing == {0: {1: -1, -2: 1}, 1: {2: 0, -3: 2}, 2: {3: 1, -4: -2}}
    assert flows == {1: 1.0, 2: 1.0, 3: 1.0, 4: 1.0}
