"""Ensemble path: every member of a batched device solve equals the oracle run on that member's modified inputs
(rate constants, flows, boundary values and initial state scaled on the *input files*, not inside the solver)."""
import configparser
import os

import numpy as np
import pytest

import cases as C
from _util import setup_case
from openccm_b200.config import ConfigParser, GroupedBCs

pytestmark = pytest.mark.gpu


def _scaled_inputs(name, tmp_path, m, k_row, q, bc, ic):
    """CONFIG + reactions file + network of member m with the scales applied textually."""
    case = C.CASES[name]
    p = configparser.ConfigParser(); p.optionxform = str; p.read_string(C.RXN[case['rxn']])
    lines = ['[REACTIONS]'] + [f'{k}: {v}' for k, v in p['REACTIONS'].items()] + ['', '[RATE CONSTANTS]']
    for r, (k, v) in enumerate(p['RATE CONSTANTS'].items()):
        lines.append(f'{k}: {float(v) * float(k_row[r])!r}')
    rxn_path = os.path.join(str(tmp_path), f'rxn_{name}_{m}')
    with open(rxn_path, 'w') as fh:
        fh.write('\n'.join(lines) + '\n')
    cp = ConfigParser(text=C.config_text(name, 'RK45', rxn_path))
    pad = "\n" + " " * 22
    ics = []
    for line in case['ic']:
        sp_, val = [v.strip() for v in line.split('->')]
        ics.append(f'{sp_} -> {float(val) * float(ic)!r}')
    bcs = []
    for line in case['bc']:
        sp_, rest = [v.strip() for v in line.split(':')]
        bname, val = [v.strip() for v in rest.split('->')]
        bcs.append(f'{sp_}: {bname} -> {float(val) * float(bc)!r}')
    cp['SIMULATION']['initial_conditions'] = pad.join(ics)
    cp['SIMULATION']['boundary_conditions'] = pad.join(bcs)
    net = case['net']()
    net.flows = net.flows * float(q)
    return cp, C.make_cmesh(name, net, GroupedBCs(cp)), net.to_model_network()


@pytest.mark.parametrize('name', ['cstr_second_order', 'cstr_net12_nacl', 'pfr_net8_nacl'])
def test_ensemble_members_match_oracle(name, tmp_path):
    from oracle import openccm_oracle as O
    from openccm_b200.ensemble import EnsembleSolver
    M = 6
    rng = np.random.default_rng(42)
    cp, cmesh, net, mn = setup_case(name, 'B200-RK45', tmp_path)
    if name != 'cstr_second_order':     # give the all-zero initial state something to scale
        species = cp.get_list(['SIMULATION', 'specie_names'], str)
        C.CASES[name] = dict(C.CASES[name], ic=[f'{s} -> {0.2 * (i + 1)}' for i, s in enumerate(species)])
        cp, cmesh, net, mn = setup_case(name, 'B200-RK45', tmp_path)
    try:
        solver = EnsembleSolver(C.CASES[name]['model'], mn, cp, cmesh)
        R = solver.host.tables.n_rxn
        ks = 10.0 ** rng.uniform(-0.5, 0.5, (M, R)); qs = rng.uniform(0.5, 2.0, M)
        bs = rng.uniform(0.5, 1.5, M); ics = rng.uniform(0.1, 1.0, M)
        res = solver.solve(M, k_scale=ks, q_scale=qs, bc_scale=bs, ic_scale=ics, save='all', batch=4)
        assert np.all(res.status == 1)
        host = solver.host
        for m in range(M):
            cpm, cmeshm, mnm = _scaled_inputs(name, tmp_path, m, ks[m], qs[m], bs[m], ics[m])
            c_ref, t_ref, *_ = O.solve_system(C.CASES[name]['model'], mnm, cpm, cmeshm)
            c_dev = res.y[m].reshape(len(res.t), host.n_points, host.S).transpose(2, 1, 0)
            np.testing.assert_allclose(res.t, t_ref, rtol=0, atol=1e-14)
            assert np.all(np.abs(c_dev - c_ref) <= 1e-10 + 1e-6 * np.abs(c_ref)), (name, m, np.abs(c_dev - c_ref).max())
    finally:
        import importlib
        importlib.reload(C)


def test_ensemble_rtd_moments(tmp_path):
    """Inflow sweep on a tracer network: device RTD moments per member vs oracle (1e-6)."""
    from oracle import openccm_oracle as O
    from openccm_b200.ensemble import EnsembleSolver
    name = 'cstr_net9_tracer'
    cp, cmesh, net, mn = setup_case(name, 'B200-RK45', tmp_path)
    solver = EnsembleSolver('cstr', mn, cp, cmesh)
    qs = np.array([0.5, 1.0, 1.7])
    res = solver.solve(3, q_scale=qs, save='outlets', rtd=True)
    for m, q in enumerate(qs):
        netm = C.CASES[name]['net'](); netm.flows = netm.flows * q
        cpo, cmesho, _, _ = setup_case(name, 'RK45', tmp_path)
        mno = netm.to_model_network()
        r = O.solve_system('cstr', mno, cpo, cmesho)
        rtd_ref = O.network_to_rtd(r, cmesho, cpo, mno)
        np.testing.assert_allclose(res.rtd[m][:, :, 2], rtd_ref[:, :, 2], rtol=1e-6, atol=2e-7)
        np.testing.assert_allclose(res.moments[m], O.rtd_moments(rtd_ref), rtol=1e-6)
    assert res.moment_stats['mean'].shape == (1, 3)
