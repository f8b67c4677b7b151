"""
Pins the CPU oracle (oracle/openccm_oracle.py) against outputs of the UNMODIFIED reference
(tests/golden/*.npz, made by tests/golden/make_golden.py in the build container) and against the
reference's analytic known answers.  CPU only.
"""
import json
import os

import numpy as np
import pytest

import cases as C
from _util import GOLDEN_DIR, load_golden, setup_case
from oracle import openccm_oracle as O

ALL = sorted(C.CASES)
FAST = [n for n in ALL if n != 'openfoam_2d_pfr']


def test_parser_kats():
    """Rate strings of reactions.py:335-392 and their sympy-simplified forms."""
    import configparser
    import sympy as sp
    kats = json.load(open(os.path.join(GOLDEN_DIR, 'parser_kats.json')))
    for key, text in C.RXN.items():
        p = configparser.ConfigParser(); p.optionxform = str; p.read_string(text)
        raw = O.parse_reactions(dict(p['REACTIONS']), dict(p['RATE CONSTANTS']))
        assert raw == kats[key]['raw'], key
        assert {k: str(sp.parse_expr(v)) for k, v in raw.items()} == kats[key]['sympy'], key


@pytest.mark.parametrize('name', ALL)
def test_c0_and_raw_ddt(name, tmp_path):
    """c0 (ICs + folded BCs) and the raw right-hand side at random states vs the reference's numba ddt."""
    g = load_golden(name)
    cp, cmesh, net, mn = setup_case(name, 'RK45', tmp_path)
    s = O.build_system(C.CASES[name]['model'], mn, cp, cmesh)
    np.testing.assert_allclose(s.y0, g['c0'], rtol=1e-14, atol=1e-15)
    for y, t, ref in zip(g['ddt_y'], g['ddt_t'], g['ddt_out']):
        out = s.fun(float(t), y.copy())
        np.testing.assert_allclose(out, ref, rtol=1e-12, atol=1e-13 * max(1.0, np.abs(ref).max()))


@pytest.mark.parametrize('name', FAST)
def test_rk45_solve(name, tmp_path):
    """Full RK45 solve: same scipy integrator around the restated ddt -> agreement to round-off."""
    g = load_golden(name)
    cp, cmesh, net, mn = setup_case(name, 'RK45', tmp_path)
    c, t, imap, omap, res = O.solve_system(C.CASES[name]['model'], mn, cp, cmesh, return_result=True)
    np.testing.assert_allclose(t, g['t_RK45'], rtol=0, atol=1e-14)
    np.testing.assert_allclose(c, g['c_RK45'], rtol=1e-9, atol=1e-12)
    assert abs(res.nfev - int(g['nfev_RK45'])) <= 12       # a round-off-flipped accept/reject costs 6 evaluations


@pytest.mark.parametrize('name', ['cstr_irreversible', 'cstr_net12_nacl', 'pfr_net8', 'cstr_net20_timebc'])
def test_rk45_all_steps(name, tmp_path):
    """t_eval = all: the solver's own step sequence is reproduced."""
    g = load_golden(name)
    cp, cmesh, net, mn = setup_case(name, 'RK45', tmp_path, t_eval='all')
    c, t, *_ = O.solve_system(C.CASES[name]['model'], mn, cp, cmesh)
    assert t.shape == g['t_RK45_all'].shape
    np.testing.assert_allclose(t, g['t_RK45_all'], rtol=1e-9)
    np.testing.assert_allclose(c, g['c_RK45_all'], rtol=1e-8, atol=1e-12)


@pytest.mark.parametrize('name', ['cstr_irreversible', 'cstr_net12_nacl', 'pfr_net8', 'pfr_chains_tracer'])
def test_lsoda_solve(name, tmp_path):
    g = load_golden(name)
    cp, cmesh, net, mn = setup_case(name, 'LSODA', tmp_path)
    c, t, *_ = O.solve_system(C.CASES[name]['model'], mn, cp, cmesh)
    np.testing.assert_allclose(c, g['c_LSODA'], rtol=1e-7, atol=1e-10)


@pytest.mark.parametrize('name', [n for n in ALL if C.CASES[n]['rtd'] and n != 'openfoam_2d_pfr'])
def test_rtd(name, tmp_path):
    g = load_golden(name)
    cp, cmesh, net, mn = setup_case(name, 'RK45', tmp_path)
    res = O.solve_system(C.CASES[name]['model'], mn, cp, cmesh)
    rtd = O.network_to_rtd(res, cmesh, cp, mn)
    np.testing.assert_allclose(rtd, g['rtd_RK45'], rtol=1e-8, atol=1e-11)
    # RTD computed by the oracle from the reference's own concentrations is bit-comparable
    rtd2 = O.network_to_rtd((g['c_RK45'], g['t_RK45'], res[2], res[3]), cmesh, cp, mn)
    np.testing.assert_allclose(rtd2, g['rtd_RK45'], rtol=1e-13, atol=1e-15)


def test_openfoam_2d_lsoda_and_rtd(tmp_path):
    """C2: the real 444-PFR network of examples/OpenFOAM/pipe_with_recirc (tracer step, RTD)."""
    g = load_golden('openfoam_2d_pfr')
    cp, cmesh, net, mn = setup_case('openfoam_2d_pfr', 'LSODA', tmp_path)
    res = O.solve_system('pfr', mn, cp, cmesh)
    np.testing.assert_allclose(res[0], g['c_LSODA'], rtol=1e-6, atol=1e-9)
    rtd = O.network_to_rtd(res, cmesh, cp, mn)
    np.testing.assert_allclose(rtd[:, :, 2], g['rtd_LSODA'][:, :, 2], rtol=1e-6, atol=1e-9)


def test_analytic_irreversible_cstr(tmp_path):
    """examples/simple_reactors/cstr/irreversible/cstr_analysis.py:40-53,60-90: mean |err| <= tol."""
    k1, k2, tau, Ca0 = 0.5, 0.2, 10.0, 2.0
    k1T = k1 * tau; bA = (k1T + 1) / tau; a1 = k1 * Ca0 / (k1T + 1); a2 = a1 * k1T; bB = (k2 * tau + 1) / tau
    Ca = lambda t: (Ca0 / (1 + k1T)) * (1 + k1T * np.exp(-t * bA))
    Cb = lambda t: (a1 / bB) * (1 - np.exp(-bB * t)) + (a2 / (bB - bA)) * (np.exp(-bA * t) - np.exp(-bB * t))
    for tol in (1e-3, 1e-7, 1e-10):
        cp, cmesh, net, mn = setup_case('cstr_irreversible', 'LSODA', tmp_path, rtol=tol, atol=tol, t_eval='all')
        c, t, *_ = O.solve_system('cstr', mn, cp, cmesh)
        assert np.mean(np.abs(c[0, 0] - Ca(t))) <= tol
        assert np.mean(np.abs(c[1, 0] - Cb(t))) <= tol


def test_paper_equilibrium(tmp_path):
    """publications/2403_JOSS/paper.md:160-171: 2NaCl+CaCO3 <-> Na2CO3+CaCl2 equilibrium x ~= 0.1147 satisfies
    k_f (1-2x)^2 (1-x) = k_r x^2; a closed (no-flow-limited) long-time CSTR outlet approaches the inlet-fed steady state,
    so only the algebraic identity of the mechanism is checked here through the oracle's reaction function."""
    cp, cmesh, net, mn = setup_case('cstr_net12_nacl', 'RK45', tmp_path)
    species = cp.get_list(['SIMULATION', 'specie_names'], str)
    rx = O.make_reactions(cp, species, None)
    x = 0.1147
    C_eq = np.array([[1 - 2 * x], [1 - x], [x], [x]])
    d = np.zeros_like(C_eq)
    rx(C_eq, d)
    assert np.all(np.abs(d) < 2e-4)


def test_sparse_cstr_assembly_is_pinned(tmp_path):
    """The oracle's sparse CSTR path (`build_cstr(sparse=True)`: the CPU arm of bench.py and the yard-stick of the full-size GPU
    tests, where the reference's dense N x N operator does not fit) against the UNMODIFIED reference's raw ddt at N = 2000 and
    against the oracle's own verbatim dense path."""
    g = load_golden('cstr_net2000')
    cp, cmesh, net, mn = setup_case('cstr_net2000', 'RK45', tmp_path)
    sparse = O.build_cstr(mn, cp, cmesh, sparse=True)
    dense = O.build_cstr(mn, cp, cmesh, sparse=False)
    import scipy.sparse
    assert scipy.sparse.issparse(sparse.extra['A']) and isinstance(dense.extra['A'], np.ndarray)
    np.testing.assert_allclose(sparse.extra['A'].toarray(), dense.extra['A'].T, rtol=1e-15, atol=0)      # A_sparse = Q_div_v^T
    np.testing.assert_allclose(sparse.y0, g['c0'], rtol=1e-14, atol=1e-15)
    for y, t, ref in zip(g['ddt_y'], g['ddt_t'], g['ddt_out']):
        out_s, out_d = sparse.fun(float(t), y.copy()), dense.fun(float(t), y.copy())
        scale = max(1.0, np.abs(ref).max())
        np.testing.assert_allclose(out_s, ref, rtol=1e-12, atol=1e-13 * scale)
        np.testing.assert_allclose(out_s, out_d, rtol=1e-13, atol=1e-13 * scale)
