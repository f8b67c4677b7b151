"""Device BDF (Newton + block-Jacobi GMRES) against the reference's golden solves.

Two adaptive integrators agree only to about their tolerance, so (SURVEY.md §7 "hard parts") the device BDF is run at
tolerances tighter than the acceptance band and compared with the tight LSODA goldens of the unmodified reference:
band rtol 1e-6 / atol 1e-10 is asserted wherever the golden itself is that accurate (its own rtol is 1e-8 / 1e-10)."""
import numpy as np
import pytest

import cases as C
from _util import load_golden, setup_case

pytestmark = pytest.mark.gpu

CASES = ['cstr_irreversible', 'cstr_monod', 'cstr_net12_nacl', 'cstr_net20_timebc', 'pfr_single', 'pfr_net8', 'pfr_net8_nacl',
         'cstr_net10_pointmass']


@pytest.mark.parametrize('name', CASES)
def test_bdf_matches_reference(name, tmp_path):
    """Device BDF at rtol 1e-9 vs the oracle (restated ddt + scipy LSODA) at rtol 1e-11: north-star band at every output."""
    from oracle import openccm_oracle as O
    from openccm_b200.system_solvers import solve_system
    g = load_golden(name)
    cp, cmesh, net, mn = setup_case(name, 'B200-BDF', tmp_path, rtol=1e-9, atol=1e-12)
    c, t, imap, omap, info = solve_system(C.CASES[name]['model'], mn, cp, cmesh, return_info=True)
    assert info['status'] == 1, info
    cpo, cmesho, _, mno = setup_case(name, 'LSODA', tmp_path, rtol=1e-11, atol=1e-13)
    ref, t_ref, *_ = O.solve_system(C.CASES[name]['model'], mno, cpo, cmesho)
    np.testing.assert_allclose(t, t_ref, rtol=0, atol=1e-14)
    assert np.all(np.abs(c - ref) <= 1e-10 + 1e-6 * np.abs(ref)), np.abs(c - ref).max()
    # the reference's own (looser, rtol 1e-8) golden solve agrees within its tolerance
    assert np.all(np.abs(c - g['c_LSODA']) <= 5e-7 + 1e-6 * np.abs(g['c_LSODA']))


def test_bdf_openfoam_network(tmp_path):
    """C2: the stiff 444-PFR network (residence times 8e-4 .. 4e3) — where scipy RK45 needs 8e5 RHS calls."""
    from openccm_b200.system_solvers import solve_system
    g = load_golden('openfoam_2d_pfr')
    cp, cmesh, net, mn = setup_case('openfoam_2d_pfr', 'B200-BDF', tmp_path, rtol=1e-9, atol=1e-12)
    c, t, imap, omap, info = solve_system('pfr', mn, cp, cmesh, return_info=True)
    assert info['status'] == 1, info
    ref = g['c_LSODA']                       # LSODA at rtol 1e-10 / atol 1e-12
    assert np.all(np.abs(c - ref) <= 1e-8 + 1e-6 * np.abs(ref)), np.abs(c - ref).max()
    assert info['nfev'] < 400_000


def test_bdf_openfoam_cstr_nacl(tmp_path):
    """C3 (BASELINE.json configs[2]): 76 CSTRs of the OpenFOAM example, 2NaCl + CaCO3 <-> Na2CO3 + CaCl2, against the
    unmodified reference's LSODA solve at rtol 1e-10 / atol 1e-12."""
    from openccm_b200.system_solvers import solve_system
    g = load_golden('openfoam_2d_cstr_nacl')
    cp, cmesh, net, mn = setup_case('openfoam_2d_cstr_nacl', 'B200-BDF', tmp_path, rtol=1e-9, atol=1e-12)
    c, t, imap, omap, info = solve_system('cstr', mn, cp, cmesh, return_info=True)
    assert info['status'] == 1, info
    ref = g['c_LSODA']
    np.testing.assert_allclose(t, g['t_LSODA'], rtol=0, atol=1e-14)
    assert np.all(np.abs(c - ref) <= 1e-8 + 1e-6 * np.abs(ref)), np.abs(c - ref).max()


def test_bdf_batched_members(tmp_path):
    """Members with different rate constants / flows advance in lock step with individual orders and step sizes."""
    from oracle import openccm_oracle as O
    from openccm_b200.ensemble import EnsembleSolver
    import scipy.integrate
    name = 'cstr_net12_nacl'
    cp, cmesh, net, mn = setup_case(name, 'B200-BDF', tmp_path, rtol=1e-9, atol=1e-12)
    solver = EnsembleSolver('cstr', mn, cp, cmesh)
    M = 4
    rng = np.random.default_rng(3)
    ks = 10.0 ** rng.uniform(-1, 1, (M, 2)); qs = rng.uniform(0.5, 2, M)
    res = solver.solve(M, k_scale=ks, q_scale=qs, save='all')
    assert np.all(res.status == 1)
    cpo, cmesho, _, mno = setup_case(name, 'RK45', tmp_path, rtol=1e-10, atol=1e-13)
    host = solver.host
    for m in range(M):
        netm = C.CASES[name]['net'](); netm.flows = netm.flows * float(qs[m])
        syso = O.build_cstr(netm.to_model_network(), cpo, cmesho)
        rx = O.make_reactions_scaled(cpo, host.species, ks[m])
        A, bcs = syso.extra['A'], syso.extra['bcs']
        S, N = syso.c_shape
        def fun(t, y):
            cc = y.reshape(S, N); d = cc @ A; bcs(t, d); rx(cc, d); return d.ravel()
        r = scipy.integrate.solve_ivp(fun, (0, 30), syso.y0, method='LSODA', rtol=1e-10, atol=1e-13, t_eval=res.t)
        c_dev = res.y[m].reshape(len(res.t), N, S).transpose(2, 1, 0).reshape(S * N, -1)
        assert np.all(np.abs(c_dev - r.y) <= 5e-9 + 1e-6 * np.abs(r.y)), (m, np.abs(c_dev - r.y).max())


@pytest.mark.parametrize('name', ['cstr_net12_nacl', 'pfr_net8_nacl'])
def test_bdf_stored_and_on_the_fly_preconditioner_agree(name, tmp_path, monkeypatch):
    """precond_mode 0 (block inverses kept in HBM, rebuilt lazily) and 1 (rebuilt on every application) solve the same
    Newton systems: the solutions agree far inside the tolerance band, and the stored variant rebuilds rarely."""
    import torch
    from openccm_b200.bdf import solve_bdf
    from openccm_b200.device import DeviceSystem
    from openccm_b200.system_solvers.export import export_system
    cp, cmesh, net, mn = setup_case(name, 'B200-BDF', tmp_path, rtol=1e-8, atol=1e-11)
    host = export_system(C.CASES[name]['model'], mn, cp, cmesh)
    dev = DeviceSystem(host)
    y0 = torch.as_tensor(host.to_device_layout(host.c0.reshape(-1)), device='cuda').reshape(1, -1)
    te = np.linspace(0.0, 10.0, 11)
    res = [solve_bdf(dev, y0, (0.0, 10.0), rtol=1e-8, atol=1e-11, first_step=1e-4, t_eval=te, precond_mode=mode) for mode in (0, 1)]
    assert all(np.all(r.status == 1) for r in res)
    a, b = res[0].y.cpu().numpy(), res[1].y.cpu().numpy()
    assert np.all(np.abs(a - b) <= 1e-10 + 1e-6 * np.abs(b)), np.abs(a - b).max()
    assert res[1].n_precond_factor.sum() == 0
    assert 0 < res[0].n_precond_factor[0] < res[0].n_accepted[0]
