"""Device BDF (Newton + preconditioned GMRES) against tight solutions of the unmodified reference.

Two adaptive integrators agree only to about their tolerance (SURVEY.md §7 "hard parts"), so the yard-stick is a tight
solution (tests/golden/tight_*.npz: the reference with LSODA at rtol 1e-13 / atol 1e-15, confirmed by DOP853 / Radau) and the
device runs at the USER tolerance of the north star, rtol 1e-6 / atol 1e-10, through `solver = B200-BDF` — whose documented
tolerance-mapping policy (openccm_b200.bdf.device_tolerances) is what makes it meet the band.  For context the goldens also
hold the reference's own LSODA solve at the same user tolerance; the tests print how it fares."""
import numpy as np
import pytest

import cases as C
from _util import load_golden, setup_case

pytestmark = pytest.mark.gpu

CASES = ['cstr_irreversible', 'cstr_monod', 'cstr_net12_nacl', 'cstr_net20_timebc', 'pfr_single', 'pfr_net8', 'pfr_net8_nacl',
         'cstr_net10_pointmass', 'pfr_net6_pointmass', 'cstr_net9_tracer', 'pfr_chains_tracer']


def band_units(c, tight):
    return np.abs(c - tight) / (1e-10 + 1e-6 * np.abs(tight))


def solve_at_user_tolerance(name, tmp_path):
    from openccm_b200.system_solvers import solve_system
    cp, cmesh, net, mn = setup_case(name, 'B200-BDF', tmp_path, rtol=1e-6, atol=1e-10)
    c, t, imap, omap, info = solve_system(C.CASES[name]['model'], mn, cp, cmesh, return_info=True)
    assert info['status'] == 1, info
    tg = load_golden(f'tight_{name}')
    np.testing.assert_allclose(t, tg['t'], rtol=0, atol=1e-14)
    e_dev, e_ref = band_units(c, tg['c_tight']), band_units(tg['c_user_LSODA'], tg['c_tight'])
    print(f'{name}: device max {e_dev.max():.3f} band units (nfev {info["nfev"]}); reference LSODA at the same tolerance: '
          f'max {e_ref.max():.2f}, {100 * (e_ref <= 1).mean():.1f} % of the outputs in band (nfev {int(tg["nfev_user_LSODA"])})')
    return e_dev, e_ref, info


@pytest.mark.parametrize('name', CASES)
def test_bdf_at_the_user_tolerance_is_inside_the_band(name, tmp_path):
    """North-star band 1e-10 + 1e-6 |c| around the tight solution, at every output time of every DOF — not only where the
    reference's own solver at that tolerance happens to be inside it."""
    e_dev, e_ref, info = solve_at_user_tolerance(name, tmp_path)
    assert e_dev.max() <= 1.0, e_dev.max()


def test_bdf_openfoam_network(tmp_path):
    """C2: the stiff 444-PFR network (residence times 8e-4 .. 4e3) — where scipy RK45 needs 8e5 RHS calls and the
    reference's LSODA 37 534 (36 408 of them for dense finite-difference Jacobians)."""
    e_dev, e_ref, info = solve_at_user_tolerance('openfoam_2d_pfr', tmp_path)
    assert e_dev.max() <= 1.0, e_dev.max()
    assert info['nfev'] < 40_000


def test_bdf_openfoam_cstr_nacl(tmp_path):
    """C3 (BASELINE.json configs[2]): 76 CSTRs of the OpenFOAM example, 2NaCl + CaCO3 <-> Na2CO3 + CaCl2."""
    e_dev, e_ref, info = solve_at_user_tolerance('openfoam_2d_cstr_nacl', tmp_path)
    assert e_dev.max() <= 1.0, e_dev.max()


def test_bdf_batched_members(tmp_path):
    """Members with different rate constants / flows advance in lock step with individual orders and step sizes."""
    from oracle import openccm_oracle as O
    from openccm_b200.ensemble import EnsembleSolver
    import scipy.integrate
    name = 'cstr_net12_nacl'
    cp, cmesh, net, mn = setup_case(name, 'B200-BDF', tmp_path, rtol=1e-7, atol=1e-10)      # the device integrates at 1e-9 / 1e-12
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
