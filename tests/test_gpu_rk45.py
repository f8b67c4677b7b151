"""GPU parity of the batched RK45 integrator (through the reference-shaped ``solve_system``) against the
reference's golden solves.  Acceptance band of the north star: rtol 1e-6 / atol 1e-10 at every output time; the
step-exact mirror is held to a much tighter band here (rtol 1e-9 / atol 1e-12)."""
import numpy as np
import pytest

import cases as C
from _util import load_golden, setup_case

pytestmark = pytest.mark.gpu

import os

from _util import GOLDEN_DIR

# Linear cases that sit at steady state for most of the interval: there the RK45 error estimate is pure round-off, the step
# sequence is governed by noise, and two runs of the same algorithm (scipy vs scipy with its RHS perturbed by 1e-16) differ by
# ~1e-8.  Comparing two such answers with each other needs a relaxed band; comparing EACH with a tight-tolerance solution of
# the unmodified reference (tests/golden/tight_*.npz: LSODA at rtol 1e-13, confirmed by DOP853) does not.
ROUNDOFF_LIMITED = {'cstr_net9_tracer', 'cstr_reversible', 'pfr_chains_tracer'}

RK_CASES = [n for n in sorted(C.CASES) if 'RK45' in C.CASES[n].get('solvers', ('RK45', 'LSODA'))]
TIGHT_RK_CASES = [n for n in RK_CASES if os.path.exists(os.path.join(GOLDEN_DIR, f'tight_{n}.npz'))]


def band_units(c, tight):
    """|c - tight| in units of the north-star acceptance band 1e-10 + 1e-6 |tight|."""
    return np.abs(c - tight) / (1e-10 + 1e-6 * np.abs(tight))


@pytest.mark.parametrize('name', RK_CASES)
def test_solve_system_matches_reference_rk45(name, tmp_path):
    from openccm_b200.system_solvers import solve_system
    g = load_golden(name)
    cp, cmesh, net, mn = setup_case(name, 'B200-RK45', tmp_path)
    c, t, imap, omap, info = solve_system(C.CASES[name]['model'], mn, cp, cmesh, return_info=True)
    assert info['status'] == 1
    assert c.shape == g['c_RK45'].shape
    np.testing.assert_allclose(t, g['t_RK45'], rtol=0, atol=1e-14)
    ref = g['c_RK45']
    if name in ROUNDOFF_LIMITED:
        # both the reference's solve and the device's sit inside the acceptance band around the tight solution, at every output
        tight = load_golden(f'tight_{name}')['c_tight']
        assert band_units(ref, tight).max() <= 1.0
        assert band_units(c, tight).max() <= 1.0, band_units(c, tight).max()
        assert abs(info['nfev'] - int(g['nfev_RK45'])) <= 0.05 * int(g['nfev_RK45'])      # step sequence governed by round-off
    else:
        # north-star acceptance band at every output time, and the much tighter band a step-exact mirror reaches
        assert np.all(np.abs(c - ref) <= 1e-10 + 1e-6 * np.abs(ref))
        np.testing.assert_allclose(c, ref, rtol=1e-9, atol=2e-12)
        assert abs(info['nfev'] - int(g['nfev_RK45'])) <= 12


@pytest.mark.parametrize('name', TIGHT_RK_CASES)
def test_rk45_at_the_user_tolerance_is_in_band_wherever_scipy_is(name, tmp_path):
    """rtol 1e-6 / atol 1e-10 (the north star's numbers).  At that tolerance the reference's own RK45 is up to 250 bands away
    from the tight solution on some cases (tests/golden/make_golden.py --tight prints them), so the assertion is: the device
    is inside the band wherever scipy is, never further out than scipy, and spends the same number of RHS calls."""
    from openccm_b200.system_solvers import solve_system
    tg = load_golden(f'tight_{name}')
    cp, cmesh, net, mn = setup_case(name, 'B200-RK45', tmp_path, rtol=1e-6, atol=1e-10)
    c, t, imap, omap, info = solve_system(C.CASES[name]['model'], mn, cp, cmesh, return_info=True)
    assert info['status'] == 1
    np.testing.assert_allclose(t, tg['t'], rtol=0, atol=1e-14)
    e_dev, e_ref = band_units(c, tg['c_tight']), band_units(tg['c_user_RK45'], tg['c_tight'])
    if name in ROUNDOFF_LIMITED:
        # accepted steps are decided by round-off in the error estimate: WHICH outputs land inside the band differs between two
        # runs of the same algorithm, so the statement is about the distribution — as many outputs in band as scipy, none further out
        assert (e_dev <= 1.0).mean() >= (e_ref <= 1.0).mean() - 0.02, ((e_dev <= 1.0).mean(), (e_ref <= 1.0).mean())
        # the largest deviation is one output of a run whose accept / reject decisions flip with the last bit of the error norm:
        # pfr_chains_tracer gives 6.4 bands with scipy, 6.7 with the generic row walk (131 attempts, 3 rejected) and 7.2 with the
        # arithmetic of the batched stage kernels, which the persistent kernel's transport path reproduces bit for bit (132
        # attempts, 4 rejected) — three correct evaluations of the same right-hand side at rtol 1e-6
        slack_max = 0.15
        slack = 0.05
    else:
        assert np.all(e_dev[e_ref <= 0.98] <= 1.0), e_dev[e_ref <= 0.98].max()
        slack_max = slack = 0.02
    assert e_dev.max() <= (1 + slack_max) * e_ref.max() + 0.02, (e_dev.max(), e_ref.max())
    assert abs(info['nfev'] - int(tg['nfev_user_RK45'])) <= slack * int(tg['nfev_user_RK45'])


@pytest.mark.parametrize('name', ['cstr_irreversible', 'cstr_net12_nacl', 'pfr_net8', 'cstr_net20_timebc'])
def test_all_steps_mode_reproduces_scipy_step_sequence(name, tmp_path):
    from openccm_b200.system_solvers import solve_system
    g = load_golden(name)
    cp, cmesh, net, mn = setup_case(name, 'B200-RK45', tmp_path, t_eval='all')
    c, t, *_ = solve_system(C.CASES[name]['model'], mn, cp, cmesh)
    assert t.shape == g['t_RK45_all'].shape
    # same number of accepted steps; step times agree up to the round-off sensitivity of the error estimate
    # (a difference of small numbers at rtol 1e-8), the states are compared at (slightly) shifted times
    np.testing.assert_allclose(t, g['t_RK45_all'], rtol=1e-5)
    np.testing.assert_allclose(c, g['c_RK45_all'], rtol=2e-5, atol=1e-8)
