"""`solver = B200-LSODA`: automatic method selection in the spirit of the reference's default solver (LSODA,
expanded_config_parser.py:55): an explicit RK45 probe, then the device BDF from the probe's state when the step size is
stability limited.  No step-sequence parity is claimed for this name (ODEPACK's Adams / BDF switching is a different algorithm);
what is asserted is the acceptance band around the tight solutions of the unmodified reference."""
import numpy as np
import pytest

import cases as C
from _util import load_golden, setup_case

pytestmark = pytest.mark.gpu


def band_units(c, tight):
    return np.abs(c - tight) / (1e-10 + 1e-6 * np.abs(tight))


def test_stiff_network_switches_to_bdf(tmp_path):
    """C2, the 444-PFR network (residence times 8e-4 .. 4e3): RK45 alone needs 134 861 steps; the probe sees it and hands over."""
    from openccm_b200.system_solvers import solve_system
    tg = load_golden('tight_openfoam_2d_pfr')
    cp, cmesh, net, mn = setup_case('openfoam_2d_pfr', 'B200-LSODA', tmp_path, rtol=1e-6, atol=1e-10)
    c, t, imap, omap, info = solve_system('pfr', mn, cp, cmesh, return_info=True)
    assert info['status'] == 1 and info['switched_at'] is not None and 0.0 < info['switched_at'] < 5.0
    np.testing.assert_allclose(t, tg['t'], rtol=0, atol=1e-14)
    e = band_units(c, tg['c_tight'])
    e_ref = band_units(tg['c_user_LSODA'], tg['c_tight'])
    print(f"B200-LSODA: switched at t = {info['switched_at']:.3g}, nfev {info['nfev']} (reference LSODA: {int(tg['nfev_user_LSODA'])}), "
          f"max {e.max():.2f} band units (reference: {e_ref.max():.2f})")
    assert e.max() <= max(1.0, e_ref.max()), e.max()
    assert info['nfev'] < int(tg['nfev_user_LSODA'])


@pytest.mark.parametrize('name', ['cstr_irreversible', 'cstr_net12_nacl'])
def test_nonstiff_case_stays_explicit(name, tmp_path):
    from openccm_b200.system_solvers import solve_system
    cp, cmesh, net, mn = setup_case(name, 'B200-LSODA', tmp_path)
    c, t, imap, omap, info = solve_system(C.CASES[name]['model'], mn, cp, cmesh, return_info=True)
    cp2, cmesh2, _, mn2 = setup_case(name, 'B200-RK45', tmp_path)
    c2, t2, *_ = solve_system(C.CASES[name]['model'], mn2, cp2, cmesh2)
    assert info['status'] == 1 and info['switched_at'] is None
    np.testing.assert_array_equal(c, c2)
