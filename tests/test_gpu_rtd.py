"""RTD kernel (occm_rtd) against the reference's golden rtd arrays and the oracle's moments (bar: 1e-6, north star)."""
import numpy as np
import pytest

import cases as C
from _util import load_golden, setup_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name', ['cstr_net9_tracer', 'pfr_chains_tracer'])
def test_rtd_from_reference_concentrations(name, tmp_path):
    """Feeding the reference's own concentrations isolates the RTD kernel: agreement to round-off."""
    from oracle import openccm_oracle as O
    from openccm_b200.postprocessing import network_to_rtd
    from openccm_b200.system_solvers.export import export_system
    g = load_golden(name)
    cp, cmesh, net, mn = setup_case(name, 'B200-RK45', tmp_path)
    host = export_system(C.CASES[name]['model'], mn, cp, cmesh)
    res = (g['c_RK45'], g['t_RK45'], host.inlet_map, host.outlet_map)
    rtd, mom = network_to_rtd(res, cmesh, cp, mn, return_moments=True)
    np.testing.assert_allclose(rtd, g['rtd_RK45'], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(mom, O.rtd_moments(g['rtd_RK45']), rtol=1e-10)


@pytest.mark.parametrize('name', ['cstr_net9_tracer', 'pfr_chains_tracer'])
def test_rtd_end_to_end(name, tmp_path):
    """Device solve + device RTD vs the reference's solve + RTD: moments to 1e-6."""
    from oracle import openccm_oracle as O
    from openccm_b200.postprocessing import network_to_rtd
    from openccm_b200.system_solvers import solve_system
    g = load_golden(name)
    cp, cmesh, net, mn = setup_case(name, 'B200-RK45', tmp_path)
    res = solve_system(C.CASES[name]['model'], mn, cp, cmesh)
    rtd, mom = network_to_rtd(res, cmesh, cp, mn, return_moments=True)
    # (both cases sit at steady state for most of the interval, where RK45's error estimate is round-off and scipy
    #  itself moves by ~1e-8 under 1e-16 perturbations of the RHS, see tests/test_gpu_rk45.py)
    np.testing.assert_allclose(rtd[:, :, 2], g['rtd_RK45'][:, :, 2], rtol=1e-6, atol=5e-7)
    ref_mom = O.rtd_moments(g['rtd_RK45'])
    np.testing.assert_allclose(mom[:, :2], ref_mom[:, :2], rtol=1e-6)          # m0, mean residence time
    np.testing.assert_allclose(mom[:, 2], ref_mom[:, 2], rtol=2e-5)            # variance amplifies the tail noise


def test_rtd_openfoam_golden_concentrations(tmp_path):
    """C2: RTD of the real 444-PFR network from the reference's LSODA concentrations."""
    from openccm_b200.postprocessing import network_to_rtd
    from openccm_b200.system_solvers.export import export_system
    g = load_golden('openfoam_2d_pfr')
    cp, cmesh, net, mn = setup_case('openfoam_2d_pfr', 'B200-RK45', tmp_path)
    host = export_system('pfr', mn, cp, cmesh)
    rtd = network_to_rtd((g['c_LSODA'], g['t_LSODA'], host.inlet_map, host.outlet_map), cmesh, cp, mn)
    np.testing.assert_allclose(rtd, g['rtd_LSODA'], rtol=1e-11, atol=1e-13)
