"""The UNMODIFIED reference pipeline on the device: `openccm.run(CONFIG)` of examples/OpenFOAM/pipe_with_recirc with
`openccm_b200.dropin.install()` and `solver = B200-BDF` (north star: "openccm runs unchanged with a device solver in place").

The reference is installed into baseline/_ref and the example's inputs are copied to baseline/_ref_examples by
scripts/install_reference.sh (build container only; both directories travel to the GPU box with the snapshot, neither is
committed).  Everything upstream of the solver is the reference's own code: ConfigParser, OpenFOAM ingest, CMesh,
compartmentalisation, network construction, the cache pickles, np.save and the call of network_to_rtd at run.py:205.
Outputs are compared with the tight solutions the unmodified reference produced for the same CONFIG (tests/golden/tight_*.npz).
"""
import os
import shutil
import sys

import numpy as np
import pytest

from _util import load_golden

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(REPO, 'baseline', '_ref')
EXAMPLE = os.path.join(REPO, 'baseline', '_ref_examples', 'pipe_with_recirc')


def _reference():
    if not (os.path.isdir(os.path.join(REF, 'openccm')) and os.path.isdir(EXAMPLE)):
        pytest.skip('reference not installed (scripts/install_reference.sh, build container)')
    try:
        import pyparsing  # noqa: F401
    except ModuleNotFoundError:                      # the reference imports it at module level (reactions.py:29)
        import pip._vendor.pyparsing as _pp
        sys.modules['pyparsing'] = _pp
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import openccm
    assert os.path.realpath(openccm.__file__).startswith(os.path.realpath(REF)), openccm.__file__
    return openccm


def band_units(c, tight):
    return np.abs(c - tight) / (1e-10 + 1e-6 * np.abs(tight))


@pytest.fixture()
def case_dir(tmp_path):
    dst = tmp_path / 'case'
    shutil.copytree(EXAMPLE, dst)
    old = os.getcwd()
    os.chdir(dst)
    yield dst
    os.chdir(old)


def _run(openccm, solver, **sim):
    """One `openccm.run` of the CONFIG in the current directory with the given [SIMULATION] overrides."""
    import scipy.integrate
    cp = openccm.ConfigParser('CONFIG')
    cp['SIMULATION']['solver'] = solver
    cp['SIMULATION'].update({k: str(v) for k, v in sim.items()})
    if solver.upper() not in ('B200-RK45', 'B200-BDF'):
        # harness-side work-around for the reference path only: reactions.py:214 sets np.seterr(all='raise')
        np.seterr(all='warn', under='ignore')
    timing = openccm.run(cp)
    model = cp.get_item(['COMPARTMENT MODELLING', 'model'], str).lower()
    out = cp.get_item(['SETUP', 'output_folder_path'], str)
    load = lambda name: np.load(os.path.join(out, f'{model}_{name}.npy'))
    return timing, load('concentrations'), load('t'), (load('rtd') if cp.get_item(['POST-PROCESSING', 'calculate_rtd'], bool) else None)


def test_openccm_run_pfr_tracer_and_rtd_on_device(case_dir):
    """C2: 444 PFRs x 2 points, tracer step, RTD — concentrations, times and the RTD array written by run() itself."""
    openccm = _reference()
    from openccm_b200 import dropin
    sim = dict(t_eval='5.0, linear', rtol=1e-6, atol=1e-10, first_timestep=1e-4)
    dropin.install()
    try:
        run_mod = sys.modules['openccm.run']
        assert run_mod.solve_system.__module__ == 'openccm_b200.dropin' and run_mod.network_to_rtd.__module__ == 'openccm_b200.dropin'
        timing, c, t, rtd = _run(openccm, 'B200-BDF', **sim)
    finally:
        dropin.uninstall()
    tg = load_golden('tight_openfoam_2d_pfr')
    assert c.shape == tg['c_tight'].shape                          # the reference built the same 444-PFR network on this box
    np.testing.assert_allclose(t, tg['t'], rtol=0, atol=1e-12)
    assert band_units(c, tg['c_tight']).max() <= 1.0, band_units(c, tg['c_tight']).max()
    F, F_ref = rtd[:, :, 2], tg['rtd_tight'][:, :, 2]
    assert np.all(np.abs(F - F_ref) <= 1e-10 + 1e-6 * np.abs(F_ref))
    np.testing.assert_allclose(rtd[:, :, 0], tg['rtd_tight'][:, :, 0], rtol=0, atol=1e-12)
    np.testing.assert_allclose(rtd[:, :, 1], tg['rtd_tight'][:, :, 1], rtol=1e-4, atol=1e-9)   # E = dF/dt (second-order differences)
    # the same CONFIG through the reference's own scipy path (its default solver), from the same cache pickles
    timing_ref, c_ref, t_ref, rtd_ref = _run(openccm, 'LSODA', **sim)
    e_dev, e_ref = band_units(c, tg['c_tight']).max(), band_units(c_ref, tg['c_tight']).max()
    dev_s, ref_s = timing['Solve model'] * 1e-9, timing_ref['Solve model'] * 1e-9
    print(f"\n[dropin] PFR tracer: 'Solve model' device {dev_s:.2f} s vs reference LSODA {ref_s:.2f} s ({ref_s / dev_s:.1f}x); "
          f"max deviation from the tight solution: device {e_dev:.2f}, reference {e_ref:.2f} band units; "
          f"'RTD Calcs' device {timing['RTD Calcs'] * 1e-9:.3f} s vs reference {timing_ref['RTD Calcs'] * 1e-9:.3f} s")
    assert e_dev <= e_ref
    assert dev_s < ref_s


def test_openccm_run_cstr_nacl_on_device(case_dir):
    """C3: the same mesh as a CSTR network (76 CSTRs) with the 2NaCl + CaCO3 <-> Na2CO3 + CaCl2 kinetics of
    examples/OpenCMP/pipe_with_recirc_2d/reactions."""
    openccm = _reference()
    from openccm_b200 import dropin
    cp_over = {'COMPARTMENT MODELLING': {'model': 'cstr'}, 'POST-PROCESSING': {'calculate_rtd': 'False'}}
    sim = dict(t_eval='5.0, linear', rtol=1e-6, atol=1e-10, first_timestep=1e-4, points_per_pfr=1,
               specie_names='NaCl, CaCO3, Na2CO3, CaCl2', reactions_file_path='reactions_nacl',
               initial_conditions='NaCl -> 0\nCaCO3 -> 0\nNa2CO3 -> 0\nCaCl2 -> 0',
               boundary_conditions='NaCl: inlet -> 1\nCaCO3: inlet -> 1\nNa2CO3: inlet -> 0\nCaCl2: inlet -> 0')
    # section overrides go through the file so that ConfigParser's own validation sees them
    text = open('CONFIG').read().replace('model = PFR', 'model = CSTR').replace('calculate_rtd          = True', 'calculate_rtd          = False')
    open('CONFIG', 'w').write(text)
    dropin.install()
    try:
        timing, c, t, _ = _run(openccm, 'B200-BDF', **sim)
    finally:
        dropin.uninstall()
    tg = load_golden('tight_openfoam_2d_cstr_nacl')
    assert c.shape == tg['c_tight'].shape
    np.testing.assert_allclose(t, tg['t'], rtol=0, atol=1e-12)
    assert band_units(c, tg['c_tight']).max() <= 1.0, band_units(c, tg['c_tight']).max()
    print(f"\n[dropin] CSTR NaCl: 'Solve model' device {timing['Solve model'] * 1e-9:.2f} s, max deviation {band_units(c, tg['c_tight']).max():.2f} band units")
