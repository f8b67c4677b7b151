"""C-ABI behaviour that is not numerics: error reporting, pause/resume of the all-steps mode, step budget, per-point fields."""
import numpy as np
import pytest

import cases as C
from _util import load_golden, setup_case

pytestmark = pytest.mark.gpu


def _system(name, tmp_path, **kw):
    import torch
    from openccm_b200.device import DeviceSystem
    from openccm_b200.system_solvers.export import export_system
    cp, cmesh, net, mn = setup_case(name, 'B200-RK45', tmp_path, **kw)
    host = export_system(C.CASES[name]['model'], mn, cp, cmesh)
    dev = DeviceSystem(host)
    y0 = torch.as_tensor(host.to_device_layout(host.c0.reshape(-1)), device='cuda').reshape(1, -1)
    return host, dev, y0


def test_first_step_validation_matches_scipy_messages(tmp_path):
    from openccm_b200._lib import OccmError
    host, dev, y0 = _system('cstr_irreversible', tmp_path)
    with pytest.raises(OccmError, match='exceeds bounds'):
        dev.rk45(y0, (0.0, 1.0), rtol=1e-6, atol=1e-9, first_step=2.0, t_eval=[0.0, 1.0])
    with pytest.raises(OccmError, match='must be positive'):
        dev.rk45(y0, (0.0, 1.0), rtol=1e-6, atol=1e-9, first_step=0.0, t_eval=[0.0, 1.0])
    with pytest.raises(ValueError, match='not within'):
        dev.rk45(y0, (0.0, 1.0), rtol=1e-6, atol=1e-9, first_step=1e-3, t_eval=[0.0, 2.0])
    with pytest.raises(ValueError, match='sorted'):
        dev.rk45(y0, (0.0, 1.0), rtol=1e-6, atol=1e-9, first_step=1e-3, t_eval=[0.5, 0.2])


def test_all_steps_mode_pause_and_resume(tmp_path):
    """A tiny output buffer forces the pause / harvest / resume cycle; the step sequence must not change."""
    g = load_golden('cstr_net12_nacl')
    host, dev, y0 = _system('cstr_net12_nacl', tmp_path)
    res = dev.rk45(y0, (0.0, 30.0), rtol=1e-8, atol=1e-10, first_step=1e-4, t_eval=None, all_capacity=16)
    t = res.t[0]
    assert t.shape == g['t_RK45_all'].shape
    np.testing.assert_allclose(t, g['t_RK45_all'], rtol=1e-5)
    assert np.all(np.diff(t) > 0) and t[0] == 0.0 and t[-1] == 30.0


def test_step_budget_stops_members_and_returns_partial_rows(tmp_path):
    host, dev, y0 = _system('cstr_net12_nacl', tmp_path)
    res = dev.rk45(y0, (0.0, 30.0), rtol=1e-8, atol=1e-10, first_step=1e-4, t_eval=np.linspace(0, 30, 61), max_steps=20)
    assert res.status[0] == -3 and res.n_accepted[0] + res.n_rejected[0] == 20
    assert 0 < res.n_rows[0] < 61


def test_per_point_field_in_rate_expression(tmp_path):
    """[EXTRA TERMS] `kf: CFD, file` (reactions.py:216-231): a per-DOF array inside the rate expression."""
    import os
    import torch
    from oracle import openccm_oracle as O
    from openccm_b200.device import DeviceSystem
    from openccm_b200.system_solvers.export import export_system
    name = 'cstr_net12_nacl'
    cp, cmesh, net, mn = setup_case(name, 'B200-RK45', tmp_path)
    rxn = os.path.join(str(tmp_path), 'rxn_field')
    with open(rxn, 'w') as fh:
        fh.write("[REACTIONS]\nR1: 2NaCl + CaCO3 -> Na2CO3 + CaCl2\nR2: Na2CO3 + CaCl2 -> 2NaCl + CaCO3\n\n"
                 "[RATE CONSTANTS]\nR1: kf\nR2: 2\n\n[EXTRA TERMS]\nkf: CFD, kf_field\n")
    cp['SIMULATION']['reactions_file_path'] = rxn
    kf = np.random.default_rng(1).uniform(0.01, 0.2, net.n_models)
    host = export_system('cstr', mn, cp, cmesh, fields={'kf': kf})
    dev = DeviceSystem(host)
    species = cp.get_list(['SIMULATION', 'specie_names'], str)
    rx = O.make_reactions(cp, species, None, extra_arrays={'kf': kf})
    rng = np.random.default_rng(2)
    c = rng.uniform(0.1, 1.0, (host.S, host.n_points))
    ref = np.zeros_like(c)
    rx(c, ref)
    # transport + BC part from the oracle system built without reactions
    cp0, cmesh0, _, mn0 = setup_case(name, 'RK45', tmp_path)
    cp0['SIMULATION']['reactions_file_path'] = 'None'
    lin = O.build_cstr(mn0, cp0, cmesh0).fun(0.3, c.reshape(-1).copy()).reshape(c.shape)
    y = torch.as_tensor(host.to_device_layout(c.reshape(-1)), device='cuda').reshape(1, -1)
    out = host.to_reference_layout(dev.rhs(y, 0.3).cpu().numpy().reshape(-1)).reshape(c.shape)
    np.testing.assert_allclose(out, lin + ref, rtol=1e-12, atol=1e-13)
