"""recorded solution -> mesh elements (export_to_vtu_openfoam's buffers).

Goldens (`tests/golden/vtu_*.npz`) are the buffers the UNMODIFIED reference handed to its file writer
(`make_golden.py --vtu`).  CPU: the oracle restatement and the vectorised interpolation table reproduce them bit for bit;
GPU: so does `occm_dofs_to_elements` through the host mirror."""
import os

import numpy as np
import pytest

import cases as C
from _util import load_golden

HERE = os.path.dirname(os.path.abspath(__file__))
VTU_CASES = ['pfr_net8', 'cstr_net20_timebc', 'pfr_single', 'openfoam_2d_pfr']


def _inputs(name):
    g = load_golden(name)
    v = np.load(os.path.join(HERE, 'golden', f'vtu_{name}.npz'))
    solver = 'RK45' if 'c_RK45' in g.files else 'LSODA'
    c = np.ascontiguousarray(g[f'c_{solver}'][:, :, v['time_index']])
    net = C.CASES[name]['net']()
    m2e = net.to_model_network()[-1]
    return c, m2e, C.CASES[name]['P'] if C.CASES[name]['model'] == 'pfr' else 1, v['fields']


@pytest.mark.parametrize('name', VTU_CASES)
def test_oracle_element_fields_match_reference(name):
    from oracle import openccm_oracle as O
    c, m2e, P, fields = _inputs(name)
    out = O.element_fields(c, m2e, P, fields.shape[2])
    np.testing.assert_array_equal(out, fields)


@pytest.mark.parametrize('name', VTU_CASES)
def test_vectorised_interpolation_table_matches_reference_loop(name):
    """export.element_interpolation == create_dof_to_element_map (restated in the oracle), element by element."""
    from oracle import openccm_oracle as O
    from openccm_b200.system_solvers.export import element_interpolation
    c, m2e, P, fields = _inputs(name)
    n_el = fields.shape[2]
    el_dof, el_other, el_w = element_interpolation(m2e, len(m2e), P, n_el)
    ref_dof = -np.ones(n_el, dtype=np.int64); ref_other = np.zeros(n_el, dtype=np.int64); ref_w = np.zeros(n_el)
    for dof, mapping in enumerate(O.create_dof_to_element_map(m2e, P)):
        for element, other, w in mapping:
            ref_dof[element] = dof; ref_other[element] = other % (len(m2e) * P); ref_w[element] = w
    np.testing.assert_array_equal(el_dof, ref_dof)
    mapped = ref_dof >= 0
    np.testing.assert_array_equal(el_other[mapped], ref_other[mapped])
    np.testing.assert_array_equal(el_w[mapped], ref_w[mapped])
    # and the numpy evaluation of the table gives the reference's buffers
    S, n_points, T = c.shape
    val = c[:, el_dof.clip(0), :] * el_w[None, :, None] + c[:, el_other, :] * (1 - el_w)[None, :, None]
    val = np.where(mapped[None, :, None], val, -1.0).transpose(2, 0, 1)
    np.testing.assert_array_equal(val, fields)


def test_interpolation_table_edge_cases():
    """Empty models, a model whose elements all sit at the inlet, unsorted ties: same table as the reference's walk."""
    from oracle import openccm_oracle as O
    from openccm_b200.system_solvers.export import element_interpolation
    m2e = [[], [(0.0, 4), (0.0, 1)], [(0.1, 0), (0.5, 2), (0.5, 3), (0.99, 5), (1.0, 6)], [(0.7, 7)]]
    for P in (2, 3, 6):
        el_dof, el_other, el_w = element_interpolation(m2e, len(m2e), P, 9)
        ref = {}
        for dof, mapping in enumerate(O.create_dof_to_element_map(m2e, P)):
            for element, other, w in mapping:
                ref[element] = (dof, other, w)
        for e in range(9):
            if e in ref:
                assert (el_dof[e], el_other[e], el_w[e]) == ref[e], (P, e)
            else:
                assert el_dof[e] == -1


@pytest.mark.gpu
@pytest.mark.parametrize('name', VTU_CASES)
def test_device_element_fields_match_reference(name):
    from openccm_b200.postprocessing.vtu_output import element_fields
    c, m2e, P, fields = _inputs(name)
    out = element_fields(c, m2e, P, fields.shape[2])
    np.testing.assert_array_equal(out, fields)


@pytest.mark.gpu
def test_device_element_fields_from_recorded_output(tmp_path):
    """No host round trip: the integrator's recorded output (device layout) goes straight into the element kernel."""
    import torch
    from _util import setup_case
    from oracle import openccm_oracle as O
    from openccm_b200.device import DeviceSystem
    from openccm_b200.postprocessing.vtu_output import element_fields
    from openccm_b200.system_solvers.export import export_system
    name = 'pfr_net8'
    cp, cmesh, net, mn = setup_case(name, 'B200-RK45', tmp_path)
    host = export_system('pfr', mn, cp, cmesh)
    dev = DeviceSystem(host)
    y0 = torch.as_tensor(host.to_device_layout(host.c0.reshape(-1)), device='cuda').reshape(1, -1)
    te = np.linspace(0, 10, 6)
    res = dev.rk45(y0, (0.0, 10.0), rtol=1e-8, atol=1e-10, first_step=1e-4, t_eval=te)
    n_el = len(cmesh.element_sizes)
    out = element_fields(res.y[0].contiguous(), mn[-1], host.P, n_el, device_layout=True)
    c = res.y[0].cpu().numpy().reshape(len(te), host.n_points, host.S).transpose(2, 1, 0)
    np.testing.assert_array_equal(out, O.element_fields(np.ascontiguousarray(c), mn[-1], host.P, n_el))


def test_export_writes_the_reference_folder_layout(tmp_path, monkeypatch):
    """Host logic of the export mirror (folders, symlinks, one buffer per time and species) with a stub kernel result."""
    from openccm_b200.postprocessing import vtu_output as V
    name = 'pfr_single'
    c, m2e, P, fields = _inputs(name)
    monkeypatch.setattr(V, 'element_fields', lambda *a, **k: fields)
    from _util import setup_case
    cp, cmesh, net, mn = setup_case(name, 'B200-RK45', tmp_path)
    cp['SETUP']['output_folder_path'] = str(tmp_path) + '/out/'
    cp['POST-PROCESSING']['vtu_dir'] = 'vtu/'
    ts = np.array([0.0, 1.0, 2.5, 20.0])
    os.makedirs(tmp_path / 'out' / 'vtu' / '0.0')
    (tmp_path / 'out' / 'vtu' / '0.0' / 'compartments').write_text('x')
    seen = []
    V.export_to_vtu_openfoam(c, ts, m2e, cp, cmesh, write_buffer=lambda buf, path, t, var, cm: seen.append((path, t, var, buf.copy())))
    assert len(seen) == len(ts) * c.shape[0]
    assert os.path.islink(tmp_path / 'out' / 'vtu' / '2.5' / 'compartments')
    assert seen[0][0].endswith('/vtu/0.0/c_a') and seen[-1][0].endswith('/vtu/20.0/c_b')
    np.testing.assert_array_equal(seen[3][3], fields[1, 1])
