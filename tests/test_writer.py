"""Asynchronous result writer (f-3): the files are the ones a synchronous np.save of the reference layout would have written."""
import os

import numpy as np
import pytest
import torch


def test_writer_host_arrays_and_layout(tmp_path):
    from openccm_b200.postprocessing import AsyncResultWriter, to_reference_layout
    T, NP, S = 7, 5, 3
    y = np.arange(T * NP * S, dtype=np.float64).reshape(T, NP * S)          # device layout: [T, point * S + species]
    seen = []
    with AsyncResultWriter(max_pending=1) as w:
        w.submit(str(tmp_path / 'c.npy'), y, to_reference_layout(NP, S))
        w.submit(str(tmp_path / 't.npy'), np.linspace(0, 1, T))
        w.submit(lambda a: seen.append(a.sum()), torch.ones(4, dtype=torch.float64))
    assert w.n_written == 3 and seen == [4.0]
    c = np.load(tmp_path / 'c.npy')
    assert c.shape == (S, NP, T)                                             # cstr_system.py:152-154
    for s in range(S):
        for p in range(NP):
            np.testing.assert_array_equal(c[s, p], y[:, p * S + s])
    np.testing.assert_array_equal(np.load(tmp_path / 't.npy'), np.linspace(0, 1, T))


def test_writer_surfaces_worker_errors(tmp_path):
    from openccm_b200.postprocessing import AsyncResultWriter
    w = AsyncResultWriter()
    w.submit(str(tmp_path / 'missing_dir' / 'x.npy'), np.zeros(3))
    with pytest.raises(OSError):
        w.close()


@pytest.mark.gpu
def test_writer_device_tensors_match_synchronous_save(tmp_path):
    """Recorded output of a device solve written while the next solve runs: identical to the synchronous reference-layout save."""
    import cases as C
    from _util import setup_case
    from openccm_b200.device import DeviceSystem
    from openccm_b200.postprocessing import AsyncResultWriter, to_reference_layout
    from openccm_b200.system_solvers.export import export_system
    cp, cmesh, net, mn = setup_case('cstr_net12_nacl', 'B200-RK45', tmp_path)
    host = export_system('cstr', mn, cp, cmesh)
    dev = DeviceSystem(host)
    y0 = torch.as_tensor(host.to_device_layout(host.c0.reshape(-1)), device='cuda').reshape(1, -1)
    te = np.linspace(0.0, 30.0, 61)
    with AsyncResultWriter() as w:
        outs = []
        for k in range(3):
            res = dev.rk45(y0 * (1.0 + 0.1 * k), (0.0, 30.0), rtol=1e-8, atol=1e-10, first_step=1e-4, t_eval=te)
            w.submit(str(tmp_path / f'c{k}.npy'), res.y[0], to_reference_layout(host.n_points, host.S))
            outs.append(res.y[0].clone())
    for k in range(3):
        ref = outs[k].cpu().numpy().reshape(len(te), host.n_points, host.S).transpose(2, 1, 0)
        np.testing.assert_array_equal(np.load(tmp_path / f'c{k}.npy'), ref)
