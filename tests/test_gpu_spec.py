"""Mechanism-specialised kernel images (csrc/occm_spec.h) on the GPU: the plain RHS and the RK45 solution of a system with an
image loaded are those of the table-driven kernels — bit for bit for the RHS (same operations, same roundings), to round-off of
the error norm's summation order for complete solves — and the loader refuses images that do not belong to the system."""
import numpy as np
import pytest

import cases as C
from _util import setup_case

pytestmark = pytest.mark.gpu


def _case_host(name, tmp_path):
    from openccm_b200.system_solvers.export import export_system
    cp, cmesh, net, mn = setup_case(name, 'B200-RK45', tmp_path)
    return export_system(C.CASES[name]['model'], mn, cp, cmesh)


def _synthetic_host(n_cstr, n_species, n_reactions, seed):
    from openccm_b200.system_solvers.export import export_system
    from openccm_b200.workloads import cstr_ensemble_workload
    model, net, cp, cmesh = cstr_ensemble_workload(n_cstr=n_cstr, n_species=n_species, n_reactions=n_reactions, seed=seed)
    return export_system(model, net, cp, cmesh)


def _members(host, M, seed):
    import torch
    from openccm_b200.device import MemberParams
    rng = np.random.default_rng(seed)
    ks = 10.0 ** rng.uniform(-0.5, 0.5, (M, max(host.tables.n_rxn, 1)))
    qs = rng.uniform(0.5, 2.0, M)
    return MemberParams(M, torch.as_tensor(ks, device='cuda'), torch.as_tensor(qs, device='cuda'), None)


def _compare(host, M, t_end, seed=3):
    import torch
    from openccm_b200.device import DeviceSystem
    rng = np.random.default_rng(seed)
    y = torch.as_tensor(np.abs(rng.normal(0.5, 0.5, (M, host.ndof))), device='cuda')
    t = torch.as_tensor(rng.uniform(0, 2, M), device='cuda')
    mem = _members(host, M, seed + 1)
    plain = DeviceSystem(host, specialize='off')
    spec = DeviceSystem(host, specialize='on')
    assert not plain.specialized and spec.specialized and spec.spec_image
    assert spec.lib.occm_system_is_specialized(spec.handle) == 1 and plain.lib.occm_system_is_specialized(plain.handle) == 0
    for members in (None, mem):
        a = plain.rhs(y, t, members)
        b = spec.rhs(y, t, members)
        assert torch.equal(a, b), float((a - b).abs().max())
    kw = dict(rtol=1e-7, atol=1e-10, first_step=1e-4, t_eval=np.linspace(0, t_end, 9), members=mem)
    ra = plain.rk45(y, (0.0, t_end), **kw)
    rb = spec.rk45(y, (0.0, t_end), **kw)
    assert np.all(ra.status == 1) and np.all(rb.status == 1)
    # the error norm is summed in another order (per-warp partials of other chunk sizes): step sizes agree to round-off
    np.testing.assert_allclose(rb.y.cpu().numpy(), ra.y.cpu().numpy(), rtol=1e-10, atol=1e-13)
    assert np.all(np.abs(ra.nfev - rb.nfev) <= 6)
    return plain, spec


@pytest.mark.parametrize('name', ['cstr_net12_nacl', 'openfoam_2d_cstr_nacl', 'cstr_reversible', 'cstr_zeroth', 'cstr_quirk_sum'])
def test_reference_cases(name, tmp_path):
    """Small networks of the reference's own cases (boundary rows, pulses and wide rows go through occm_fixup)."""
    _compare(_case_host(name, tmp_path), M=5, t_end=3.0)


@pytest.mark.parametrize('n_cstr,S,R,M', [(3000, 10, 5, 6), (2500, 12, 8, 3), (5000, 4, 3, 4), (4100, 16, 12, 2), (333, 4, 2, 37)])
def test_synthetic_networks(n_cstr, S, R, M):
    """Ring + permutation networks (ELL width 3): chunk tails, members that wrap around the persistent grid, 4..16 species."""
    _compare(_synthetic_host(n_cstr, S, R, seed=S), M=M, t_end=1.0)


@pytest.mark.parametrize('name', ['pfr_net8', 'pfr_net8_nacl', 'pfr_single'])
def test_pfr_reference_cases(name, tmp_path):
    """PFR networks: one upwind entry per point, inlet nodes through occm_fixup."""
    _compare(_case_host(name, tmp_path), M=3, t_end=2.0)


def test_pfr_bdf_newton_krylov_is_unchanged():
    """Stiff PFR chains through BDF + GMRES: the residual (plain RHS) and the matrix-free Jacobian action (RHS of y + eps z, the
    stage-1 kernel) of the image are bit-identical to the table-driven ones, so the whole solve is — same steps, same orders,
    same iteration counts, same numbers."""
    import torch
    from openccm_b200.bdf import solve_bdf
    from openccm_b200.device import DeviceSystem
    from openccm_b200.system_solvers.export import export_system
    from openccm_b200.workloads import pfr_stiff_workload
    model, net, cp, cmesh = pfr_stiff_workload(n_chains=6, chain_len=5, P=40, t_end=2.0)
    host = export_system(model, net, cp, cmesh)
    y0 = torch.as_tensor(host.to_device_layout(host.c0.reshape(-1)), device='cuda')[None].contiguous()
    out = []
    for how in ('off', 'on'):
        dev = DeviceSystem(host, specialize=how)
        assert dev.specialized == (how == 'on')
        out.append(solve_bdf(dev, y0, (0.0, 2.0), rtol=1e-8, atol=1e-12, first_step=1e-5, t_eval=np.linspace(0, 2, 9)))
    a, b = out
    assert int(a.status[0]) == 1 and int(b.status[0]) == 1
    assert int(a.nfev[0]) == int(b.nfev[0]) and int(a.n_accepted[0]) == int(b.n_accepted[0])
    assert torch.equal(a.y, b.y)


@pytest.mark.parametrize('kind', ['pfr', 'cstr'])
def test_bdf_ensembles_with_members_finishing_at_different_attempts(kind):
    """Batched BDF: members with other flow / rate scales take other step sequences and finish at different attempts, so the RHS
    and J*v launches run under the active-member mask; image and table-driven kernels give identical solves."""
    import torch
    from openccm_b200.bdf import solve_bdf
    from openccm_b200.device import DeviceSystem, MemberParams
    from openccm_b200.system_solvers.export import export_system
    from openccm_b200.workloads import cstr_ensemble_workload, pfr_stiff_workload
    if kind == 'pfr':
        model, net, cp, cmesh = pfr_stiff_workload(n_chains=4, chain_len=4, P=30, t_end=1.5)
    else:
        model, net, cp, cmesh = cstr_ensemble_workload(n_cstr=1500, n_species=6, n_reactions=4, seed=7)
    host = export_system(model, net, cp, cmesh)
    M = 5
    rng = np.random.default_rng(2)
    ks = 10.0 ** rng.uniform(-0.7, 0.7, (M, host.tables.n_rxn))
    qs = rng.uniform(0.3, 3.0, M)
    mem = MemberParams(M, torch.as_tensor(ks, device='cuda'), torch.as_tensor(qs, device='cuda'), None)
    y0 = torch.as_tensor(host.to_device_layout(host.c0.reshape(-1)), device='cuda')[None].repeat(M, 1).contiguous()
    if kind == 'cstr':
        y0 = y0 * torch.as_tensor(rng.uniform(0.2, 1.0, M), device='cuda')[:, None]
    out = []
    for how in ('off', 'on'):
        dev = DeviceSystem(host, specialize=how)
        assert dev.specialized == (how == 'on')
        out.append(solve_bdf(dev, y0, (0.0, 1.5), rtol=1e-7, atol=1e-11, first_step=1e-5, t_eval=np.linspace(0, 1.5, 7), members=mem))
    a, b = out
    assert np.all(a.status == 1) and np.all(b.status == 1)
    assert len(set((a.n_accepted + a.n_rejected).tolist())) > 1          # the members really take different numbers of attempts
    assert np.array_equal(a.nfev, b.nfev) and np.array_equal(a.n_accepted, b.n_accepted)
    assert torch.equal(a.y, b.y)


def test_ring_network_ell_width_two():
    """A pure ring (diagonal + predecessor: ELL width 2, every source inside or next to the staged chunk)."""
    from openccm_b200 import synthetic as syn
    from openccm_b200.device import DeviceSystem
    from openccm_b200.system_solvers.export import export_system
    from openccm_b200.workloads import cstr_ensemble_workload
    model, _, cp, cmesh = cstr_ensemble_workload(n_cstr=2600, n_species=8, n_reactions=5, seed=3)
    net = syn.random_network(2600, n_io=16, seed=3, permutation_edges=False)
    host = export_system(model, net, cp, cmesh)
    plain, spec = _compare(host, M=4, t_end=1.0)
    assert spec.ell_k == 2


def test_auto_mode_specialises_large_networks_only(monkeypatch):
    from openccm_b200.device import DeviceSystem
    monkeypatch.delenv('OCCM_B200_SPECIALIZE', raising=False)
    assert not DeviceSystem(_synthetic_host(1000, 10, 5, seed=0)).specialized
    assert DeviceSystem(_synthetic_host(5000, 10, 5, seed=0)).specialized
    monkeypatch.setenv('OCCM_B200_SPECIALIZE', '0')
    assert not DeviceSystem(_synthetic_host(5000, 10, 5, seed=0)).specialized


def test_ineligible_systems_keep_the_table_driven_kernels(tmp_path):
    import torch
    from openccm_b200.device import DeviceSystem
    for name in ('pfr_net6_pointmass', 'cstr_net20_timebc', 'cstr_monod'):
        host = _case_host(name, tmp_path)
        dev = DeviceSystem(host, specialize='on')
        assert not dev.specialized
        y = torch.as_tensor(np.full((1, host.ndof), 0.3), device='cuda')
        assert torch.isfinite(dev.rhs(y, 0.0)).all()


def test_loader_rejects_foreign_images(tmp_path):
    """An image generated from another mechanism, for another ELL width, or that is not a cubin at all."""
    from openccm_b200 import _lib, specialize as sp
    from openccm_b200.device import DeviceSystem
    host = _synthetic_host(3000, 10, 5, seed=1)
    other = _synthetic_host(3000, 10, 5, seed=2)                       # same shape, other rate constants
    dev = DeviceSystem(host, specialize='off')
    load = lambda img: dev.lib.occm_system_load_specialized(dev.handle, img, len(img))
    with open(sp.build_image(other.tables, other.tables.pack(), other.S, dev.ell_k), 'rb') as fh:
        assert load(fh.read()) == -1 and 'table hash' in dev.lib.occm_last_error().decode()
    with open(sp.build_image(host.tables, dev.tables_blob, host.S, dev.ell_k + 1), 'rb') as fh:
        assert load(fh.read()) == -1 and 'does not match' in dev.lib.occm_last_error().decode()
    assert load(b'\x7fELF' + b'\0' * 200) == -2
    assert dev.lib.occm_system_is_specialized(dev.handle) == 0
    with open(sp.build_image(host.tables, dev.tables_blob, host.S, dev.ell_k), 'rb') as fh:
        assert load(fh.read()) == 0
    assert dev.lib.occm_system_is_specialized(dev.handle) == 1
    pfr = DeviceSystem(_case_host('pfr_net8', tmp_path), specialize='off')
    with open(sp.build_image(host.tables, dev.tables_blob, host.S, dev.ell_k), 'rb') as fh:
        img = fh.read()
    assert pfr.lib.occm_system_load_specialized(pfr.handle, img, len(img)) == -1 and 'model' in pfr.lib.occm_last_error().decode()


def test_finished_members_are_skipped():
    """Members that reach t_bound early stop being integrated (status != running): the specialised stage kernels skip them
    exactly like the table-driven ones (same outputs, same evaluation counts per member)."""
    import torch
    from openccm_b200.device import DeviceSystem, MemberParams
    host = _synthetic_host(4200, 10, 5, seed=4)
    M = 12
    rng = np.random.default_rng(0)
    y = torch.as_tensor(np.abs(rng.normal(0.5, 0.3, (M, host.ndof))), device='cuda')
    ks = np.ones((M, host.tables.n_rxn)); ks[::3] *= 30.0              # every third member is much stiffer: many more steps
    mem = MemberParams(M, torch.as_tensor(ks, device='cuda'), torch.as_tensor(np.ones(M), device='cuda'), None)
    out = []
    for how in ('off', 'on'):
        dev = DeviceSystem(host, specialize=how)
        r = dev.rk45(y, (0.0, 2.0), rtol=1e-6, atol=1e-10, first_step=1e-4, t_eval=np.linspace(0, 2, 5), members=mem)
        assert np.all(r.status == 1)
        out.append(r)
    assert out[0].nfev.max() > 1.3 * out[0].nfev.min()                 # the members really finish at different launches
    np.testing.assert_allclose(out[1].y.cpu().numpy(), out[0].y.cpu().numpy(), rtol=1e-10, atol=1e-13)
    assert np.all(np.abs(out[0].nfev - out[1].nfev) <= 6)
