"""GPU parity of the fused right-hand side (occm_rhs through the C ABI) against the oracle and the reference's
golden raw-ddt values.  Bar: rel 1e-12 (fp64 round-off of differently ordered sums)."""
import numpy as np
import pytest

import cases as C
from _util import load_golden, setup_case

pytestmark = pytest.mark.gpu

ALL = sorted(C.CASES)


@pytest.mark.parametrize('name', ALL)
def test_rhs_matches_reference_golden(name, tmp_path):
    import torch
    from openccm_b200.device import DeviceSystem
    from openccm_b200.system_solvers.export import export_system
    g = load_golden(name)
    cp, cmesh, net, mn = setup_case(name, 'B200-RK45', tmp_path)
    host = export_system(C.CASES[name]['model'], mn, cp, cmesh)
    np.testing.assert_allclose(host.c0.reshape(-1), g['c0'], rtol=1e-13, atol=1e-15)
    dev = DeviceSystem(host)
    for y, t, ref in zip(g['ddt_y'], g['ddt_t'], g['ddt_out']):
        yd = torch.as_tensor(host.to_device_layout(y), device='cuda').reshape(1, -1)
        out = host.to_reference_layout(dev.rhs(yd, float(t)).cpu().numpy().reshape(-1))
        np.testing.assert_allclose(out, ref, rtol=1e-12, atol=1e-13 * max(1.0, np.abs(ref).max()))


@pytest.mark.parametrize('name', ['cstr_net12_nacl', 'pfr_net8', 'cstr_net20_timebc'])
def test_rhs_batched_members_vs_oracle(name, tmp_path):
    """Several members with per-member times in one launch == oracle called member by member."""
    import torch
    from oracle import openccm_oracle as O
    from openccm_b200.device import DeviceSystem
    from openccm_b200.system_solvers.export import export_system
    cp, cmesh, net, mn = setup_case(name, 'B200-RK45', tmp_path)
    host = export_system(C.CASES[name]['model'], mn, cp, cmesh)
    ora = O.build_system(C.CASES[name]['model'], mn, cp, cmesh)
    dev = DeviceSystem(host)
    rng = np.random.default_rng(5)
    M = 7
    ys = np.abs(rng.normal(0.5, 0.5, (M, host.ndof)))
    ts = rng.uniform(0.0, 2.0, M)
    yd = torch.as_tensor(host.to_device_layout(ys), device='cuda')
    out = dev.rhs(yd, torch.as_tensor(ts, device='cuda')).cpu().numpy()
    for m in range(M):
        ref = ora.fun(float(ts[m]), ys[m].copy())
        np.testing.assert_allclose(host.to_reference_layout(out[m]), ref, rtol=1e-12, atol=1e-13 * max(1.0, np.abs(ref).max()))


@pytest.mark.parametrize('name', ['cstr_net12_nacl', 'cstr_reversible', 'cstr_net20_timebc', 'pfr_single', 'pfr_net8', 'pfr_net8_nacl'])
def test_generic_and_fast_kernels_agree(name, tmp_path, monkeypatch):
    """The 128-bit fast kernels (even S; CSTR rows and PFR upwind rows; term lists in registers or, forced here, in
    shared memory) and the generic kernels give the same RHS and the same RK45 solution."""
    import torch
    from openccm_b200.device import DeviceSystem
    from openccm_b200.system_solvers.export import export_system
    cp, cmesh, net, mn = setup_case(name, 'B200-RK45', tmp_path)
    host = export_system(C.CASES[name]['model'], mn, cp, cmesh)
    rng = np.random.default_rng(11)
    M = 5
    y = torch.as_tensor(np.abs(rng.normal(0.5, 0.5, (M, host.ndof))), device='cuda')
    t = torch.as_tensor(rng.uniform(0, 2, M), device='cuda')
    outs, sols = [], []
    for mode in ('fast', 'generic', 'wide'):
        monkeypatch.delenv('OCCM_B200_NO_FAST', raising=False)
        monkeypatch.delenv('OCCM_B200_FORCE_WIDE', raising=False)
        if mode == 'generic':
            monkeypatch.setenv('OCCM_B200_NO_FAST', '1')
        if mode == 'wide':
            monkeypatch.setenv('OCCM_B200_FORCE_WIDE', '1')
        dev = DeviceSystem(host)
        outs.append(dev.rhs(y, t).cpu().numpy())
        res = dev.rk45(y, (0.0, 3.0), rtol=1e-8, atol=1e-10, first_step=1e-4, t_eval=np.linspace(0, 3, 7))
        assert np.all(res.status == 1)
        sols.append(res.y.cpu().numpy())
    for k in (1, 2):
        np.testing.assert_allclose(outs[0], outs[k], rtol=1e-13, atol=1e-14)
        np.testing.assert_allclose(sols[0], sols[k], rtol=1e-9, atol=1e-12)
    np.testing.assert_array_equal(outs[0], outs[2])          # registers vs shared memory: the same arithmetic, bit for bit


@pytest.mark.parametrize('model', ['cstr', 'pfr'])
def test_wide_mechanism_vs_oracle(model, tmp_path, monkeypatch):
    """A mechanism with more terms per species than the register form holds (stiff C5 mechanism: 6) runs on the
    shared-memory term-list kernels: RHS against the oracle, and against the generic table walk."""
    import torch
    from oracle import openccm_oracle as O
    from openccm_b200 import synthetic as syn
    from openccm_b200.device import DeviceSystem
    from openccm_b200.system_solvers.export import export_system
    from openccm_b200.workloads import _cmesh, _config, pfr_stiff_workload
    if model == 'pfr':
        _, net, cp, cmesh = pfr_stiff_workload(n_chains=3, chain_len=4, P=7, t_end=2.0, n_out=5, workdir=str(tmp_path))
    else:
        net = syn.random_network(40, n_io=3, seed=4)
        species, rxn_text = syn.stiff_mechanism(12, 20, seed=0)
        path = tmp_path / 'rxn_wide'
        path.write_text(rxn_text)
        cp = _config('cstr', species, str(path), t_span=(0, 2.0), t_eval='0.5, linear')
        cmesh = _cmesh(cp, 40, 9)
    mn = net.to_model_network()
    host = export_system(model, mn, cp, cmesh)
    n_terms = np.bincount([tm.specie for tm in host.tables.terms], minlength=host.S)
    assert n_terms.max() > 4
    ora = O.build_system(model, mn, cp, cmesh)
    rng = np.random.default_rng(2)
    M = 3
    ys = np.abs(rng.normal(0.5, 0.3, (M, host.ndof)))
    yd = torch.as_tensor(host.to_device_layout(ys), device='cuda')
    ts = torch.zeros(M, dtype=torch.float64, device='cuda')
    monkeypatch.delenv('OCCM_B200_NO_FAST', raising=False)
    fast = DeviceSystem(host).rhs(yd, ts).cpu().numpy()
    monkeypatch.setenv('OCCM_B200_NO_FAST', '1')
    generic = DeviceSystem(host).rhs(yd, ts).cpu().numpy()
    for m in range(M):
        ref = ora.fun(0.0, ys[m].copy())
        scale = max(1.0, np.abs(ref).max())
        np.testing.assert_allclose(host.to_reference_layout(fast[m]), ref, rtol=1e-11, atol=1e-12 * scale)
        np.testing.assert_allclose(host.to_reference_layout(generic[m]), ref, rtol=1e-11, atol=1e-12 * scale)


def _random_mechanism(n_species, n_reactions, max_order, seed):
    """Reactions file text: random reactants (total order <= max_order, stoichiometric coefficients 1..2), 1-2 products."""
    from openccm_b200 import synthetic as syn
    rng = np.random.default_rng(seed)
    names = syn.species_names(n_species)
    lines_r, lines_k, used = ['[REACTIONS]'], ['[RATE CONSTANTS]'], set()

    def side(idx, coeffs):
        return ' + '.join((f'{c}' if c > 1 else '') + names[i] for i, c in zip(idx, coeffs))
    r = 0
    for _ in range(n_reactions):
        n_react = int(rng.integers(1, 3))
        reactants = rng.choice(n_species, size=n_react, replace=False)
        coeffs = [1] * n_react
        while sum(coeffs) < max_order and rng.random() < 0.5:
            coeffs[int(rng.integers(0, n_react))] += 1
        rest = [i for i in range(n_species) if i not in reactants]
        products = rng.choice(rest, size=min(len(rest), int(rng.integers(1, 3))), replace=False)
        r += 1
        lines_r.append(f'R{r}: {side(reactants, coeffs)} -> {side(products, [int(rng.integers(1, 3)) for _ in products])}')
        lines_k.append(f'R{r}: {10 ** rng.uniform(-1.5, 0.5):.6e}')
        used.update(int(i) for i in reactants); used.update(int(i) for i in products)
    for s in range(n_species):
        if s not in used:
            r += 1
            lines_r.append(f'R{r}: {names[s]} -> {names[(s + 1) % n_species]}')
            lines_k.append(f'R{r}: 0.3')
    return names, '\n'.join(lines_r) + '\n\n' + '\n'.join(lines_k) + '\n'


@pytest.mark.parametrize('model,n_species,n_reactions,max_order,seed', [
    ('cstr', 2, 1, 1, 0), ('cstr', 2, 2, 2, 1), ('cstr', 4, 3, 3, 2), ('cstr', 6, 9, 3, 3), ('cstr', 16, 12, 2, 4),
    ('pfr', 2, 1, 2, 5), ('pfr', 4, 6, 3, 6), ('pfr', 8, 14, 2, 7), ('pfr', 6, 2, 1, 8), ('cstr', 3, 4, 3, 9),
    ('cstr', 34, 24, 2, 10), ('cstr', 64, 40, 2, 11), ('pfr', 22, 30, 3, 12)])
def test_random_mechanisms_all_kernel_forms(model, n_species, n_reactions, max_order, seed, tmp_path, monkeypatch):
    """Random mechanisms (orders 1..3, short and long term lists, even and odd species counts) on random networks: every
    kernel form the dispatcher can pick — register lists, shared-memory lists, generic table walk — against the oracle."""
    import torch
    from oracle import openccm_oracle as O
    from openccm_b200 import synthetic as syn
    from openccm_b200.device import DeviceSystem
    from openccm_b200.system_solvers.export import export_system
    from openccm_b200.workloads import _cmesh, _config
    species, text = _random_mechanism(n_species, n_reactions, max_order, seed)
    path = tmp_path / 'rxn'
    path.write_text(text)
    n_models = 30 if model == 'cstr' else 9
    net = syn.random_network(n_models, n_io=2, seed=seed + 20, elements_per_model=2 if model == 'pfr' else 1)
    cp = _config(model, species, str(path), t_span=(0, 1.0), t_eval='0.5, linear', P=7)
    cmesh = _cmesh(cp, n_models * (2 if model == 'pfr' else 1), seed)
    mn = net.to_model_network()
    host = export_system(model, mn, cp, cmesh)
    ora = O.build_system(model, mn, cp, cmesh)
    rng = np.random.default_rng(seed)
    M = 3
    ys = np.abs(rng.normal(0.6, 0.3, (M, host.ndof)))
    yd = torch.as_tensor(host.to_device_layout(ys), device='cuda')
    ts = torch.zeros(M, dtype=torch.float64, device='cuda')
    outs = {}
    for mode in ('fast', 'wide', 'generic'):
        monkeypatch.delenv('OCCM_B200_NO_FAST', raising=False)
        monkeypatch.delenv('OCCM_B200_FORCE_WIDE', raising=False)
        if mode == 'generic':
            monkeypatch.setenv('OCCM_B200_NO_FAST', '1')
        if mode == 'wide':
            monkeypatch.setenv('OCCM_B200_FORCE_WIDE', '1')
        outs[mode] = DeviceSystem(host).rhs(yd, ts).cpu().numpy()
    for m in range(M):
        ref = ora.fun(0.0, ys[m].copy())
        scale = max(1.0, np.abs(ref).max())
        for mode, out in outs.items():
            np.testing.assert_allclose(host.to_reference_layout(out[m]), ref, rtol=1e-11, atol=1e-12 * scale, err_msg=mode)
    np.testing.assert_array_equal(outs['fast'], outs['wide'])
