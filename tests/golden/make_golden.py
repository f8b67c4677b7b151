#!/usr/bin/env python
"""
Generate the golden vectors by running the UNMODIFIED reference (``/root/reference``) on every case of
``cases.py``.  Runs only in the build container (the GPU box has no /root/reference); its outputs
(``<case>.npz``, ``parser_kats.json``) are committed so that the tests never need the reference.

    python tests/golden/make_golden.py [case ...]          # all cases when none is named
    python tests/golden/make_golden.py --openfoam-fixture  # (re)build the 444-PFR network fixture first

Harness-side work-arounds for reference defects; the reference files are untouched:
 1. ``pyparsing`` is not installed -> alias pip's vendored copy before importing ``reactions.py``.
 2. ``reactions.py:214`` sets ``np.seterr(all='raise')`` globally -> reset inside the ``solve_ivp`` wrapper.
 3. scipy BDF/Radau pass non-contiguous ``y`` -> ``np.ascontiguousarray`` in the wrapper.
"""
from __future__ import annotations

import json
import os
import pickle
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

try:
    import pyparsing  # noqa: F401
except ModuleNotFoundError:
    import pip._vendor.pyparsing as _pp
    sys.modules['pyparsing'] = _pp
sys.path.insert(0, '/root/reference')

import numpy as np
import scipy.integrate

import cases as C  # noqa: E402

from openccm import ConfigParser  # noqa: E402  (the reference)
from openccm.mesh import GroupedBCs  # noqa: E402
import openccm.system_solvers.cstr_system as ref_cstr  # noqa: E402
import openccm.system_solvers.pfr_system as ref_pfr  # noqa: E402
from openccm.system_solvers import solve_system as ref_solve_system  # noqa: E402
from openccm.system_solvers.reactions import parse_reactions  # noqa: E402
from openccm.postprocessing.analysis import network_to_rtd  # noqa: E402

CAPTURE = {}


def _patched_solve_ivp(fun, t_span, y0, **kw):
    np.seterr(all='warn', under='ignore')
    args = kw.pop('args', ())
    f = lambda t, y: fun(t, np.ascontiguousarray(y), *args)
    res = scipy.integrate.solve_ivp(f, t_span, y0, **kw)
    CAPTURE.update(fun=fun, args=args, y0=np.array(y0, copy=True), res=res)
    return res


ref_cstr.solve_ivp = _patched_solve_ivp
ref_pfr.solve_ivp = _patched_solve_ivp


def run_reference(name: str, solver: str, t_eval_override=None, workdir=None, **sim_overrides):
    case = C.CASES[name]
    net = case['net']()
    os.makedirs('cache', exist_ok=True)
    if case['rxn']:
        with open('reactions', 'w') as fh:
            fh.write(C.RXN[case['rxn']])
    with open('CONFIG', 'w') as fh:
        fh.write(C.config_text(name, solver, 'reactions'))
    cp = ConfigParser('CONFIG')
    cp['SETUP']['tmp_folder_path'] = 'cache/'
    cp.need_to_update_paths = False
    if t_eval_override is not None:
        cp['SIMULATION']['t_eval'] = t_eval_override
    for k, v in sim_overrides.items():
        cp['SIMULATION'][k] = str(v)
    gb = GroupedBCs(cp)
    cmesh = C.make_cmesh(name, net, gb)
    model_network = net.to_model_network()
    c, t, inlet_map, outlet_map = ref_solve_system(case['model'], model_network, cp, cmesh)
    return dict(c=c, t=t, inlet_map=inlet_map, outlet_map=outlet_map, cp=cp, cmesh=cmesh, net=model_network,
                capture=dict(CAPTURE))


def golden_for(name: str) -> None:
    case = C.CASES[name]
    solvers = case.get('solvers', ('RK45', 'LSODA'))
    out = {}
    old = os.getcwd()
    tmp = tempfile.mkdtemp(prefix=f'golden_{name}_')
    os.chdir(tmp)
    try:
        for solver in solvers:
            r = run_reference(name, solver)
            res = r['capture']['res']
            out[f'c_{solver}'] = r['c']
            out[f't_{solver}'] = r['t']
            out[f'nfev_{solver}'] = np.int64(res.nfev)
            if case['rtd']:
                out[f'rtd_{solver}'] = network_to_rtd((r['c'], r['t'], r['inlet_map'], r['outlet_map']),
                                                      r['cmesh'], r['cp'], r['net'])
        # raw RHS evaluations at random states (inner boundary)
        cap = r['capture']
        rng = np.random.default_rng(C.hash_name(name) + 1)
        y0 = cap['y0']
        t0, tf = [float(v) for v in case['t_span'].split(', ')]
        n_states = 6
        ys = np.empty((n_states, y0.size)); ts = np.empty(n_states); dys = np.empty((n_states, y0.size))
        for k in range(n_states):
            ys[k] = y0 if k == 0 else np.abs(y0 + rng.normal(0.0, 0.5, y0.size))
            ts[k] = t0 if k == 0 else rng.uniform(t0, min(tf, t0 + 2.0))
            if k == 1:
                ts[k] = t0 + 0.0005   # inside the smoothed-Heaviside ramp of H(10*t)
            np.seterr(all='warn', under='ignore')
            dys[k] = cap['fun'](ts[k], ys[k].copy(), *cap['args'])
        out.update(c0=y0, ddt_y=ys, ddt_t=ts, ddt_out=dys)
        # step-exactness data: RK45 with t_eval = all (the solver's own steps)
        if 'RK45' in solvers and name in ('cstr_irreversible', 'cstr_net12_nacl', 'pfr_net8', 'cstr_net20_timebc'):
            r = run_reference(name, 'RK45', t_eval_override='all')
            out['c_RK45_all'] = r['c']; out['t_RK45_all'] = r['t']
            out['nfev_RK45_all'] = np.int64(r['capture']['res'].nfev)
    finally:
        os.chdir(old)
        shutil.rmtree(tmp, ignore_errors=True)
    np.savez_compressed(os.path.join(HERE, f'{name}.npz'), **out)
    print(f'[golden] {name}: ' + ', '.join(f'{k}{tuple(np.shape(v))}' for k, v in out.items()))


def parser_kats() -> None:
    """Reaction-parser known answers: the per-species rate strings of reactions.py:270-392 and their
    sympy-simplified form (reactions.py:421,451)."""
    import configparser
    import sympy as sp
    kats = {}
    for key, text in C.RXN.items():
        p = configparser.ConfigParser(); p.optionxform = str; p.read_string(text)
        raw = parse_reactions(dict(p['REACTIONS']), dict(p['RATE CONSTANTS']))
        kats[key] = {'raw': raw, 'sympy': {k: str(sp.parse_expr(v)) for k, v in raw.items()}}
    with open(os.path.join(HERE, 'parser_kats.json'), 'w') as fh:
        json.dump(kats, fh, indent=1, sort_keys=True)
    print('[golden] parser_kats.json')


VTU_CASES = ('pfr_net8', 'cstr_net20_timebc', 'pfr_single', 'openfoam_2d_pfr')


def vtu_golden(name: str) -> None:
    """f-3: the per-(time, species) element buffers the UNMODIFIED `export_to_vtu_openfoam` hands to its file writer, for
    the golden solution of `name` (the writer is replaced by a capture, the folder layout is prepared in a scratch dir)."""
    import openccm.postprocessing.vtu_output as ref_vtu
    case = C.CASES[name]
    g = np.load(os.path.join(HERE, f'{name}.npz'))
    solver = 'RK45' if 'c_RK45' in g.files else 'LSODA'
    c, ts = g[f'c_{solver}'], g[f't_{solver}']
    keep = np.unique(np.linspace(0, len(ts) - 1, 4).astype(int))           # a few output times are enough
    c, ts = np.ascontiguousarray(c[:, :, keep]), ts[keep]
    net = case['net']()
    tmp = tempfile.mkdtemp(prefix=f'golden_vtu_{name}_')
    old = os.getcwd(); os.chdir(tmp)
    captured = []
    orig = ref_vtu.write_buffer_to_file_openfoam
    ref_vtu.write_buffer_to_file_openfoam = lambda buffer, path, t, var, cmesh: captured.append(np.array(buffer, copy=True))
    try:
        os.makedirs('cache', exist_ok=True)
        with open('CONFIG', 'w') as fh:
            fh.write(C.config_text(name, solver, 'reactions'))
        cp = ConfigParser('CONFIG')
        cp['SETUP']['output_folder_path'] = tmp + '/out/'
        cp['POST-PROCESSING']['vtu_dir'] = 'vtu/'
        cp.need_to_update_paths = False
        t0 = cp.get_list(['SIMULATION', 't_span'], float)[0]
        os.makedirs(os.path.join(tmp, 'out', 'vtu', str(t0)), exist_ok=True)
        if ts[0] != t0:                      # the reference creates the first folder with exist_ok only for i_t == 0
            os.makedirs(os.path.join(tmp, 'out', 'vtu', str(ts[0])), exist_ok=True)
        gb = GroupedBCs(cp)
        cmesh = C.make_cmesh(name, net, gb)
        m2e = net.to_model_network()[-1]
        ref_vtu.export_to_vtu_openfoam(c, ts, m2e, cp, cmesh)
    finally:
        ref_vtu.write_buffer_to_file_openfoam = orig
        os.chdir(old)
        shutil.rmtree(tmp, ignore_errors=True)
    S = c.shape[0]
    fields = np.array(captured).reshape(len(ts), S, -1)
    np.savez_compressed(os.path.join(HERE, f'vtu_{name}.npz'), fields=fields, time_index=keep)
    print(f'[golden] vtu_{name}: fields{fields.shape} unmapped {(fields[0, 0] == -1).sum()}')


FLOW_CASES = ((8, 1), (40, 2), (300, 3), (1500, 4))


def flow_case(n: int, seed: int):
    """Inputs of a flow-balancing case (shared with tests/test_flow_balance.py): a random network whose flows are perturbed
    by 0.1 % so that no model conserves mass."""
    from types import SimpleNamespace
    from openccm_b200 import synthetic as syn
    net = syn.random_network(n, n_io=2, seed=seed)
    conn = net.to_model_network()[0]
    rng = np.random.default_rng(seed)
    flows = net.flows * (1 + 1e-3 * rng.standard_normal(net.flows.shape))
    ids_in = sorted({o for (i, _) in conn.values() for o in i.values() if o < 0})
    ids_out = sorted({o for (_, oo) in conn.values() for o in oo.values() if o < 0})
    return conn, flows, SimpleNamespace(domain_inlets=tuple(ids_in), domain_outlets=tuple(ids_out))


def flows_golden() -> None:
    """f-2: volumetric flows after the UNMODIFIED `tweak_final_flows` (helpers.py:268-438)."""
    from openccm.compartment_models.helpers import tweak_final_flows
    out = {}
    for n, seed in FLOW_CASES:
        conn, flows, gb = flow_case(n, seed)
        tweak_final_flows(conn, flows, gb, 1e-9)
        out[f'flows_{n}_{seed}'] = flows
    np.savez_compressed(os.path.join(HERE, 'flow_balance.npz'), **out)
    print('[golden] flow_balance.npz: ' + ', '.join(f'{k}{v.shape}' for k, v in out.items()))


def openfoam_cstr_fixture() -> None:
    """C3: the same OpenFOAM 2-D example compartmentalised with model = CSTR (76 CSTRs); network + geometry fixture."""
    import openccm
    tmp = tempfile.mkdtemp(prefix='golden_openfoam_cstr_')
    dst = os.path.join(tmp, 'case')
    shutil.copytree('/root/reference/examples/OpenFOAM/pipe_with_recirc', dst)
    old = os.getcwd(); os.chdir(dst)
    try:
        cp = ConfigParser('CONFIG')
        cp['COMPARTMENT MODELLING']['model'] = 'cstr'
        cp['SIMULATION']['points_per_pfr'] = '1'
        cp['SIMULATION']['run'] = 'False'
        cp['POST-PROCESSING']['calculate_rtd'] = 'False'
        openccm.run(cp)
        with open('cache/cstr_network.pickle', 'rb') as fh:
            network = pickle.load(fh)
        with open('cache/cmesh.pickle', 'rb') as fh:
            cmesh = pickle.load(fh)
    finally:
        os.chdir(old)
    from openccm_b200.network import NetworkArrays
    na = NetworkArrays.from_model_network(network)
    d = {k: getattr(na, k) for k in ('n_models', 'volumes', 'flows', 'in_model', 'in_conn', 'in_other',
                                     'out_model', 'out_conn', 'out_other')}
    d['model_to_element_map'] = [[(float(a), int(b)) for a, b in m] for m in na.model_to_element_map]
    with open(os.path.join(HERE, 'openfoam_2d_cstr_network.pickle'), 'wb') as fh:
        pickle.dump(d, fh, protocol=4)
    print(f'[golden] openfoam cstr fixture: {na.n_models} CSTRs, {na.flows.size} connections')
    shutil.rmtree(tmp, ignore_errors=True)


def cstr_builder_golden() -> None:
    """f-2: inputs and outputs of the UNMODIFIED `create_cstr_network` (compartment_models/cstr.py:138-251) on the OpenFOAM 2-D
    example compartmentalised with model = CSTR.  Inputs are stored flattened (no reference classes needed to load them)."""
    import openccm
    from openccm.compartment_models.cstr import create_cstr_network
    tmp = tempfile.mkdtemp(prefix='golden_cstr_builder_')
    dst = os.path.join(tmp, 'case')
    shutil.copytree('/root/reference/examples/OpenFOAM/pipe_with_recirc', dst)
    old = os.getcwd(); os.chdir(dst)
    try:
        cp = ConfigParser('CONFIG')
        cp['COMPARTMENT MODELLING']['model'] = 'cstr'
        cp['SIMULATION']['points_per_pfr'] = '1'
        cp['SIMULATION']['run'] = 'False'
        cp['POST-PROCESSING']['calculate_rtd'] = 'False'
        openccm.run(cp)
        load = lambda n: pickle.load(open(f'cache/{n}.pickle', 'rb'))
        cn, comp, cmesh = load('cstr_compartment_network'), load('cstr_compartments_post'), load('cmesh')
        fu = np.load('cache/flows_and_upwind.npy', allow_pickle=True)
        connections, volumes, flows, c2m, m2e = create_cstr_network(comp, cn, cmesh, fu, None, cp)
    finally:
        os.chdir(old)
    g_comp, g_nb, g_len, facets, elems = [], [], [], [], []
    for c, nbs in cn.items():
        for nb, fd in nbs.items():
            g_comp.append(c); g_nb.append(nb); g_len.append(len(fd)); facets.extend(fd.keys()); elems.extend(fd.values())
    facets = np.asarray(facets, dtype=np.int64)
    comp_ptr = np.concatenate([[0], np.cumsum([len(comp[c]) for c in sorted(comp)])])
    comp_el = np.concatenate([np.fromiter(comp[c], dtype=np.int64, count=len(comp[c])) for c in sorted(comp)])
    from openccm_b200.network import NetworkArrays
    na = NetworkArrays.from_model_network((connections, volumes, flows, c2m, m2e))
    np.savez_compressed(os.path.join(HERE, 'openfoam_2d_cstr_builder.npz'),
                        g_comp=np.asarray(g_comp), g_nb=np.asarray(g_nb), g_len=np.asarray(g_len), facets=facets, elems=np.asarray(elems, dtype=np.int64),
                        facet_el=np.asarray([(tuple(cmesh.facet_elements[f]) + (-1, -1))[:2] for f in facets.tolist()], dtype=np.int64),   # boundary facets: one element
                        facet_flow=fu[facets, 0].astype(np.float64), facet_flag=fu[facets, 1].astype(np.int64),
                        comp_ptr=comp_ptr, comp_el=comp_el, element_sizes=np.asarray(cmesh.element_sizes, dtype=np.float64),
                        domain_inlets=np.asarray(cmesh.grouped_bcs.domain_inlets), domain_outlets=np.asarray(cmesh.grouped_bcs.domain_outlets),
                        flow_threshold=cp.get_item(['COMPARTMENT MODELLING', 'flow_threshold'], float),
                        atol_opt=cp.get_item(['COMPARTMENT MODELLING', 'atol_opt'], float),
                        volumes=volumes, flows=flows, in_model=na.in_model, in_conn=na.in_conn, in_other=na.in_other,
                        out_model=na.out_model, out_conn=na.out_conn, out_other=na.out_other)
    print(f'[golden] openfoam_2d_cstr_builder: {len(g_comp)} compartment pairs, {len(facets)} facets -> {len(volumes)} CSTRs, {len(flows)} connections')
    shutil.rmtree(tmp, ignore_errors=True)


def openfoam_fixture() -> None:
    """Run the reference's own pipeline on examples/OpenFOAM/pipe_with_recirc (a writable copy) and freeze the
    444-PFR model network + element geometry as a fixture (inputs only)."""
    import openccm
    tmp = tempfile.mkdtemp(prefix='golden_openfoam_')
    dst = os.path.join(tmp, 'case')
    shutil.copytree('/root/reference/examples/OpenFOAM/pipe_with_recirc', dst)
    old = os.getcwd(); os.chdir(dst)
    try:
        cp = ConfigParser('CONFIG')
        cp['SIMULATION']['run'] = 'False'
        cp['POST-PROCESSING']['calculate_rtd'] = 'False'
        openccm.run(cp)
        with open('cache/pfr_network.pickle', 'rb') as fh:
            network = pickle.load(fh)
        with open('cache/cmesh.pickle', 'rb') as fh:
            cmesh = pickle.load(fh)
    finally:
        os.chdir(old)
    from openccm_b200.network import NetworkArrays
    na = NetworkArrays.from_model_network(network)
    d = {k: getattr(na, k) for k in ('n_models', 'volumes', 'flows', 'in_model', 'in_conn', 'in_other',
                                     'out_model', 'out_conn', 'out_other')}
    d['model_to_element_map'] = [[(float(a), int(b)) for a, b in m] for m in na.model_to_element_map]
    with open(os.path.join(HERE, 'openfoam_2d_pfr_network.pickle'), 'wb') as fh:
        pickle.dump(d, fh, protocol=4)
    np.savez_compressed(os.path.join(HERE, 'openfoam_2d_pfr_geometry.npz'),
                        centroids=np.asarray(cmesh.element_centroids), sizes=np.asarray(cmesh.element_sizes))
    print(f'[golden] openfoam fixture: {na.n_models} PFRs, {na.flows.size} connections, '
          f'{len(cmesh.element_sizes)} elements; bc ids inlet={cmesh.grouped_bcs.id("inlet")} '
          f'outlet={cmesh.grouped_bcs.id("outlet")}')
    shutil.rmtree(tmp, ignore_errors=True)


# Tight-tolerance solutions (the yard-stick of the north-star acceptance band 1e-10 + 1e-6 |ref|): two adaptive integrators
# only agree to about their tolerance, so "device vs scipy at the same tolerance" compares two noisy answers.  Each tight
# golden holds the UNMODIFIED reference run with LSODA at rtol 1e-13 / atol 1e-15, its agreement with a second tight solver
# in band units (evidence that it is converged), and the reference's own solves at the user tolerance (rtol 1e-6 / atol
# 1e-10, LSODA = the reference default, expanded_config_parser.py:55) so the tests can state where scipy itself is in band.
TIGHT_CASES = {
    'cstr_net9_tracer': 'DOP853', 'cstr_reversible': 'DOP853', 'pfr_chains_tracer': 'DOP853',
    'cstr_irreversible': 'DOP853', 'cstr_monod': 'DOP853', 'cstr_net12_nacl': 'DOP853', 'cstr_net20_timebc': 'DOP853',
    'pfr_single': 'DOP853', 'pfr_net8': 'DOP853', 'pfr_net8_nacl': 'DOP853', 'cstr_net10_pointmass': 'DOP853',
    'pfr_net6_pointmass': 'DOP853', 'openfoam_2d_pfr': 'Radau', 'openfoam_2d_cstr_nacl': 'Radau',
}
USER_RTOL, USER_ATOL = 1e-6, 1e-10


def tight_golden(name: str) -> None:
    case = C.CASES[name]
    out = {}
    old = os.getcwd()
    tmp = tempfile.mkdtemp(prefix=f'golden_tight_{name}_')
    os.chdir(tmp)
    try:
        r = run_reference(name, 'LSODA', rtol=1e-13, atol=1e-15)
        out['c_tight'] = r['c']; out['t'] = r['t']
        if case['rtd']:
            out['rtd_tight'] = network_to_rtd((r['c'], r['t'], r['inlet_map'], r['outlet_map']), r['cmesh'], r['cp'], r['net'])
        second = TIGHT_CASES[name]
        r2 = run_reference(name, second, rtol=1e-12, atol=1e-14)
        band = 1e-10 + 1e-6 * np.abs(out['c_tight'])
        out['tight_agreement_band_units'] = np.float64((np.abs(r2['c'] - out['c_tight']) / band).max())
        out['tight_second_solver'] = np.array(second)
        user = ['LSODA'] + (['RK45'] if 'RK45' in case.get('solvers', ('RK45', 'LSODA')) else [])
        for solver in user:
            ru = run_reference(name, solver, rtol=USER_RTOL, atol=USER_ATOL)
            out[f'c_user_{solver}'] = ru['c']
            out[f'nfev_user_{solver}'] = np.int64(ru['capture']['res'].nfev)
            if case['rtd']:
                out[f'rtd_user_{solver}'] = network_to_rtd((ru['c'], ru['t'], ru['inlet_map'], ru['outlet_map']), ru['cmesh'], ru['cp'], ru['net'])
    finally:
        os.chdir(old)
        shutil.rmtree(tmp, ignore_errors=True)
    np.savez_compressed(os.path.join(HERE, f'tight_{name}.npz'), **out)
    msg = ', '.join(f"{s}@user max {(np.abs(out[f'c_user_{s}'] - out['c_tight']) / band).max():.2f} band units" for s in user)
    print(f"[golden] tight_{name}: second solver {second} agrees to {float(out['tight_agreement_band_units']):.2e} band units; {msg}")


if __name__ == '__main__':
    argv = sys.argv[1:]
    if '--tight' in argv:
        argv.remove('--tight')
        for n in (argv or list(TIGHT_CASES)):
            tight_golden(n)
        sys.exit(0)
    if '--openfoam-fixture' in argv:
        argv.remove('--openfoam-fixture')
        openfoam_fixture()
    if '--openfoam-cstr-fixture' in argv:
        argv.remove('--openfoam-cstr-fixture')
        openfoam_cstr_fixture()
    if '--cstr-builder' in argv:
        argv.remove('--cstr-builder')
        cstr_builder_golden()
        if not argv:
            sys.exit(0)
    if '--flows' in argv:
        argv.remove('--flows')
        flows_golden()
        if not argv:
            sys.exit(0)
    if '--vtu' in argv:
        argv.remove('--vtu')
        for n in VTU_CASES:
            vtu_golden(n)
        if not argv:
            sys.exit(0)
    if '--kats' in argv:
        argv.remove('--kats')
        parser_kats()
        if not argv:
            sys.exit(0)
    names = argv or list(C.CASES)
    if not argv:
        parser_kats()
    for n in names:
        golden_for(n)
