"""
Definitions of the golden cases: inputs only (network, CONFIG text, reactions text).

Shared by
* ``make_golden.py`` — runs the UNMODIFIED reference (/root/reference, in the build container) on each case and
  stores its outputs as ``<case>.npz``;
* ``tests/test_oracle_vs_golden.py`` — pins the CPU oracle against those outputs;
* ``tests/test_gpu_*.py`` — compares the CUDA path with the oracle / the golden outputs.

Nothing here imports the reference, the oracle or the device library.
"""
from __future__ import annotations

import os
import pickle
from types import SimpleNamespace
from typing import Dict

import numpy as np

from openccm_b200 import synthetic as syn
from openccm_b200.network import NetworkArrays

HERE = os.path.dirname(os.path.abspath(__file__))

RXN = {
    'irreversible': "[REACTIONS]\nR1: a -> b\nR2: b -> c\n\n[RATE CONSTANTS]\nR1: 0.5\nR2: 0.2\n",       # examples/simple_reactors/cstr/irreversible/reactions
    'reversible':   "[REACTIONS]\nR1: a -> b\nR2: b -> a\n\n[RATE CONSTANTS]\nR1: 1\nR2: 1\n",           # .../cstr/reversible/reactions
    'monod':        "[REACTIONS]\nR1: s + X -> 1.5X\n\n[RATE CONSTANTS]\nR1: 1/(0.25 + s)\n",            # .../cstr/monod/reactions
    'top_level':    "[REACTIONS]\nR1: a -> 2b\nR2: 3b -> c\n\n[RATE CONSTANTS]\nR1: 10\nR2:  5\n",       # /root/reference/CONFIG_REACTIONS
    'second_order': "[REACTIONS]\nR1: a + b -> c\n\n[RATE CONSTANTS]\nR1: 2.0\n",
    'zeroth':       "[REACTIONS]\nR1: a -> b\n\n[RATE CONSTANTS]\nR1: z=0.3\n",
    'quirk_sum':    "[REACTIONS]\nR1: a -> b\n\n[RATE CONSTANTS]\nR1: 1 + 2\n",                          # reactions.py:364 (no parentheses)
    'nacl':         "[REACTIONS]\nR1: 2NaCl   + CaCO3 ->  Na2CO3 + CaCl2\nR2:  Na2CO3 + CaCl2 -> 2NaCl   + CaCO3\n\n"
                    "[RATE CONSTANTS]\nR1: 5e-2\nR2: 2\n",                                               # examples/OpenCMP/pipe_with_recirc_2d/reactions
    'pfr_a2b':      "[REACTIONS]\nR1: a -> 2b\n\n[RATE CONSTANTS]\nR1: 0.1\n",                           # examples/simple_reactors/pfr/reactions
    'extra_terms':  "[REACTIONS]\nR1: a + b -> c\nR2: c -> a\n\n[RATE CONSTANTS]\nR1: kf*Ea/T\nR2: kr\n\n"
                    "[EXTRA TERMS]\nT: 300.0\nEa: 150.0\nkf: 2.5\nkr: kf/4\n",   # (functions such as exp() are a NameError in the reference's generated module)
}


def _cfg(model, species, ic, bc, t_span, t_eval, rxn, *, P=2, rtol=1e-8, atol=1e-10, first=1e-4, rtd=False):
    return dict(model=model, species=species, ic=ic, bc=bc, t_span=t_span, t_eval=t_eval, rxn=rxn, P=P,
                rtol=rtol, atol=atol, first=first, rtd=rtd)


def _openfoam_cstr_network() -> NetworkArrays:
    with open(os.path.join(HERE, 'openfoam_2d_cstr_network.pickle'), 'rb') as fh:
        d = pickle.load(fh)
    return NetworkArrays(**d)


def _openfoam_network() -> NetworkArrays:
    with open(os.path.join(HERE, 'openfoam_2d_pfr_network.pickle'), 'rb') as fh:
        d = pickle.load(fh)
    return NetworkArrays(**d)


CASES: Dict[str, dict] = {
    # ---- C1: examples/simple_reactors (one CSTR, V=1, Q=0.1)
    'cstr_irreversible': dict(net=lambda: syn.single_model(), **_cfg(
        'cstr', 'a, b, c', ['a -> 2', 'b -> 0', 'c -> 0'], ['a: inlet -> 2', 'b: inlet -> 0', 'c: inlet -> 0'],
        '0, 50', '1.0, linear', 'irreversible')),
    'cstr_reversible': dict(net=lambda: syn.single_model(), **_cfg(
        'cstr', 'a, b', ['a -> 0', 'b -> 0'], ['a: inlet -> 1', 'b: inlet -> 0'], '0, 70', '1.0, linear', 'reversible')),
    'cstr_monod': dict(net=lambda: syn.single_model(), **_cfg(
        'cstr', 'X, s', ['X -> 0.3', 's -> 0.01'], ['X: inlet -> 0', 's: inlet -> 1'], '0, 50', '1.0, linear', 'monod')),
    'cstr_top_level': dict(net=lambda: syn.single_model(), **_cfg(
        'cstr', 'a, b, c', ['a -> 1', 'b -> 0', 'c -> 0'], ['a: inlet -> 1', 'b: inlet -> 0', 'c: inlet -> 0'],
        '0, 5', '0.1, linear', 'top_level')),
    'cstr_second_order': dict(net=lambda: syn.single_model(), **_cfg(
        'cstr', 'a, b, c', ['a -> 1', 'b -> 0.5', 'c -> 0'], ['a: inlet -> 1', 'b: inlet -> 2', 'c: inlet -> 0'],
        '0, 50', '0.0, 0.3, 1.7, 5.0, 12.5, 50.0', 'second_order')),
    'cstr_zeroth': dict(net=lambda: syn.single_model(), **_cfg(
        'cstr', 'a, b', ['a -> 5', 'b -> 0'], ['a: inlet -> 5', 'b: inlet -> 0'], '0, 5', '0.5, linear', 'zeroth')),
    'cstr_quirk_sum': dict(net=lambda: syn.single_model(), **_cfg(
        'cstr', 'a, b', ['a -> 1', 'b -> 0'], ['a: inlet -> 1', 'b: inlet -> 0'], '0, 3', '0.25, linear', 'quirk_sum')),
    'cstr_extra_terms': dict(net=lambda: syn.single_model(), **_cfg(
        'cstr', 'a, b, c', ['a -> 1', 'b -> 1', 'c -> 0'], ['a: inlet -> 1', 'b: inlet -> 1', 'c: inlet -> 0'],
        '0, 20', '1.0, linear', 'extra_terms')),
    'cstr_log_teval': dict(net=lambda: syn.single_model(), **_cfg(
        'cstr', 'a, b, c', ['a -> 2', 'b -> 0', 'c -> 0'], ['a: inlet -> 2', 'b: inlet -> 0', 'c: inlet -> 0'],
        '0, 50', 'log, 12', 'irreversible')),
    # ---- C3-like: CSTR networks
    'cstr_net12_nacl': dict(net=lambda: syn.random_network(12, n_io=2, seed=3), **_cfg(
        'cstr', 'NaCl, CaCO3, Na2CO3, CaCl2', ['NaCl -> 0', 'CaCO3 -> 0', 'Na2CO3 -> 0', 'CaCl2 -> 0'],
        ['NaCl: inlet -> 1', 'CaCO3: inlet -> 1', 'Na2CO3: inlet -> 0', 'CaCl2: inlet -> 0'],
        '0, 30', '0.5, linear', 'nacl')),
    'cstr_net20_timebc': dict(net=lambda: syn.random_network(20, n_io=3, seed=7, elements_per_model=2), **_cfg(
        'cstr', 'a, b, c', ['a -> x + 2*y', 'b -> 0.1', 'b -> 1 @ x > 0.5', 'c -> 0'],
        ['a: inlet -> 1 + 0.5*sin(t)', 'b: inlet -> H(t - 1)', 'c: inlet -> 0'],
        '0, 10', '0.25, linear', 'second_order')),
    'cstr_net9_tracer': dict(net=lambda: syn.random_network(9, n_io=2, seed=11), **_cfg(
        'cstr', 'a', ['a -> 0'], ['a: inlet -> 1'], '0, 40', '0.5, linear', None, rtd=True)),
    # N = 2000: the largest size at which the reference's dense N x N operator is still cheap; pins the oracle's sparse
    # ("reference-sparse") assembly against the unmodified reference, odd species count
    'cstr_net2000': dict(net=lambda: syn.random_network(2000, n_io=4, seed=23), **_cfg(
        'cstr', 'a, b, c', ['a -> 1', 'b -> 0.5', 'c -> 0'], ['a: inlet -> 1', 'b: inlet -> 2', 'c: inlet -> 0'],
        '0, 2', '0.5, linear', 'second_order', rtol=1e-6), solvers=('RK45',)),
    'cstr_net10_pointmass': dict(net=lambda: syn.random_network(10, n_io=1, seed=13), **_cfg(
        'cstr', 'a', ['a -> 2.0 @ point(0.3, 0.4)'], [], '0, 5', '0.1, linear', None, first=1e-6)),
    # ---- C1 / C2-like: PFRs
    'pfr_single': dict(net=lambda: syn.single_model(), **_cfg(
        'pfr', 'a, b', ['a -> 0', 'b -> 0'], ['a: inlet -> H(10*t)', 'b: inlet -> 0'], '0, 20', '0.5, linear',
        'pfr_a2b', P=21)),
    'pfr_net8': dict(net=lambda: syn.random_network(8, n_io=2, seed=5, elements_per_model=3), **_cfg(
        'pfr', 'a, b', ['a -> 0.2*x', 'b -> 0'], ['a: inlet -> H(10*t)', 'b: inlet -> 0.3'], '0, 10', '0.25, linear',
        'pfr_a2b', P=5)),
    'pfr_chains_tracer': dict(net=lambda: syn.pfr_chains(3, 4, seed=2), **_cfg(
        'pfr', 'a', ['a -> 0'], ['a: inlet -> 1'], '0, 30', '0.25, linear', None, P=4, rtd=True)),
    'pfr_net6_pointmass': dict(net=lambda: syn.random_network(6, n_io=1, seed=17, elements_per_model=2), **_cfg(
        'pfr', 'a', ['a -> 1 @ point(0.9, 0.9)'], [], '0, 4', '0.1, linear', None, P=3, first=1e-6)),
    'pfr_net8_nacl': dict(net=lambda: syn.random_network(8, n_io=2, seed=19), **_cfg(
        'pfr', 'NaCl, CaCO3, Na2CO3, CaCl2', ['NaCl -> 0', 'CaCO3 -> 0', 'Na2CO3 -> 0', 'CaCl2 -> 0'],
        ['NaCl: inlet -> 1', 'CaCO3: inlet -> 1', 'Na2CO3: inlet -> 0', 'CaCl2: inlet -> 0'],
        '0, 20', '0.5, linear', 'nacl', P=2)),
    # ---- C2: real network of examples/OpenFOAM/pipe_with_recirc (444 PFRs), tracer + RTD
    'openfoam_2d_pfr': dict(net=_openfoam_network, **_cfg(
        'pfr', 'a', ['a -> 0'], ['a: inlet -> 1'], '0, 300', '5.0, linear', None, P=2, rtol=1e-10, atol=1e-12, rtd=True),
        solvers=('LSODA',), centroids='fixture'),
    # ---- C3: the same mesh as a CSTR network (76 CSTRs), 2NaCl + CaCO3 <-> Na2CO3 + CaCl2 (BASELINE.json configs[2])
    'openfoam_2d_cstr_nacl': dict(net=_openfoam_cstr_network, **_cfg(
        'cstr', 'NaCl, CaCO3, Na2CO3, CaCl2', ['NaCl -> 0', 'CaCO3 -> 0', 'Na2CO3 -> 0', 'CaCl2 -> 0'],
        ['NaCl: inlet -> 1', 'CaCO3: inlet -> 1', 'Na2CO3: inlet -> 0', 'CaCl2: inlet -> 0'],
        '0, 300', '5.0, linear', 'nacl', rtol=1e-10, atol=1e-12), centroids='fixture'),
}


def n_elements(net: NetworkArrays) -> int:
    if net.model_to_element_map is None:
        return net.n_models
    return 1 + max(e for m in net.model_to_element_map for _, e in m)


def element_geometry(name: str, net: NetworkArrays):
    """Deterministic element centroids (2-D, in the unit square) and sizes for a synthetic case."""
    case = CASES[name]
    if case.get('centroids') == 'fixture':
        d = np.load(os.path.join(HERE, 'openfoam_2d_pfr_geometry.npz'))
        return d['centroids'], d['sizes']
    n_el = n_elements(net)
    rng = np.random.default_rng(abs(hash_name(name)))
    return rng.uniform(0.0, 1.0, (n_el, 2)), rng.uniform(0.5, 1.5, n_el)


def hash_name(name: str) -> int:
    h = 0
    for ch in name:
        h = (h * 131 + ord(ch)) % (2 ** 31)
    return h


def config_text(name: str, solver: str, reactions_path: str = 'None') -> str:
    c = CASES[name]
    pad = "\n" + " " * 22
    lines = [
        "[INPUT]", "openfoam_sol_folder_path = ./", "openfoam_sim_folder_to_use = 0", "no_flux_names = None",
        "ignored_names = None", 'domain_inlet_names = ("inlet",)', 'domain_outlet_names = ("outlet",)', "",
        "[COMPARTMENT MODELLING]", f"model = {c['model'].upper()}", "",
        "[SIMULATION]", "run = True", f"t_span = {c['t_span']}", f"t_eval = {c['t_eval']}",
        f"first_timestep = {c['first']}", f"points_per_pfr = {c['P']}", f"solver = {solver}",
        f"rtol = {c['rtol']}", f"atol = {c['atol']}", f"specie_names = {c['species']}",
        f"reactions_file_path = {reactions_path if c['rxn'] else 'None'}",
        "initial_conditions = " + pad.join(c['ic']),
    ]
    if c['bc']:
        lines.append("boundary_conditions = " + pad.join(c['bc']))
    lines += ["", "[POST-PROCESSING]", "inlet_bc_name = inlet", "outlet_bc_name = outlet",
              f"calculate_rtd = {c['rtd']}", ""]
    return "\n".join(lines)


def make_cmesh(name: str, net: NetworkArrays, grouped_bcs) -> SimpleNamespace:
    centroids, sizes = element_geometry(name, net)
    return SimpleNamespace(grouped_bcs=grouped_bcs, element_sizes=sizes, element_centroids=centroids)
