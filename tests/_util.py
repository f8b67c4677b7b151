"""Shared helpers for the tests: build (config parser, cmesh, network) for a golden case without the reference."""
from __future__ import annotations

import os

import numpy as np

import cases as C
from openccm_b200.config import ConfigParser, GroupedBCs

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_golden(name: str):
    return np.load(os.path.join(GOLDEN_DIR, f'{name}.npz'))


def setup_case(name: str, solver: str, workdir, **overrides):
    """Returns (cp, cmesh, net (NetworkArrays), model_network (5-tuple)); the reactions file is written to workdir."""
    case = C.CASES[name]
    rxn_path = 'None'
    if case['rxn']:
        rxn_path = os.path.join(str(workdir), f'reactions_{name}')
        with open(rxn_path, 'w') as fh:
            fh.write(C.RXN[case['rxn']])
    cp = ConfigParser(text=C.config_text(name, solver, rxn_path))
    cp['SETUP']['tmp_folder_path'] = os.path.join(str(workdir), 'cache') + '/'
    for k, v in overrides.items():
        cp['SIMULATION'][k] = str(v)
    net = case['net']()
    cmesh = C.make_cmesh(name, net, GroupedBCs(cp))
    return cp, cmesh, net, net.to_model_network()
