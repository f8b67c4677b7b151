"""
Named synthetic workloads of BASELINE.json: deterministic inputs shared by the device path,
the oracle and the benchmarks.  Each builder returns (model, NetworkArrays, ConfigParser, cmesh) — exactly the
arguments of ``solve_system`` — so that every consumer goes through the same reference-shaped boundary.
"""
from __future__ import annotations

import os
import tempfile
from types import SimpleNamespace
from typing import Optional, Tuple

import numpy as np

from . import synthetic as syn
from .config import ConfigParser, GroupedBCs


def _config(model: str, species, rxn_path: str, *, t_span, t_eval: str, P: int = 2, rtol=1e-6, atol=1e-10,
            first=1e-4, solver='B200-RK45', ic_value=1.0, bc_values=None) -> ConfigParser:
    pad = "\n" + " " * 22
    bc_values = bc_values if bc_values is not None else [1.0 if i % 2 == 0 else 0.5 for i in range(len(species))]
    text = "\n".join([
        "[INPUT]", "openfoam_sol_folder_path = ./", "openfoam_sim_folder_to_use = 0", "no_flux_names = None",
        "ignored_names = None", 'domain_inlet_names = ("inlet",)', 'domain_outlet_names = ("outlet",)', "",
        "[COMPARTMENT MODELLING]", f"model = {model.upper()}", "",
        "[SIMULATION]", "run = True", f"t_span = {t_span[0]}, {t_span[1]}", f"t_eval = {t_eval}",
        f"first_timestep = {first}", f"points_per_pfr = {P}", f"solver = {solver}", f"rtol = {rtol}", f"atol = {atol}",
        "specie_names = " + ", ".join(species), f"reactions_file_path = {rxn_path}",
        "initial_conditions = " + pad.join(f"{s} -> {ic_value}" for s in species),
        "boundary_conditions = " + pad.join(f"{s}: inlet -> {v}" for s, v in zip(species, bc_values)),
        "", "[POST-PROCESSING]", "inlet_bc_name = inlet", "outlet_bc_name = outlet", "calculate_rtd = False", ""])
    return ConfigParser(text=text)


def _cmesh(cp, n_elements: int, seed: int) -> SimpleNamespace:
    rng = np.random.default_rng(seed)
    return SimpleNamespace(grouped_bcs=GroupedBCs(cp), element_sizes=np.ones(n_elements),
                           element_centroids=rng.uniform(0.0, 1.0, (n_elements, 2)))


def cstr_ensemble_workload(n_cstr: int = 100_000, n_species: int = 10, n_reactions: int = 5, n_io: int = 16,
                           t_end: float = 10.0, n_out: int = 101, seed: int = 0, workdir: Optional[str] = None,
                           rtol: float = 1e-6, atol: float = 1e-10, solver: str = 'B200-RK45'):
    """C4: synthetic N-CSTR network (ring + random permutation edges, exactly balanced flows), S species,
    second-order mechanism with k ~ logU(1e-2, 1e1), t in [0, t_end], n_out linear outputs."""
    workdir = workdir or tempfile.mkdtemp(prefix='occm_c4_')
    net = syn.random_network(n_cstr, n_io=n_io, seed=seed)
    species, rxn_text = syn.second_order_mechanism(n_species, n_reactions, seed=seed)
    rxn_path = os.path.join(workdir, 'reactions_c4')
    with open(rxn_path, 'w') as fh:
        fh.write(rxn_text)
    dt = t_end / (n_out - 1)
    cp = _config('cstr', species, rxn_path, t_span=(0, t_end), t_eval=f"{dt!r}, linear", rtol=rtol, atol=atol, solver=solver)
    return 'cstr', net, cp, _cmesh(cp, n_cstr, seed + 100)


def pfr_stiff_workload(n_chains: int = 500, chain_len: int = 20, P: int = 200, n_species: int = 12,
                       n_reactions: int = 20, t_end: float = 20.0, n_out: int = 41, seed: int = 0,
                       workdir: Optional[str] = None, rtol: float = 1e-6, atol: float = 1e-10, solver: str = 'B200-BDF'):
    """C5: Np = n_chains * chain_len PFRs in recycle chains, P FD points each, stiff 20-reaction mechanism."""
    workdir = workdir or tempfile.mkdtemp(prefix='occm_c5_')
    net = syn.pfr_chains(n_chains, chain_len, seed=seed)
    species, rxn_text = syn.stiff_mechanism(n_species, n_reactions, seed=seed)
    rxn_path = os.path.join(workdir, 'reactions_c5')
    with open(rxn_path, 'w') as fh:
        fh.write(rxn_text)
    dt = t_end / (n_out - 1)
    cp = _config('pfr', species, rxn_path, t_span=(0, t_end), t_eval=f"{dt!r}, linear", P=P, rtol=rtol, atol=atol,
                 solver=solver, ic_value=0.0)
    return 'pfr', net, cp, _cmesh(cp, n_chains * chain_len, seed + 100)
