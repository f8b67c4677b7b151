"""
Sparsity pattern of the ODE Jacobian in the REFERENCE state layout (y[s * n_points + p]) — SURVEY.md §8(f) item 4.

The reference calls ``solve_ivp`` without ``jac`` / ``jac_sparsity`` (cstr_system.py:151, pfr_system.py:191), so scipy builds a dense
finite-difference Jacobian: N_dof + 1 RHS calls per Jacobian and a dense LU (97 % of LSODA's RHS calls on the real 444-PFR
network).  Passing the pattern below as ``jac_sparsity=`` to scipy's BDF / Radau groups the columns and cuts that to a
handful of RHS calls and a sparse LU — a fairer CPU baseline, and useful to users who stay on the CPU path.

J = (I + W) (A_transport (x) I_S + blockdiag(dr/dc)): transport couples equal species of connected points, reactions couple the
species of one point, the PFR inlet coupling W adds the rows of the upstream outlet nodes to the inlet-node rows.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .export import HostSystem


def species_coupling(host: HostSystem) -> np.ndarray:
    """Boolean [S, S]: entry (i, j) set when d r_i / d c_j can be non-zero."""
    S = host.S
    out = np.eye(S, dtype=bool)
    for tm in host.tables.terms:
        for spj, _ in tm.factors:
            out[tm.specie, spj] = True
        if tm.prog >= 0:                # byte-code factor: may depend on any species
            out[tm.specie, :] = True
    return out


def jacobian_sparsity(host: HostSystem) -> sp.csr_matrix:
    S, Np = host.S, host.n_points
    rows_m = np.repeat(np.arange(host.n_models), np.diff(host.row_ptr))
    src = host.row_src
    internal = src >= 0
    if host.model == 'cstr':
        point_pat = sp.coo_matrix((np.ones(internal.sum(), dtype=bool), (rows_m[internal], src[internal])), shape=(Np, Np)).tocsr()
        point_pat = point_pat + sp.identity(Np, dtype=bool, format='csr')
        inlet_rows = None
    else:
        P = host.P
        p = np.arange(Np)
        interior = p % P != 0
        r = np.concatenate([p[interior], p[interior]]); c = np.concatenate([p[interior], p[interior] - 1])
        point_pat = sp.coo_matrix((np.ones(len(r), dtype=bool), (r, c)), shape=(Np, Np)).tocsr()
        inlet_rows = (rows_m[internal] * P, (src[internal] + 1) * P - 1)       # inlet node <- outlet node of the upstream PFR
    react = sp.csr_matrix(species_coupling(host))
    eyeS = sp.identity(S, dtype=bool, format='csr')
    # reference layout: index = s * Np + p  ->  kron(species block, point block)
    J = sp.kron(eyeS, point_pat, format='csr').astype(bool)
    diag_pts = sp.identity(Np, dtype=bool, format='csr')
    if host.model != 'cstr':                   # no reaction on PFR inlet nodes (pfr_system.py:308, reactions.py:436-447)
        d = np.ones(Np, dtype=bool); d[np.arange(0, Np, host.P)] = False
        diag_pts = sp.diags(d.astype(np.int8), format='csr').astype(bool)
    J = (J + sp.kron(react, diag_pts, format='csr').astype(bool)).astype(bool)
    if inlet_rows is not None and len(inlet_rows[0]):
        W = sp.coo_matrix((np.ones(len(inlet_rows[0]), dtype=np.int8), inlet_rows), shape=(Np, Np)).tocsr()
        Wfull = sp.kron(sp.identity(S, dtype=np.int8, format='csr'), W, format='csr')
        J = (J.astype(np.int8) + (Wfull @ J.astype(np.int8))).astype(bool)
    J.eliminate_zeros()
    return J.tocsr()
