"""
Deterministic synthetic compartment networks and mechanisms.

Every generator is a pure function of its arguments and a numpy ``default_rng(seed)``;
the oracle, the golden-vector script and the device path all consume the same
``NetworkArrays`` so their inputs are identical by construction.

Flow construction (exact node balance by construction, no LP needed — the reference gets
its balance from ``tweak_final_flows``, compartment_models/helpers.py:268-438):

* a ring 0 -> 1 -> ... -> n-1 -> 0 carrying a base circulation,
* the edges i -> pi(i) of a random permutation, with one flow value per permutation cycle
  (every node then gains and loses the same amount),
* ``n_io`` (inlet, outlet) pairs; the through-flow entering at an inlet travels along the
  ring to its outlet (added with a difference array + cumulative sum).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np

from .network import NetworkArrays

INLET_ID = -1   # GroupedBCs.id('inlet')  with domain_inlet_names=("inlet",), domain_outlet_names=("outlet",)
OUTLET_ID = -2  # GroupedBCs.id('outlet')   (mesh/cmesh.py:108-111: position in inlet names + outlet names)


def single_model(volume: float = 1.0, flow: float = 0.1) -> NetworkArrays:
    """One CSTR/PFR with one inlet and one outlet: the `examples/simple_reactors` geometry
    (input_data_generation/input_sim_result_generator.py:35-36 -> V=1, Q=0.1, tau=10)."""
    i64 = lambda *v: np.array(v, dtype=np.int64)
    return NetworkArrays(1, np.array([volume]), np.array([flow, flow]),
                         i64(0), i64(0), i64(INLET_ID), i64(0), i64(1), i64(OUTLET_ID),
                         [[(0.0, 0)]])


def random_network(n: int, n_io: int = 1, seed: int = 0, *, through_flow: float = 1.0,
                   ring_flow: Tuple[float, float] = (0.25, 0.75), recycle_flow: Tuple[float, float] = (0.0, 0.5),
                   volume: Tuple[float, float] = (0.5, 2.0), permutation_edges: bool = True,
                   elements_per_model: int = 1) -> NetworkArrays:
    """Strongly connected random digraph with exactly balanced flows (see module docstring)."""
    rng = np.random.default_rng(seed)
    assert n >= 2 and 1 <= n_io <= n // 2
    volumes = rng.uniform(volume[0], volume[1], n)

    # ring edges e=i : i -> (i+1)%n
    ring = np.full(n, rng.uniform(*ring_flow))
    # through-flow: inlet node a_k -> ... -> outlet node b_k along the ring
    io_nodes = rng.choice(n, size=2 * n_io, replace=False)
    inl, outl = io_nodes[:n_io], io_nodes[n_io:]
    q_io = np.full(n_io, through_flow / n_io)
    diff = np.zeros(n + 1)
    for a, b, q in zip(inl.tolist(), outl.tolist(), q_io.tolist()):
        # ring edges a, a+1, ..., b-1 (mod n)
        if a < b:
            diff[a] += q; diff[b] -= q
        else:
            diff[a] += q; diff[n] -= q; diff[0] += q; diff[b] -= q
    ring = ring + np.cumsum(diff)[:n]

    srcs = [np.arange(n)]; dsts = [(np.arange(n) + 1) % n]; fl = [ring]
    if permutation_edges:
        perm = rng.permutation(n)
        # label permutation cycles
        label = np.full(n, -1, dtype=np.int64)
        n_cycles = 0
        for start in range(n):
            if label[start] >= 0:
                continue
            j = start
            while label[j] < 0:
                label[j] = n_cycles
                j = perm[j]
            n_cycles += 1
        cyc_flow = rng.uniform(recycle_flow[0], recycle_flow[1], n_cycles)
        keep = perm != np.arange(n)
        srcs.append(np.arange(n)[keep]); dsts.append(perm[keep]); fl.append(cyc_flow[label][keep])
    src = np.concatenate(srcs); dst = np.concatenate(dsts); flow_int = np.concatenate(fl)
    n_int = src.size

    # connection ids: internal edges first, then inlet links, then outlet links
    flows = np.concatenate([flow_int, q_io, q_io])
    conn_int = np.arange(n_int)
    conn_in = n_int + np.arange(n_io)
    conn_out = n_int + n_io + np.arange(n_io)

    # inlet-side entries (model = destination), grouped by model in ascending order (stable)
    in_model = np.concatenate([dst, inl]); in_conn = np.concatenate([conn_int, conn_in])
    in_other = np.concatenate([src, np.full(n_io, INLET_ID)])
    o = np.argsort(in_model, kind='stable')
    in_model, in_conn, in_other = in_model[o], in_conn[o], in_other[o]
    out_model = np.concatenate([src, outl]); out_conn = np.concatenate([conn_int, conn_out])
    out_other = np.concatenate([dst, np.full(n_io, OUTLET_ID)])
    o = np.argsort(out_model, kind='stable')
    out_model, out_conn, out_other = out_model[o], out_conn[o], out_other[o]

    if elements_per_model == 1:
        m2e = None
    else:
        m2e = [[((k + 0.5) / elements_per_model, i * elements_per_model + k) for k in range(elements_per_model)]
               for i in range(n)]
    i64 = lambda x: np.ascontiguousarray(x, dtype=np.int64)
    return NetworkArrays(n, volumes, flows, i64(in_model), i64(in_conn), i64(in_other),
                         i64(out_model), i64(out_conn), i64(out_other), m2e)


def pfr_chains(n_chains: int, chain_len: int, seed: int = 0, *, recycle_fraction: float = 0.3,
               volume: Tuple[float, float] = (0.5, 2.0), elements_per_model: int = 1) -> NetworkArrays:
    """C5 topology: ``n_chains`` chains of ``chain_len`` PFRs in series; each chain is fed from the domain
    inlet, discharges to the domain outlet and recycles a fraction of its outlet stream to its own head."""
    rng = np.random.default_rng(seed)
    n = n_chains * chain_len
    volumes = rng.uniform(volume[0], volume[1], n)
    q_feed = rng.uniform(0.5, 1.5, n_chains)
    q_rec = recycle_fraction * q_feed
    head = np.arange(n_chains) * chain_len
    tail = head + chain_len - 1
    # series edges inside each chain carry feed + recycle
    s_src = np.setdiff1d(np.arange(n), tail); s_dst = s_src + 1
    s_flow = np.repeat(q_feed + q_rec, chain_len - 1)
    src = np.concatenate([s_src, tail]); dst = np.concatenate([s_dst, head]); flow_int = np.concatenate([s_flow, q_rec])
    n_int = src.size
    flows = np.concatenate([flow_int, q_feed, q_feed])
    conn_int = np.arange(n_int); conn_in = n_int + np.arange(n_chains); conn_out = n_int + n_chains + np.arange(n_chains)
    in_model = np.concatenate([dst, head]); in_conn = np.concatenate([conn_int, conn_in])
    in_other = np.concatenate([src, np.full(n_chains, INLET_ID)])
    o = np.argsort(in_model, kind='stable'); in_model, in_conn, in_other = in_model[o], in_conn[o], in_other[o]
    out_model = np.concatenate([src, tail]); out_conn = np.concatenate([conn_int, conn_out])
    out_other = np.concatenate([dst, np.full(n_chains, OUTLET_ID)])
    o = np.argsort(out_model, kind='stable'); out_model, out_conn, out_other = out_model[o], out_conn[o], out_other[o]
    m2e = None if elements_per_model == 1 else \
        [[((k + 0.5) / elements_per_model, i * elements_per_model + k) for k in range(elements_per_model)] for i in range(n)]
    i64 = lambda x: np.ascontiguousarray(x, dtype=np.int64)
    return NetworkArrays(n, volumes, flows, i64(in_model), i64(in_conn), i64(in_other),
                         i64(out_model), i64(out_conn), i64(out_other), m2e)


# ---------------------------------------------------------------------------------------------- mechanisms
def species_names(n: int) -> List[str]:
    """Alphabetic species names that avoid the reference's reserved symbols x, y, z, t, S
    (config_functions/expanded_config_parser.py:153) and sympy's single-letter constants (E, I, N, O, Q)."""
    pool = [c for c in "abcdefghijklmnopqruvw"]
    if n <= len(pool):
        return pool[:n]
    return [f"sp{chr(ord('a') + i // 26)}{chr(ord('a') + i % 26)}" for i in range(n)]


def second_order_mechanism(n_species: int = 10, n_reactions: int = 5, seed: int = 0,
                           log10_k: Tuple[float, float] = (-2.0, 1.0)) -> Tuple[List[str], str]:
    """C4 mechanism: ``A + B -> C`` type reactions over ``n_species`` species, k ~ logU(1e-2, 1e1).
    Returned as (species names, text of a reactions file in the reference's format,
    /root/reference/CONFIG_REACTIONS:1-9) so the oracle and the device parse the same text."""
    rng = np.random.default_rng(seed)
    names = species_names(n_species)
    lines_r, lines_k = ["[REACTIONS]"], ["[RATE CONSTANTS]"]
    used = set()
    for r in range(n_reactions):
        a, b, c = rng.choice(n_species, size=3, replace=False)
        used.update((a, b, c))
        lines_r.append(f"R{r + 1}: {names[a]} + {names[b]} -> {names[c]}")
        lines_k.append(f"R{r + 1}: {10 ** rng.uniform(*log10_k):.6e}")
    # every species must appear in some reaction (reactions.py:416 raises KeyError otherwise)
    r = n_reactions
    for s in range(n_species):
        if s not in used:
            r += 1
            tgt = (s + 1) % n_species
            lines_r.append(f"R{r}: {names[s]} -> {names[tgt]}")
            lines_k.append(f"R{r}: {10 ** rng.uniform(*log10_k):.6e}")
            used.update((s, tgt))
    return names, "\n".join(lines_r) + "\n\n" + "\n".join(lines_k) + "\n"


def stiff_mechanism(n_species: int = 12, n_reactions: int = 20, seed: int = 0,
                    log10_k: Tuple[float, float] = (-3.0, 4.0)) -> Tuple[List[str], str]:
    """C5 mechanism: mixed first/second order reactions with k spanning 1e-3 .. 1e4 (stiff)."""
    rng = np.random.default_rng(seed)
    names = species_names(n_species)
    lines_r, lines_k = ["[REACTIONS]"], ["[RATE CONSTANTS]"]
    used = set()
    for r in range(n_reactions):
        if r % 2 == 0:
            a, c = rng.choice(n_species, size=2, replace=False)
            lines_r.append(f"R{r + 1}: {names[a]} -> {names[c]}"); used.update((a, c))
        else:
            a, b, c = rng.choice(n_species, size=3, replace=False)
            lines_r.append(f"R{r + 1}: {names[a]} + {names[b]} -> {names[c]}"); used.update((a, b, c))
        lines_k.append(f"R{r + 1}: {10 ** rng.uniform(*log10_k):.6e}")
    r = n_reactions
    for s in range(n_species):
        if s not in used:
            r += 1
            lines_r.append(f"R{r}: {names[s]} -> {names[(s + 1) % n_species]}")
            lines_k.append(f"R{r}: {10 ** rng.uniform(*log10_k):.6e}")
    return names, "\n".join(lines_r) + "\n\n" + "\n".join(lines_k) + "\n"


def ensemble_sweeps(n_members: int, n_reactions: int, seeds: Tuple[int, int, int] = (1, 2, 3)):
    """C4 ensemble: rate-constant scale 10^U(-0.5,0.5) per (member, reaction), inflow scale U(0.5,2),
    initial-condition scale U(0,1)."""
    k_scale = 10.0 ** np.random.default_rng(seeds[0]).uniform(-0.5, 0.5, (n_members, n_reactions))
    q_scale = np.random.default_rng(seeds[1]).uniform(0.5, 2.0, n_members)
    ic_scale = np.random.default_rng(seeds[2]).uniform(0.0, 1.0, n_members)
    return k_scale, q_scale, ic_scale
