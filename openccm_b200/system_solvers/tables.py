"""
Mechanism tables: reactions file + boundary-condition strings -> packed device blob.

The reference turns these inputs into generated Python modules compiled by numba
(/root/reference/openccm/system_solvers/reactions.py:270-467,
 boundary_and_initial_conditions.py:182-245).  Here the same strings go through the same sympy parsing and
end up as *data* for the CUDA kernels (layout: openccm_b200/csrc/occm_tables.h):

* per-species stoichiometric terms  ``coeff * k_scale[rxn] * prod c[sp]^ord * (optional program)``;
* postfix byte code for anything that is not a monomial (Monod-type rate factors, time functions g(t)).

Reaction strings are assembled exactly like reactions.py:356-364 (``sign + coeff + '*' + rate + '*' + prod``,
no parentheses around the rate), one string per (species, reaction) so that per-reaction ensemble scaling stays
possible; summing them reproduces the reference's per-species expression.
"""
from __future__ import annotations

import configparser
import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import sympy as sp

# --------------------------------------------------------------------------------------------- op codes
# keep in sync with csrc/occm_tables.h
OPS = ['CONST', 'SPECIES', 'TIME', 'FIELD', 'ADD', 'SUB', 'MUL', 'DIV', 'POW', 'NEG', 'POWI',
       'SIN', 'COS', 'TAN', 'EXP', 'LOG', 'SQRT', 'ABS', 'TANH', 'SINH', 'COSH', 'ASIN', 'ACOS', 'ATAN', 'SIGN',
       'FLOOR', 'CEIL', 'MIN', 'MAX', 'LT', 'LE', 'GT', 'GE', 'EQ', 'NE', 'AND', 'OR', 'NOT']
OP = {name: i for i, name in enumerate(OPS)}
POWI_BIAS = 1 << 22
VM_STACK = 16
MAGIC = 0x4F43434D
H_WORDS = 32
(H_MAGIC, H_BYTES, H_S, H_NRXN, H_NTERM, H_NFAC, H_NPROG, H_NCODE, H_NCONST, H_NBC, H_NFIELD,
 H_OFF_TERM_PTR, H_OFF_TERM_COEFF, H_OFF_TERM_RXN, H_OFF_TERM_PROG, H_OFF_FAC_PTR, H_OFF_FAC_SP, H_OFF_FAC_ORD,
 H_OFF_PROG_PTR, H_OFF_CODE, H_OFF_CONST, H_OFF_BC_PROG, H_MAX_TERMS, H_MAX_SLOTS, H_ANY_PROG) = range(25)

_FUNCS = {sp.sin: 'SIN', sp.cos: 'COS', sp.tan: 'TAN', sp.exp: 'EXP', sp.log: 'LOG', sp.Abs: 'ABS', sp.tanh: 'TANH',
          sp.sinh: 'SINH', sp.cosh: 'COSH', sp.asin: 'ASIN', sp.acos: 'ACOS', sp.atan: 'ATAN', sp.sign: 'SIGN',
          sp.floor: 'FLOOR', sp.ceiling: 'CEIL'}
_RELS = {sp.StrictLessThan: 'LT', sp.LessThan: 'LE', sp.StrictGreaterThan: 'GT', sp.GreaterThan: 'GE',
         sp.Equality: 'EQ', sp.Unequality: 'NE'}

T_SYM = sp.Symbol('t')


class UnsupportedExpression(ValueError):
    """Raised for constructs the device interpreter cannot evaluate. There is no CPU fallback."""


class H(sp.Function):
    """Smoothed Heaviside of the reference (system_solvers/helper_functions.py:33-53)."""

    @classmethod
    def eval(cls, arg):
        return sp.Piecewise((0, arg < 0), (1, arg > 0.01), (1 / 2 * (1 - sp.cos(arg * np.pi / 0.01)), True))


# --------------------------------------------------------------------------------------------- compiler
@dataclass
class Program:
    code: List[int] = field(default_factory=list)       # op | arg << 8
    depth: int = 0
    max_depth: int = 0

    def emit(self, op: str, arg: int = 0, delta: int = 0):
        self.code.append(OP[op] | (arg << 8))
        self.depth += delta
        self.max_depth = max(self.max_depth, self.depth)


class Compiler:
    """sympy expression -> postfix program.  ``species``: name -> index, ``fields``: name -> index."""

    def __init__(self, species: Sequence[str] = (), fields: Sequence[str] = ()):
        self.species = {n: i for i, n in enumerate(species)}
        self.fields = {n: i for i, n in enumerate(fields)}
        self.consts: List[float] = []
        self._const_idx: Dict[float, int] = {}
        self.programs: List[List[int]] = []

    def const(self, v: float) -> int:
        v = float(v)
        key = v.hex()
        if key not in self._const_idx:
            self._const_idx[key] = len(self.consts)
            self.consts.append(v)
        return self._const_idx[key]

    def add_program(self, expr) -> int:
        prog = Program()
        self._compile(sp.sympify(expr), prog)
        if prog.max_depth > VM_STACK:
            raise UnsupportedExpression(f'expression too deep for the device stack ({prog.max_depth} > {VM_STACK}): {expr}')
        self.programs.append(prog.code)
        return len(self.programs) - 1

    # -- recursive descent over the sympy tree
    def _compile(self, e, p: Program) -> None:
        if e is sp.true or e is sp.false:
            p.emit('CONST', self.const(1.0 if e is sp.true else 0.0), +1)
        elif e.is_Number or e.is_NumberSymbol:
            p.emit('CONST', self.const(float(e)), +1)
        elif e.is_Symbol:
            name = str(e)
            if name == 't':
                p.emit('TIME', 0, +1)
            elif name in self.species:
                p.emit('SPECIES', self.species[name], +1)
            elif name in self.fields:
                p.emit('FIELD', self.fields[name], +1)
            else:
                raise UnsupportedExpression(f'unknown symbol {name!r}')
        elif e.is_Add:
            args = e.as_ordered_terms()
            self._compile(args[0], p)
            for a in args[1:]:
                self._compile(a, p)
                p.emit('ADD', 0, -1)
        elif e.is_Mul:
            args = e.as_ordered_factors()
            self._compile(args[0], p)
            for a in args[1:]:
                self._compile(a, p)
                p.emit('MUL', 0, -1)
        elif e.is_Pow:
            base, ex = e.args
            self._compile(base, p)
            if ex.is_Integer and abs(int(ex)) < POWI_BIAS:
                p.emit('POWI', int(ex) + POWI_BIAS)
            elif ex == sp.Rational(1, 2):
                p.emit('SQRT')
            else:
                self._compile(ex, p)
                p.emit('POW', 0, -1)
        elif isinstance(e, sp.Piecewise):
            # the reference's rewrite e1*(c1) + e2*(c2) + e3*(not (c1 | c2))   (boundary_and_initial_conditions.py:466-477)
            conds = [c for _, c in e.args if c is not sp.true]
            for k, (ex, c) in enumerate(e.args):
                self._compile(ex, p)
                self._compile(sp.Not(sp.Or(*conds)) if c is sp.true else c, p)
                p.emit('MUL', 0, -1)
                if k:
                    p.emit('ADD', 0, -1)
        elif isinstance(e, (sp.And, sp.Or)):
            op = 'AND' if isinstance(e, sp.And) else 'OR'
            self._compile(e.args[0], p)
            for a in e.args[1:]:
                self._compile(a, p)
                p.emit(op, 0, -1)
        elif isinstance(e, sp.Not):
            self._compile(e.args[0], p)
            p.emit('NOT')
        elif type(e) in _RELS:
            self._compile(e.args[0], p)
            self._compile(e.args[1], p)
            p.emit(_RELS[type(e)], 0, -1)
        elif isinstance(e, (sp.Min, sp.Max)):
            op = 'MIN' if isinstance(e, sp.Min) else 'MAX'
            self._compile(e.args[0], p)
            for a in e.args[1:]:
                self._compile(a, p)
                p.emit(op, 0, -1)
        elif isinstance(e, sp.Heaviside):          # sympy's sharp step (from d/dt of Min/Max/Abs): (x > 0) + 0.5 (x == 0)
            zero = self.const(0.0)
            self._compile(e.args[0], p); p.emit('CONST', zero, +1); p.emit('GT', 0, -1)
            self._compile(e.args[0], p); p.emit('CONST', zero, +1); p.emit('EQ', 0, -1)
            p.emit('CONST', self.const(0.5), +1); p.emit('MUL', 0, -1)
            p.emit('ADD', 0, -1)
        elif e.func in _FUNCS:
            self._compile(e.args[0], p)
            p.emit(_FUNCS[e.func])
        else:
            raise UnsupportedExpression(f'cannot run {e.func} on the device: {e}')


def run_program(code: Sequence[int], consts: Sequence[float], species: Sequence[float] = (), t: float = 0.0,
                fields: Sequence[float] = ()) -> float:
    """Host interpreter with the device semantics (used by the CPU tests and for c0 = BC(t0))."""
    import math
    st: List[float] = []
    for ins in code:
        op, arg = OPS[ins & 0xff], ins >> 8
        if op == 'CONST': st.append(consts[arg])
        elif op == 'SPECIES': st.append(float(species[arg]))
        elif op == 'TIME': st.append(t)
        elif op == 'FIELD': st.append(float(fields[arg]))
        elif op in ('ADD', 'SUB', 'MUL', 'DIV', 'POW', 'MIN', 'MAX', 'LT', 'LE', 'GT', 'GE', 'EQ', 'NE', 'AND', 'OR'):
            b = st.pop(); a = st.pop()
            st.append({'ADD': lambda: a + b, 'SUB': lambda: a - b, 'MUL': lambda: a * b, 'DIV': lambda: a / b,
                       'POW': lambda: a ** b, 'MIN': lambda: min(a, b), 'MAX': lambda: max(a, b),
                       'LT': lambda: float(a < b), 'LE': lambda: float(a <= b), 'GT': lambda: float(a > b),
                       'GE': lambda: float(a >= b), 'EQ': lambda: float(a == b), 'NE': lambda: float(a != b),
                       'AND': lambda: float(a != 0 and b != 0), 'OR': lambda: float(a != 0 or b != 0)}[op]())
        elif op == 'NEG': st[-1] = -st[-1]
        elif op == 'NOT': st[-1] = float(st[-1] == 0)
        elif op == 'POWI': st[-1] = st[-1] ** (arg - POWI_BIAS)
        elif op == 'SIGN': st[-1] = float((st[-1] > 0) - (st[-1] < 0))
        elif op == 'ABS': st[-1] = abs(st[-1])
        elif op == 'CEIL': st[-1] = float(math.ceil(st[-1]))
        elif op == 'FLOOR': st[-1] = float(math.floor(st[-1]))
        else: st[-1] = getattr(math, op.lower())(st[-1])
    return st[-1] if st else 0.0


# --------------------------------------------------------------------------------------------- reactions
_TERM = re.compile(r'^\s*(\d[\d.]*)?\s*((?:[A-Za-z]+(?:\d[\d.]*)?)+)\s*$')


def _parse_side(side: str) -> Dict[str, str]:
    """``2NaCl + CaCO3`` -> {'NaCl': '2', 'CaCO3': '1'} (grammar of reactions.py:367-379; coefficient optional)."""
    out: Dict[str, str] = {}
    if not side.strip():
        return out
    for term in side.split('+'):
        m = _TERM.match(term)
        if m is None:
            raise ValueError(f'cannot parse reaction term {term!r}')
        out[m.group(2)] = m.group(1) or '1'
    return out


def read_reactions_file(path: str) -> Tuple[Dict[str, str], Dict[str, str], Dict[str, str]]:
    """[REACTIONS], [RATE CONSTANTS], [EXTRA TERMS] with case-sensitive keys (reactions.py:79-85)."""
    p = configparser.ConfigParser()
    p.optionxform = str
    if not p.read(path):
        raise FileNotFoundError(path)
    sec = lambda n: dict(p[n]) if p.has_section(n) else {}
    return sec('REACTIONS'), sec('RATE CONSTANTS'), sec('EXTRA TERMS')


def check_reaction_ids(reactions: Dict[str, str], rates: Dict[str, str]) -> None:
    """reactions.py:252-264."""
    if len(reactions) != len(rates):
        raise RuntimeError("Number of reactions does not match number of rates (or their reaction IDs are different). "
                           "Check reactions configuration file.")
    if sorted(reactions) != sorted(rates):
        raise RuntimeError("Cannot find rate for one or more reactions. Check that reaction and rate IDs align.")


def reaction_term_strings(reactions: Dict[str, str], rates: Dict[str, str]) -> List[Tuple[str, int, str]]:
    """[(specie, reaction index, term string)] — the pieces reactions.py:356-364 concatenates per specie."""
    check_reaction_ids(reactions, rates)
    out: List[Tuple[str, int, str]] = []
    for r, (rid, eqn) in enumerate(reactions.items()):
        if '->' not in eqn:
            raise ValueError(f'reaction {rid!r} has no "->"')
        lhs, rhs = eqn.split('->')
        reactants, products = _parse_side(lhs), _parse_side(rhs)
        k = rates[rid]
        for side, sign in ((reactants, '-'), (products, '+')):
            for specie, coeff in side.items():
                if k[:2] == 'z=':
                    out.append((specie, r, sign + coeff + '*' + k[2:]))
                else:
                    out.append((specie, r, sign + coeff + '*' + k + '*' + '*'.join(s + '**' + c for s, c in reactants.items())))
    return out


@dataclass
class Term:
    specie: int
    rxn: int
    coeff: float
    factors: List[Tuple[int, int]]     # (species index, positive integer order)
    prog: int = -1


def _split_monomial(expr, species_syms: Dict[sp.Symbol, int]):
    """expr -> (numeric coeff, [(species, order)], residual expression or None)."""
    coeff, rest = expr.as_coeff_Mul()
    factors: List[Tuple[int, int]] = []
    residual = []
    for f in sp.Mul.make_args(rest):
        if f.is_Symbol and f in species_syms:
            factors.append((species_syms[f], 1))
        elif f.is_Pow and f.base.is_Symbol and f.base in species_syms and f.exp.is_Integer and int(f.exp) > 0:
            factors.append((species_syms[f.base], int(f.exp)))
        elif f == 1:
            continue
        else:
            residual.append(f)
    return float(coeff), factors, (sp.Mul(*residual) if residual else None)


def extra_term_values(extra: Dict[str, str]) -> Tuple[Dict[str, sp.Expr], List[str]]:
    """Scalar [EXTRA TERMS] evaluated in file order with earlier names visible (the generated module executes them
    as ``name = expression`` lines, reactions.py:215-233); ``CFD, file`` entries become per-point fields."""
    if 'CFD' in extra:
        raise KeyError("'CFD' is reserved and cannot be used as the name for an extra variable, "
                       "please use a different name, or change the capitalization.")
    values: Dict[str, sp.Expr] = {}
    fields: List[str] = []
    for name, expr in extra.items():
        if expr[:3] == 'CFD':
            fields.append(name)
        else:
            values[name] = sp.parse_expr(expr, local_dict={**values, 'np': sp}).subs(values)
    return values, fields


class MechanismTables:
    """Builds and packs the blob. Usage: add_reactions(...), add_bc_group(...) ..., pack()."""

    def __init__(self, species: Sequence[str], fields: Sequence[str] = ()):
        self.species = list(species)
        self.S = len(self.species)
        self.fields = list(fields)
        self.compiler = Compiler(self.species, self.fields)
        self.terms: List[Term] = []
        self.n_rxn = 0
        self.bc_groups: List[List[int]] = []      # per group: program id per species (or -1)
        self.bc_exprs: List[Dict[int, sp.Expr]] = []

    # ---- reactions
    def add_reactions(self, reactions: Dict[str, str], rates: Dict[str, str], extra: Optional[Dict[str, str]] = None) -> None:
        if not reactions:
            return
        values, fields = extra_term_values(extra or {})
        for f in fields:
            if f not in self.compiler.fields:
                raise UnsupportedExpression(f'per-point extra term {f!r} needs its array (pass fields=...)')
        syms = {sp.Symbol(n): i for i, n in enumerate(self.species)}
        pieces = reaction_term_strings(reactions, rates)
        named = {s for s, _, _ in pieces}
        for s in self.species:
            if s not in named:
                raise KeyError(s)                       # reactions.py:416 (every specie must appear in some reaction)
        self.n_rxn = len(reactions)
        for specie, rxn, text in pieces:
            if specie not in self.compiler.species:
                continue                                # species not simulated: the reference never reads its entry
            expr = sp.parse_expr(text).subs(values)
            for addend in sp.Add.make_args(sp.expand(expr) if expr.is_Add else expr):
                coeff, factors, residual = _split_monomial(addend, syms)
                if coeff == 0.0:
                    continue
                prog = -1
                if residual is not None:
                    free = {str(x) for x in residual.free_symbols}
                    unknown = free - set(self.species) - set(self.fields)
                    if unknown:
                        raise NameError(f'undefined names in rate expression {text!r}: {sorted(unknown)}')
                    prog = self.compiler.add_program(residual)
                self.terms.append(Term(self.compiler.species[specie], rxn, coeff, factors, prog))

    # ---- boundary / pulse time functions
    def add_bc_group(self, exprs: Dict[int, sp.Expr]) -> int:
        """exprs: species index -> g(t).  Returns the group id used in row_src (= -1 - id) and pulse_fn."""
        progs = [-1] * self.S
        for s, e in exprs.items():
            e = sp.sympify(e)
            free = e.free_symbols
            if free - {T_SYM}:
                raise ValueError(f"Boundary condition {e} uses a variable other than t (time).")
            progs[s] = self.compiler.add_program(e)
        self.bc_groups.append(progs)
        self.bc_exprs.append(dict(exprs))
        return len(self.bc_groups) - 1

    def eval_bc(self, group: int, s: int, t: float) -> float:
        prog = self.bc_groups[group][s]
        return 0.0 if prog < 0 else run_program(self.compiler.programs[prog], self.compiler.consts, t=t)

    def reaction_rates_host(self, c_row: Sequence[float], k_scale: Optional[Sequence[float]] = None,
                            fields: Sequence[float] = ()) -> np.ndarray:
        """Host evaluation of the term tables for one point (CPU tests of the table builder)."""
        out = np.zeros(self.S)
        for tm in self.terms:
            v = tm.coeff * (k_scale[tm.rxn] if k_scale is not None else 1.0)
            for spi, order in tm.factors:
                v *= float(c_row[spi]) ** order
            if tm.prog >= 0:
                v *= run_program(self.compiler.programs[tm.prog], self.compiler.consts, c_row, 0.0, fields)
            out[tm.specie] += v
        return out

    # ---- packing
    def pack(self) -> np.ndarray:
        S = self.S
        terms = sorted(self.terms, key=lambda tm: tm.specie)          # stable: keeps reaction order within a specie
        term_ptr = np.zeros(S + 1, dtype=np.int32)
        for tm in terms:
            term_ptr[tm.specie + 1] += 1
        term_ptr = np.cumsum(term_ptr).astype(np.int32)
        term_coeff = np.array([tm.coeff for tm in terms], dtype=np.float64)
        term_rxn = np.array([tm.rxn for tm in terms], dtype=np.int32)
        term_prog = np.array([tm.prog for tm in terms], dtype=np.int32)
        fac_ptr = np.zeros(len(terms) + 1, dtype=np.int32)
        fac_sp: List[int] = []; fac_ord: List[int] = []
        for i, tm in enumerate(terms):
            for spi, order in tm.factors:
                fac_sp.append(spi); fac_ord.append(order)
            fac_ptr[i + 1] = len(fac_sp)
        progs = self.compiler.programs
        prog_ptr = np.zeros(len(progs) + 1, dtype=np.int32)
        code: List[int] = []
        for i, pr in enumerate(progs):
            code.extend(pr)
            prog_ptr[i + 1] = len(code)
        consts = np.array(self.compiler.consts, dtype=np.float64)
        bc_prog = np.array([p for g in self.bc_groups for p in g], dtype=np.int32)

        sections = [
            (H_OFF_TERM_PTR, term_ptr), (H_OFF_TERM_COEFF, term_coeff), (H_OFF_TERM_RXN, term_rxn),
            (H_OFF_TERM_PROG, term_prog), (H_OFF_FAC_PTR, fac_ptr), (H_OFF_FAC_SP, np.array(fac_sp, dtype=np.int32)),
            (H_OFF_FAC_ORD, np.array(fac_ord, dtype=np.int32)), (H_OFF_PROG_PTR, prog_ptr),
            (H_OFF_CODE, np.array(code, dtype=np.int32)), (H_OFF_CONST, consts), (H_OFF_BC_PROG, bc_prog),
        ]
        header = np.zeros(H_WORDS, dtype=np.int32)
        header[[H_MAGIC, H_S, H_NRXN, H_NTERM, H_NFAC, H_NPROG, H_NCODE, H_NCONST, H_NBC, H_NFIELD]] = [
            MAGIC, S, self.n_rxn, len(terms), len(fac_sp), len(progs), len(code), len(consts), len(self.bc_groups),
            len(self.fields)]
        header[H_MAX_TERMS] = int(np.diff(term_ptr).max()) if len(terms) else 0
        header[H_MAX_SLOTS] = max((sum(o for _, o in tm.factors) for tm in terms), default=0)
        header[H_ANY_PROG] = int(any(tm.prog >= 0 for tm in terms))
        chunks = [header.tobytes()]
        off = header.nbytes
        for slot, arr in sections:
            header_off = off
            raw = arr.tobytes()
            pad = (-len(raw)) % 8
            chunks.append(raw + b'\0' * pad)
            header[slot] = header_off
            off += len(raw) + pad
        header[H_BYTES] = off
        chunks[0] = header.tobytes()
        blob = np.frombuffer(b''.join(chunks), dtype=np.int32).copy()
        assert blob.nbytes == off and off % 8 == 0
        return blob
