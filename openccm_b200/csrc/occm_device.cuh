// Device-side building blocks shared by the RHS, RK45 and BDF kernels (sm_100a).
//
// Thread mapping used everywhere: one thread per flat state element e = point * S + species of one ensemble
// member, consecutive threads = consecutive addresses (species innermost), so every own-element load/store of
// y, K_i, ... is a coalesced 8-byte-per-lane access.  A CTA works on "sub-chunks" of tc = OCCM_BLOCK / S whole
// points (tc * S <= OCCM_BLOCK elements), E of them per loop trip for memory-level parallelism; because a
// sub-chunk is a multiple of S elements, thread tid always owns species tid % S and keeps that species'
// stoichiometric terms in registers.  The stage state of the sub-chunk's points is staged in shared memory so that
// the reaction terms (which couple the S species of a point) and the PFR upwind difference read it from there.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "occm_tables.h"

#define OCCM_BLOCK 256
#define OCCM_MAX_S 64
#ifndef OCCM_FAST_BLOCK
#define OCCM_FAST_BLOCK 256        // threads per CTA of the fast kernels (the fix-up kernel keeps OCCM_BLOCK)
#endif

struct OccmSysDev {
    int model, S, n_models, P;
    long long n_points, ndof;
    int tc;                     // points per sub-chunk = OCCM_BLOCK / S (every thread keeps one species)
    int ell_k;                  // CSTR: ELL width (0 = CSR only)
    const int* ell_src; const double* ell_w;   // column-major [ell_k][n_models]; padding = (own row, weight 0)
    const unsigned char* point_flags;          // [n_points] 1 = take the generic (CSR / table-walk) path
    int fast_ok;                               // mechanism fits the register-resident reaction form
    int wide_ok;                               // mechanism fits the shared-memory term-list form of the fast kernels
    const int* row_ptr; const int* row_src; const double* row_w; const double* a;
    const int* pulse_ptr; const int* pulse_fn; const double* pulse_w;
    const int* tables; int tables_bytes;
    const double* fields; int n_fields;
    int n_rxn;
    int chunks_per_member;
};

struct OccmMemDev {
    int M;
    const double* k_scale; const double* q_scale; const double* bc_scale;
    const int* active;      // optional member mask of the plain RHS launches (BDF inner loops)
    const double* z;        // optional: evaluate the RHS at y + eps[m] * z   (matrix-free Jacobian action)
    const double* eps;
};

// Views into the shared-memory copy of the table blob.
struct OccmTab {
    int S, n_rxn, n_bc;
    const int* term_ptr; const double* term_coeff; const int* term_rxn; const int* term_prog;
    const int* fac_ptr; const int* fac_sp; const int* fac_ord;
    const int* prog_ptr; const int* code; const double* consts; const int* bc_prog;
};

// Views into a shared-memory copy of the blob.
__device__ __forceinline__ OccmTab occm_tab_views(const int* smem) {
    OccmTab T;
    const char* base = reinterpret_cast<const char*>(smem);
    T.S = smem[OCCM_H_S]; T.n_rxn = smem[OCCM_H_NRXN]; T.n_bc = smem[OCCM_H_NBC];
    T.term_ptr   = reinterpret_cast<const int*>(base + smem[OCCM_H_OFF_TERM_PTR]);
    T.term_coeff = reinterpret_cast<const double*>(base + smem[OCCM_H_OFF_TERM_COEFF]);
    T.term_rxn   = reinterpret_cast<const int*>(base + smem[OCCM_H_OFF_TERM_RXN]);
    T.term_prog  = reinterpret_cast<const int*>(base + smem[OCCM_H_OFF_TERM_PROG]);
    T.fac_ptr    = reinterpret_cast<const int*>(base + smem[OCCM_H_OFF_FAC_PTR]);
    T.fac_sp     = reinterpret_cast<const int*>(base + smem[OCCM_H_OFF_FAC_SP]);
    T.fac_ord    = reinterpret_cast<const int*>(base + smem[OCCM_H_OFF_FAC_ORD]);
    T.prog_ptr   = reinterpret_cast<const int*>(base + smem[OCCM_H_OFF_PROG_PTR]);
    T.code       = reinterpret_cast<const int*>(base + smem[OCCM_H_OFF_CODE]);
    T.consts     = reinterpret_cast<const double*>(base + smem[OCCM_H_OFF_CONST]);
    T.bc_prog    = reinterpret_cast<const int*>(base + smem[OCCM_H_OFF_BC_PROG]);
    return T;
}

// Stage the blob into shared memory (all threads). `smem` must be 8-byte aligned. Ends with __syncthreads().
__device__ __forceinline__ void occm_load_tables(const int* __restrict__ blob, int bytes, int* smem) {
    const int words = bytes >> 2;
    for (int i = threadIdx.x; i < words; i += blockDim.x) smem[i] = __ldg(blob + i);
    __syncthreads();
}

__device__ __forceinline__ double occm_powi(double x, int n) {
    bool neg = n < 0;
    unsigned un = neg ? (unsigned)(-n) : (unsigned)n;
    double r = 1.0, b = x;
    while (un) { if (un & 1u) r *= b; b *= b; un >>= 1; }
    return neg ? 1.0 / r : r;
}

// Postfix evaluator. `row(sp)` returns the concentration of species sp at the current point.
template <class Row>
__device__ __noinline__ double occm_vm_eval(const OccmTab& T, int prog, const Row& row, double t, long long point,
                                            const double* __restrict__ fields, long long n_points) {
    double st[OCCM_VM_STACK];
    int sp = 0;
    const int beg = T.prog_ptr[prog], end = T.prog_ptr[prog + 1];
    for (int pc = beg; pc < end; ++pc) {
        const int ins = T.code[pc];
        const int op = ins & 0xff, arg = ins >> 8;
        switch (op) {
            case OCCM_OP_CONST:   st[sp++] = T.consts[arg]; break;
            case OCCM_OP_SPECIES: st[sp++] = row(arg); break;
            case OCCM_OP_TIME:    st[sp++] = t; break;
            case OCCM_OP_FIELD:   st[sp++] = fields ? fields[(long long)arg * n_points + point] : 0.0; break;
            case OCCM_OP_ADD: --sp; st[sp - 1] = st[sp - 1] + st[sp]; break;
            case OCCM_OP_SUB: --sp; st[sp - 1] = st[sp - 1] - st[sp]; break;
            case OCCM_OP_MUL: --sp; st[sp - 1] = st[sp - 1] * st[sp]; break;
            case OCCM_OP_DIV: --sp; st[sp - 1] = st[sp - 1] / st[sp]; break;
            case OCCM_OP_POW: --sp; st[sp - 1] = pow(st[sp - 1], st[sp]); break;
            case OCCM_OP_NEG: st[sp - 1] = -st[sp - 1]; break;
            case OCCM_OP_POWI: st[sp - 1] = occm_powi(st[sp - 1], arg - OCCM_POWI_BIAS); break;
            case OCCM_OP_SIN:  st[sp - 1] = sin(st[sp - 1]); break;
            case OCCM_OP_COS:  st[sp - 1] = cos(st[sp - 1]); break;
            case OCCM_OP_TAN:  st[sp - 1] = tan(st[sp - 1]); break;
            case OCCM_OP_EXP:  st[sp - 1] = exp(st[sp - 1]); break;
            case OCCM_OP_LOG:  st[sp - 1] = log(st[sp - 1]); break;
            case OCCM_OP_SQRT: st[sp - 1] = sqrt(st[sp - 1]); break;
            case OCCM_OP_ABS:  st[sp - 1] = fabs(st[sp - 1]); break;
            case OCCM_OP_TANH: st[sp - 1] = tanh(st[sp - 1]); break;
            case OCCM_OP_SINH: st[sp - 1] = sinh(st[sp - 1]); break;
            case OCCM_OP_COSH: st[sp - 1] = cosh(st[sp - 1]); break;
            case OCCM_OP_ASIN: st[sp - 1] = asin(st[sp - 1]); break;
            case OCCM_OP_ACOS: st[sp - 1] = acos(st[sp - 1]); break;
            case OCCM_OP_ATAN: st[sp - 1] = atan(st[sp - 1]); break;
            case OCCM_OP_SIGN: st[sp - 1] = (st[sp - 1] > 0.0) - (st[sp - 1] < 0.0); break;
            case OCCM_OP_FLOOR: st[sp - 1] = floor(st[sp - 1]); break;
            case OCCM_OP_CEIL:  st[sp - 1] = ceil(st[sp - 1]); break;
            case OCCM_OP_MIN: --sp; st[sp - 1] = fmin(st[sp - 1], st[sp]); break;
            case OCCM_OP_MAX: --sp; st[sp - 1] = fmax(st[sp - 1], st[sp]); break;
            case OCCM_OP_LT: --sp; st[sp - 1] = st[sp - 1] <  st[sp] ? 1.0 : 0.0; break;
            case OCCM_OP_LE: --sp; st[sp - 1] = st[sp - 1] <= st[sp] ? 1.0 : 0.0; break;
            case OCCM_OP_GT: --sp; st[sp - 1] = st[sp - 1] >  st[sp] ? 1.0 : 0.0; break;
            case OCCM_OP_GE: --sp; st[sp - 1] = st[sp - 1] >= st[sp] ? 1.0 : 0.0; break;
            case OCCM_OP_EQ: --sp; st[sp - 1] = st[sp - 1] == st[sp] ? 1.0 : 0.0; break;
            case OCCM_OP_NE: --sp; st[sp - 1] = st[sp - 1] != st[sp] ? 1.0 : 0.0; break;
            case OCCM_OP_AND: --sp; st[sp - 1] = (st[sp - 1] != 0.0 && st[sp] != 0.0) ? 1.0 : 0.0; break;
            case OCCM_OP_OR:  --sp; st[sp - 1] = (st[sp - 1] != 0.0 || st[sp] != 0.0) ? 1.0 : 0.0; break;
            case OCCM_OP_NOT: st[sp - 1] = st[sp - 1] == 0.0 ? 1.0 : 0.0; break;
            default: break;
        }
    }
    return sp > 0 ? st[sp - 1] : 0.0;
}

struct OccmNoRow { __device__ double operator()(int) const { return 0.0; } };

// g_{b,s}(t): boundary / pulse time function of group b for species s (0 when none was specified).
__device__ __forceinline__ double occm_bc_value(const OccmTab& T, int b, int s, double t) {
    const int prog = T.bc_prog[b * T.S + s];
    if (prog < 0) return 0.0;
    // constant programs (the overwhelmingly common `a: inlet -> 1`) short-cut the interpreter
    const int beg = T.prog_ptr[prog];
    if (T.prog_ptr[prog + 1] - beg == 1 && (T.code[beg] & 0xff) == OCCM_OP_CONST) return T.consts[T.code[beg] >> 8];
    return occm_vm_eval(T, prog, OccmNoRow(), t, 0, nullptr, 0);
}

// Net reaction rate of species s at a point whose concentrations are row(0..S-1)
// (the generated `_ddt[s, :] += ...` lines of system_solvers/reactions.py:450-455).  Generic table walk: used on the
// cold path and for mechanisms that do not fit the register-resident form below (more than OCCM_NT terms per species,
// total order above OCCM_NF, or a byte-code rate factor such as Monod).
template <class Row>
__device__ __forceinline__ double occm_reaction_tables(const OccmTab& T, int s, const Row& row, const double* ksc, double t,
                                                       int point, const OccmSysDev& sys) {
    double acc = 0.0;
    const int tb = T.term_ptr[s], te = T.term_ptr[s + 1];
    for (int k = tb; k < te; ++k) {
        double v = T.term_coeff[k];            // explicit roundings (no FMA contraction): same bits as the register / shared-list forms
        if (ksc) v = __dmul_rn(v, ksc[T.term_rxn[k]]);
        const int fb = T.fac_ptr[k], fe = T.fac_ptr[k + 1];
        for (int f = fb; f < fe; ++f) {
            const double x = row(T.fac_sp[f]);
            const int ord = T.fac_ord[f];
            for (int q = 0; q < ord; ++q) v = __dmul_rn(v, x);
        }
        const int prog = T.term_prog[k];
        if (prog >= 0) v = __dmul_rn(v, occm_vm_eval(T, prog, row, t, (long long)point, sys.fields, sys.n_points));
        acc = __dadd_rn(acc, v);
    }
    return acc;
}

// Register-resident reaction program of ONE species.  A thread keeps its species for the whole kernel (sub-chunks are a
// multiple of S elements), so its stoichiometric terms are decoded from the shared-memory tables once.  Unused terms
// have coefficient 0 and unused factor slots point at column S of the staged row, which always holds 1.0, so the
// evaluation is branch free.
#define OCCM_NT 4   // terms per species
#define OCCM_NF 3   // factor slots per term (a factor of order n occupies n slots)
#define OCCM_WIDE_MAX_TERMS 32   // longest per-species term list of the shared-memory (WIDE) form of the fast kernels
struct OccmRxnRegs {
    double coeff[OCCM_NT];
    unsigned meta[OCCM_NT];   // rxn | f0 << 8 | f1 << 16 | f2 << 24
};

// term i of species s in register form: coefficient and packed meta word (padding term when i is past the species' list)
__device__ __forceinline__ void occm_rxn_term(const OccmTab& T, int s, int i, int S, double& coeff, unsigned& meta) {
    const unsigned pad = (unsigned)S;
    const int tb = T.term_ptr[s], n = T.term_ptr[s + 1] - tb;
    coeff = 0.0;
    meta = (pad << 8) | (pad << 16) | (pad << 24);
    if (i < n) {
        const int k = tb + i;
        coeff = T.term_coeff[k];
        meta = (unsigned)T.term_rxn[k] & 0xffu;
        int slot = 0;
        for (int f = T.fac_ptr[k]; f < T.fac_ptr[k + 1]; ++f)
            for (int q = 0; q < T.fac_ord[f]; ++q) {
                if (slot < OCCM_NF) meta |= ((unsigned)T.fac_sp[f] & 0xffu) << (8 + 8 * slot);
                ++slot;
            }
        for (int q = slot; q < OCCM_NF; ++q) meta |= pad << (8 + 8 * q);
    }
}

__device__ __forceinline__ OccmRxnRegs occm_rxn_regs(const OccmTab& T, int s, int S) {
    OccmRxnRegs R;
#pragma unroll
    for (int i = 0; i < OCCM_NT; ++i) occm_rxn_term(T, s, i, S, R.coeff[i], R.meta[i]);
    return R;
}

// `row` points at the staged concentrations of the point (row[S] == 1.0); coeffm = coeff * k_scale[rxn]
__device__ __forceinline__ double occm_reaction_fast(const double (&coeffm)[OCCM_NT], const unsigned (&meta)[OCCM_NT],
                                                     const double* row, int nt_u, int nf_u) {
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < OCCM_NT; ++i) {
        if (i < nt_u) {                              // uniform bound: no divergence
            const unsigned m = meta[i];
            double v = coeffm[i];
            if (nf_u > 0) v = __dmul_rn(v, row[(m >> 8) & 0xffu]);
            if (nf_u > 1) v = __dmul_rn(v, row[(m >> 16) & 0xffu]);
            if (nf_u > 2) v = __dmul_rn(v, row[m >> 24]);
            acc = __dadd_rn(acc, v);
        }
    }
    return acc;
}

// operator arrays: read-only global memory (ld.global.nc), or — SHARED_SYS, the persistent small-network kernel — a copy in
// shared memory reached through the same struct
template <bool SHARED_SYS, class T_>
__device__ __forceinline__ T_ occm_sys_ld(const T_* p) { if (SHARED_SYS) return *p; return __ldg(p); }

template <bool SHARED_SYS = false>
__device__ __forceinline__ double occm_pulses(const OccmSysDev& sys, const OccmTab& T, int model, int s, double t) {
    double acc = 0.0;
    if (sys.pulse_ptr) {
        for (int k = occm_sys_ld<SHARED_SYS>(sys.pulse_ptr + model); k < occm_sys_ld<SHARED_SYS>(sys.pulse_ptr + model + 1); ++k)
            acc += occm_sys_ld<SHARED_SYS>(sys.pulse_w + k) * occm_bc_value(T, occm_sys_ld<SHARED_SYS>(sys.pulse_fn + k), s, t);
    }
    return acc;
}

// ---------------------------------------------------------------------------------------------------------
// Cold path: d/dt of element (point p, species s) by walking the CSR rows and the mechanism tables.  Taken for points
// whose `point_flags` bit is set (rows with a domain-boundary entry, a pulse source or more entries than the ELL
// width), for PFR inlet nodes, and for every point when the mechanism does not fit the register fast path.
//   in(e)  : stage state at flat element e of this member (global memory)
//   myrow  : staged concentrations of point p (shared memory), prevrow: those of point p - 1 (PFR upwind neighbour)
// CSTR: cstr_system.py:194-203.   PFR: pfr_system.py:290-316.
template <int MODEL, class In, bool SHARED_SYS = false>
__device__ __noinline__ double occm_point_generic(const OccmSysDev& sys, const int* tab_smem, const In& in,
                                                  const double* myrow, const double* prevrow, int p, int s, double t,
                                                  double qs, double bs, const double* ksc) {
    const OccmTab T = occm_tab_views(tab_smem);
    const int S = sys.S;
    auto row = [myrow](int sp) { return myrow[sp]; };
    if (MODEL == 0) {
        double acc = 0.0;
        const int kb = occm_sys_ld<SHARED_SYS>(sys.row_ptr + p), ke = occm_sys_ld<SHARED_SYS>(sys.row_ptr + p + 1);
        for (int k = kb; k < ke; ++k) {
            const int src = occm_sys_ld<SHARED_SYS>(sys.row_src + k);
            const double w = occm_sys_ld<SHARED_SYS>(sys.row_w + k);
            if (src >= 0) acc += w * (src == p ? myrow[s] : in(src * S + s));
            else          acc += w * (bs * occm_bc_value(T, -1 - src, s, t));
        }
        return __dadd_rn(__dadd_rn(__dmul_rn(qs, acc), occm_pulses<SHARED_SYS>(sys, T, p, s, t)), occm_reaction_tables(T, s, row, ksc, t, p, sys));
    } else {
        const int P = sys.P;
        const int k = p / P;
        const int node = p - k * P;
        if (node > 0) {
            const double conv = __dmul_rn(-(qs * occm_sys_ld<SHARED_SYS>(sys.a + k)), myrow[s] - prevrow[s]);
            return __dadd_rn(__dadd_rn(conv, occm_pulses<SHARED_SYS>(sys, T, k, s, t)), occm_reaction_tables(T, s, row, ksc, t, p, sys));
        }
        double acc = 0.0;
        const int eb = occm_sys_ld<SHARED_SYS>(sys.row_ptr + k), ee = occm_sys_ld<SHARED_SYS>(sys.row_ptr + k + 1);
        for (int e = eb; e < ee; ++e) {
            const int src = occm_sys_ld<SHARED_SYS>(sys.row_src + e);
            const double w = occm_sys_ld<SHARED_SYS>(sys.row_w + e);
            if (src >= 0) {
                const int o = P * (src + 1) - 1;                 // outlet node of the upstream PFR
                auto orow = [&in, o, S](int sp) { return in(o * S + sp); };
                double d = -(qs * occm_sys_ld<SHARED_SYS>(sys.a + src)) * (in(o * S + s) - in((o - 1) * S + s));
                d += occm_pulses<SHARED_SYS>(sys, T, src, s, t);
                d += occm_reaction_tables(T, s, orow, ksc, t, o, sys);
                acc += w * d;
            } else {
                acc += w * (bs * occm_bc_value(T, -1 - src, s, t));
            }
        }
        return acc;
    }
}

// deterministic block-wide sum (fixed tree), result valid in thread 0
__device__ __forceinline__ double occm_block_sum(double v, double* scratch) {
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        for (int w = 0; w < nw; ++w) r += scratch[w];
    }
    __syncthreads();
    return r;
}
