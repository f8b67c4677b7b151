// Layout of the packed mechanism blob ("tables"), shared by the host packer
// (openccm_b200/system_solvers/tables.py) and the device code.
//
// The reference turns the reactions file and the boundary-condition strings into *generated Python source*
// that numba compiles (system_solvers/reactions.py:395-467, boundary_and_initial_conditions.py:211-245).
// Here the same information is data: per-species lists of stoichiometric terms
//     d c_s/dt += coeff * k_scale[member][rxn] * prod_f c[sp_f]^ord_f * (prog >= 0 ? eval(prog) : 1)
// plus tiny postfix programs for non-polynomial rate factors (e.g. Monod 1/(0.25 + s)) and for the time
// functions g(t) of the boundary conditions (smoothed Heaviside ramps, their derivatives, pulses).
// The blob lives in global memory and is staged into shared memory once per CTA.
#pragma once
#include <stdint.h>

#define OCCM_TABLES_MAGIC 0x4F43434D /* 'OCCM' */

// int32 header slots
enum {
    OCCM_H_MAGIC = 0, OCCM_H_BYTES, OCCM_H_S, OCCM_H_NRXN, OCCM_H_NTERM, OCCM_H_NFAC, OCCM_H_NPROG, OCCM_H_NCODE,
    OCCM_H_NCONST, OCCM_H_NBC, OCCM_H_NFIELD,
    OCCM_H_OFF_TERM_PTR,    // int32 [S + 1]
    OCCM_H_OFF_TERM_COEFF,  // double[n_term]
    OCCM_H_OFF_TERM_RXN,    // int32 [n_term]   reaction index (k_scale column)
    OCCM_H_OFF_TERM_PROG,   // int32 [n_term]   -1 or program id of an extra multiplicative factor
    OCCM_H_OFF_FAC_PTR,     // int32 [n_term + 1]
    OCCM_H_OFF_FAC_SP,      // int32 [n_fac]    species index
    OCCM_H_OFF_FAC_ORD,     // int32 [n_fac]    positive integer order
    OCCM_H_OFF_PROG_PTR,    // int32 [n_prog + 1]
    OCCM_H_OFF_CODE,        // int32 [n_code]   op | (arg << 8)
    OCCM_H_OFF_CONST,       // double[n_const]
    OCCM_H_OFF_BC_PROG,     // int32 [n_bc * S] program id of g_{b,s}(t) or -1 (no condition given: contributes 0)
    OCCM_H_MAX_TERMS,       // max number of terms of any species (uniform unroll bound of the register fast path)
    OCCM_H_MAX_SLOTS,       // max total order (factor slots) of any term
    OCCM_H_ANY_PROG,        // 1 if any term carries a byte-code factor
    OCCM_H_WORDS = 32
};

// postfix byte code
enum {
    OCCM_OP_CONST = 0,  // push consts[arg]
    OCCM_OP_SPECIES,    // push c[arg] at this point
    OCCM_OP_TIME,       // push t
    OCCM_OP_FIELD,      // push fields[arg][point]
    OCCM_OP_ADD, OCCM_OP_SUB, OCCM_OP_MUL, OCCM_OP_DIV, OCCM_OP_POW, OCCM_OP_NEG,
    OCCM_OP_POWI,       // x ** arg (arg signed, stored biased by +2^22)
    OCCM_OP_SIN, OCCM_OP_COS, OCCM_OP_TAN, OCCM_OP_EXP, OCCM_OP_LOG, OCCM_OP_SQRT, OCCM_OP_ABS, OCCM_OP_TANH,
    OCCM_OP_SINH, OCCM_OP_COSH, OCCM_OP_ASIN, OCCM_OP_ACOS, OCCM_OP_ATAN, OCCM_OP_SIGN, OCCM_OP_FLOOR, OCCM_OP_CEIL,
    OCCM_OP_MIN, OCCM_OP_MAX,
    OCCM_OP_LT, OCCM_OP_LE, OCCM_OP_GT, OCCM_OP_GE, OCCM_OP_EQ, OCCM_OP_NE,   // push 1.0 / 0.0
    OCCM_OP_AND, OCCM_OP_OR, OCCM_OP_NOT,
    OCCM_OP_COUNT
};
#define OCCM_POWI_BIAS (1 << 22)
#define OCCM_VM_STACK 16
