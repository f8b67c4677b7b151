// Mechanism-specialised stage kernels: what liboccm_b200.so (loader + launcher, occm_solver.cu) and a specialised kernel image
// (spec/occm_spec.cu, one cubin per mechanism / species count / ELL width) have to agree on.
//
// The reference writes the mechanism as the source of one `@njit` function per reactions file and lets numba compile it
// (openccm/system_solvers/reactions.py:41-93, 395-467), i.e. the mechanism is a compile-time constant of its RHS.  The same holds here: the
// host side (openccm_b200/specialize.py) writes the stoichiometric terms of the parsed mechanism as straight-line fp64 code,
// nvcc compiles spec/occm_spec.cu around it for sm_100a, and occm_system_load_specialized() hands the image to the library.
// Without an image the table-driven kernels (occm_pipe.cuh / occm_fast.cuh) run: same arithmetic, same roundings.
#pragma once

#define OCCM_SPEC_ABI 6                 // bump whenever OccmRowGeom, a kernel signature or the meaning of occm_spec_info changes

// Consumer warps per CTA (32 points each; + one producer warp) and CTAs per SM the register budget is sized for, per stage.
// The kernels are latency bound on the out-of-chunk row gathers, so the plain RHS (one staged stream: small ring slots) runs as
// many warps as 128 registers allow; stages 2 and 7 stage two / three streams and are limited by shared memory.
#ifndef OCCM_ROW_W0
#define OCCM_ROW_W0 7
#endif
#ifndef OCCM_ROW_C0
#define OCCM_ROW_C0 2
#endif
#ifndef OCCM_ROW_W1
#define OCCM_ROW_W1 4                   // stage 1 (state y + eps z of the matrix-free Jacobian action): two staged streams like stage 2
#endif
#ifndef OCCM_ROW_C1
#define OCCM_ROW_C1 2
#endif
#ifndef OCCM_ROW_W2
#define OCCM_ROW_W2 4
#endif
#ifndef OCCM_ROW_C2
#define OCCM_ROW_C2 2
#endif
#ifndef OCCM_ROW_W7
#define OCCM_ROW_W7 3
#endif
#ifndef OCCM_ROW_C7
#define OCCM_ROW_C7 3
#endif
#define OCCM_ROW_MAX_DEPTH 8
#define OCCM_SPEC_MAX_TERMS 256
#define OCCM_SPEC_MAX_S 16

// words of the image's `occm_spec_info` array
enum { OCCM_SI_ABI = 0, OCCM_SI_S, OCCM_SI_K, OCCM_SI_NTERM, OCCM_SI_HASH, OCCM_SI_MODEL, OCCM_SI_W0, OCCM_SI_C0, OCCM_SI_W1, OCCM_SI_C1,
       OCCM_SI_W2, OCCM_SI_C2, OCCM_SI_W7, OCCM_SI_C7, OCCM_SI_WORDS };

struct OccmRowGeom {
    int cpm;           // chunks (32 * warps points) per member
    int depth;         // ring slots
    int slot_bytes;    // NS streams of stream_bytes, then K ELL columns of points * 16 bytes
    int stream_bytes;  // points * S * 8
    int err_stride;    // error partials per member (stage 7)
    int coef_off;      // byte offsets into dynamic shared memory: per-warp scaled coefficients,
    int bar_off;       //   mbarriers (full[depth], empty[depth]),
    int ring_off;      //   ring slots,
    int out_off;       //   per-warp staging of the output rows (bulk stores)
};

// FNV-1a over the packed mechanism table blob: ties an image to the mechanism it was generated from
static inline unsigned long long occm_spec_hash(const unsigned char* p, size_t n) {
    unsigned long long h = 14695981039346656037ull;
    for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}
