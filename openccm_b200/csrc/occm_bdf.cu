// Batched variable-order BDF/NDF integrator with a Jacobian-free Newton-Krylov inner solve (sm_100a).
//
// Time stepping, error control and order selection follow scipy.integrate.BDF (scipy/integrate/_ivp/bdf.py:
// compute_R / change_D :16-33, solve_bdf_system :36-70, _step_impl :310-455, BdfDenseOutput :458-483) member by member.
// What differs from scipy is the linear algebra of the Newton iteration: scipy factorises I - c*J with a dense
// finite-difference Jacobian (N_dof + 1 RHS calls and an O(N_dof^3) LU — the cost that dominates the reference on stiff
// networks); here (I - c*J) dy = b is solved by right-preconditioned restarted GMRES in the error-weight
// scaling of the Newton test, with
//   * J*v by a directional difference of the fused RHS kernel (always the Jacobian at the current predictor, so the
//     "stale Jacobian" bookkeeping of scipy disappears), and
//   * a block preconditioner: per point the S x S block I - c*(transport diagonal + d r/d c); its inverse is formed
//     once (warp-cooperative Gauss-Jordan), kept in HBM and reused — CVODE's setup policy: rebuilt when c = h/alpha
//     has moved by more than 20 %, after 20 attempts, or after a slow / failed Newton solve — so an application is one
//     S x S mat-vec per point (precond_mode 1 instead rebuilds and solves the block on every application, nothing stored).
// All members advance in lock step through the phases of a step attempt; per-member masks switch members off inside
// the Newton and GMRES loops.  Reductions go through per-chunk partial sums added in a fixed order (deterministic).
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "occm_host.h"

#define BDF_MAX_ORDER 5
// Storage type of the preconditioner's block inverses.  They are 8 S of the (32 + 8 S) bytes per unknown a preconditioner
// application moves in fp64, i.e. the kernel is a stream over them; a preconditioner needs no more than single precision (GMRES is
// flexible here — the preconditioned vectors are stored — and Newton is driven by fp64 residuals), so they are stored as float:
// (32 + 4 S) bytes per unknown.  -DOCCM_BDF_BINV_F64 keeps doubles (A/B).
#ifdef OCCM_BDF_BINV_F64
typedef double occm_binv_t;
#else
typedef float occm_binv_t;
#endif
// Layout: Binv[member][unknown i of a point][j] = (B_point^-1)(i, j), rows padded to whole 16-byte vectors: a lane reads its row
// with 128-bit loads.  Measured on C5 (S = 12, 2.4e7 unknowns, one application): float rows 0.53 ms; float column-major with
// 4-byte loads 2.27 ms; float interleaved so that the lanes of a point read contiguous 16-byte pieces 0.63 ms; fp64
// column-major (round 1) 0.67 ms, fp64 rows 0.78 ms.
#define BDF_BINV_VEC (16 / (int)sizeof(occm_binv_t))
#ifndef BDF_SWEEP_AHEAD
#define BDF_SWEEP_AHEAD 2        // points of a PFR sweep whose operands are in flight (preconditioner application)
#endif
static inline __host__ __device__ int bdf_binv_row(int S) { return (S + BDF_BINV_VEC - 1) / BDF_BINV_VEC * BDF_BINV_VEC; }
#define BDF_NEWTON_MAXITER 4
#define BDF_NVEC_D 8            // MAX_ORDER + 3 difference vectors
#define BDF_CHUNK 2048          // elements per CTA trip of the streaming kernels (256 threads x 8)
#define BDF_MAXDOT 48           // partial-sum slots per chunk (>= restart + 2)
// Second Gram-Schmidt pass only when the first one cancelled enough digits to matter: one classical pass leaves the new vector
// orthogonal to the basis to about eps |w| / |w'|, so |w'| < eta |w| with eta = 1e-4 (eta^2 here) keeps the basis orthogonal to
// 1e-12 — far below the linear tolerance.  The DGKS value eta = 1/sqrt(2) (and Rutishauser's 1/10) fired on every Krylov vector
// of C5: the preconditioned operator is close to the identity, so A v_j always has its largest component along v_j; that was a
// full extra pass over the basis per iteration (0.15-0.25 ms of 1.7 ms on C5).
#ifndef BDF_REORTH_ETA2
#define BDF_REORTH_ETA2 1e-8
#endif
#define BDF_MIN_FACTOR 0.2
#define BDF_MAX_FACTOR 10.0

struct OccmBdfDev {
    int M, S, P, model, nchunk, restart;
    long long ndof, n_points, vstride;      // vstride = M * ndof
    double* D;                               // [8][M][ndof]
    double* y; double* psi; double* d; double* f; double* f2; double* z; double* w;
    double* sc;     // error weights of the attempt, atol + rtol |y_predict| (written by bdf_predict; y - d would rebuild y_predict)
    double* V;                               // [restart + 1][M][ndof]
    double* Z;                               // [restart][M][ndof]  preconditioned Krylov vectors z_j = M^-1 S v_j (flexible GMRES)
    double* t; double* h_abs; double* t_new; double* c; double* t_old_out;
    int* order; int* n_equal; int* status; int* fresh;      // fresh: first attempt of a new step
    int* newton_active; int* newton_k; int* converged; int* n_iter; double* dy_norm_old;
    int* gmres_active; int* gmres_j; double* H; double* cs; double* sn; double* gv; double* ycoef; double* eps; double* ynorm;
    double* RU; int* need_change; int* attempt_ok; double* safety; double* err_norm;
    double* partial; double* norms;
    int* accepted_now; int* out_lo; int* out_hi; int* n_rows;
    long long* nfev; long long* n_acc; long long* n_rej; long long* n_newton_fail; long long* n_gmres; long long* n_precond;
    int* counters;                           // [0] running, [1] newton active, [2] gmres active
    double rtol, atol, t_bound, newton_tol, lin_tol;
    long long max_steps;
    const double* tdiag;                     // transport diagonal per model (CSTR: A_jj, PFR: -Q/dV)
    const double* q_scale; const double* k_scale;
    // stored block inverses (precond_mode 0): Binv[m][unknown i][j] = (B_point^-1)(i, j) (occm_binv_t rows, see above)
    int store;
    occm_binv_t* Binv; double* c_fact; int* need_factor; int* since_factor; int* force_factor; long long* n_factor;
    int* reorth; double* wn2;                // second Gram-Schmidt pass wanted (DGKS test); |w|^2 before orthogonalisation
    double* crate; int carry;                // convergence rate of the Newton iteration carried from step to step (CVODE); 0 = scipy's rule
};

struct occm_bdf {
    const occm_system* sys;
    occm_members mem; bool has_mem;
    occm_bdf_problem prob;
    OccmBdfDev d;
    int M;
    int* h_counters;     // pinned host [4]
};

// ------------------------------------------------------------------------------------------------ small helpers
__device__ __forceinline__ double bdf_next_up(double t) { return __longlong_as_double(__double_as_longlong(t) + 1ll); }

__constant__ double c_kappa[6] = {0.0, -0.1850, -1.0 / 9, -0.0823, -0.0415, 0.0};

__device__ __forceinline__ double bdf_gamma(int k) {      // gamma_k = sum_{j<=k} 1/j   (bdf.py:253)
    double g = 0.0;
    for (int j = 1; j <= k; ++j) g += 1.0 / j;
    return g;
}
__device__ __forceinline__ double bdf_alpha(int k) { return (1.0 - c_kappa[k]) * bdf_gamma(k); }
__device__ __forceinline__ double bdf_error_const(int k) { return c_kappa[k] * bdf_gamma(k) + 1.0 / (k + 1); }

// R(order, factor) and RU = R(order, factor) @ R(order, 1)   (bdf.py:16-33); RU is (order+1)x(order+1), row-major 6x6
__device__ void bdf_compute_R(int order, double factor, double (*Rm)[6]) {
    for (int i = 0; i <= order; ++i)
        for (int j = 0; j <= order; ++j) Rm[i][j] = 0.0;
    for (int j = 0; j <= order; ++j) Rm[0][j] = 1.0;
    for (int i = 1; i <= order; ++i)
        for (int j = 1; j <= order; ++j) Rm[i][j] = Rm[i - 1][j] * ((i - 1 - factor * j) / i);   // cumprod down the rows
}
__device__ void bdf_set_change(const OccmBdfDev& b, int m, int order, double factor) {
    double R[6][6], U[6][6];
    bdf_compute_R(order, factor, R);
    bdf_compute_R(order, 1.0, U);
    double* RU = b.RU + (long long)m * 36;
    for (int i = 0; i <= order; ++i)
        for (int j = 0; j <= order; ++j) {
            double acc = 0.0;
            for (int k = 0; k <= order; ++k) acc += R[i][k] * U[k][j];
            RU[i * 6 + j] = acc;
        }
    b.need_change[m] = order + 1;      // number of difference rows to transform
}

// ------------------------------------------------------------------------------------------------ reductions
// partial sums are stored [member][slot][chunk]: one slot is a contiguous run over the chunks
#define BDF_PIDX(b, m, ci, k) ((((long long)(m) * BDF_MAXDOT + (k)) * (b).nchunk) + (ci))

__global__ void __launch_bounds__(512) bdf_reduce(const double* __restrict__ partial, int nchunk, int nvals, double* __restrict__ out,
                                                  const int* __restrict__ mask) {
    __shared__ double scratch[16];
    const int m = blockIdx.x, k = blockIdx.y;            // one CTA per (member, slot): with one member a single CTA summed up to 32 slots in turn
    if (mask && !mask[m]) return;
    const double* p = partial + ((long long)m * BDF_MAXDOT + k) * nchunk;
    double v = 0.0;
    for (int i = threadIdx.x; i < nchunk; i += blockDim.x) v += p[i];
    const double s = occm_block_sum(v, scratch);
    if (threadIdx.x == 0) out[(long long)m * BDF_MAXDOT + k] = s;
}

// chunk iteration helper: CTA c handles elements [ci * BDF_CHUNK, ...) of member m
#define BDF_FOR_CHUNKS(b, mask_expr)                                                              \
    const long long total_chunks = (long long)(b).M * (b).nchunk;                                 \
    for (long long cc = blockIdx.x; cc < total_chunks; cc += gridDim.x)                           \
        if (int m = (int)(cc / (b).nchunk); true)                                                 \
            if (mask_expr)                                                                        \
                if (int ci = (int)(cc - (long long)m * (b).nchunk); true)

#define BDF_ELEMS(b, ci, e) for (long long e = (long long)(ci) * BDF_CHUNK + threadIdx.x, e_end = min((long long)((ci) + 1) * BDF_CHUNK, (b).ndof); e < e_end; e += blockDim.x)

// ------------------------------------------------------------------------------------------------ vector kernels
__global__ void __launch_bounds__(256) bdf_change_D(const OccmBdfDev b) {
    BDF_FOR_CHUNKS(b, b.need_change[m] > 0) {
        const int n = b.need_change[m];
        const double* RU = b.RU + (long long)m * 36;
        BDF_ELEMS(b, ci, e) {
            double v[6], r[6];
            for (int j = 0; j < n; ++j) v[j] = b.D[j * b.vstride + (long long)m * b.ndof + e];
            for (int i = 0; i < n; ++i) {            // D[:o+1] = RU.T @ D[:o+1]
                double acc = 0.0;
                for (int j = 0; j < n; ++j) acc += RU[j * 6 + i] * v[j];
                r[i] = acc;
            }
            for (int i = 0; i < n; ++i) b.D[i * b.vstride + (long long)m * b.ndof + e] = r[i];
        }
    }
}

// The order is uniform per member: the element loops are instantiated per order so that the difference rows live in
// registers (a runtime-indexed array went to local memory) and gamma_i is folded at compile time.
template <int O>
__device__ __forceinline__ void bdf_predict_elems(const OccmBdfDev& b, int m, int ci) {
    double gam[O + 1];
#pragma unroll
    for (int i = 0; i <= O; ++i) gam[i] = bdf_gamma(i);
    const double inv_alpha = 1.0 / bdf_alpha(O);
    BDF_ELEMS(b, ci, e) {
        const long long g = (long long)m * b.ndof + e;
        double yp = b.D[g], ps = 0.0;
#pragma unroll
        for (int i = 1; i <= O; ++i) {
            const double di = b.D[i * b.vstride + g];
            yp += di;
            ps += di * gam[i];
        }
        b.y[g] = yp; b.psi[g] = ps * inv_alpha; b.d[g] = 0.0; b.sc[g] = b.atol + b.rtol * fabs(yp);
    }
}

__global__ void __launch_bounds__(256) bdf_predict(const OccmBdfDev b) {      // bdf.py:358-362
    BDF_FOR_CHUNKS(b, b.status[m] == OCCM_MEMBER_RUNNING) {
        switch (b.order[m]) {
            case 1: bdf_predict_elems<1>(b, m, ci); break;
            case 2: bdf_predict_elems<2>(b, m, ci); break;
            case 3: bdf_predict_elems<3>(b, m, ci); break;
            case 4: bdf_predict_elems<4>(b, m, ci); break;
            default: bdf_predict_elems<5>(b, m, ci); break;
        }
    }
}

// b~ = (c f - psi - d) / scale -> V[0];  partial[0] = |b~|^2, partial[1] = |y|^2     (bdf.py:46, scaled by the error weights)
__global__ void __launch_bounds__(256) bdf_newton_rhs(const OccmBdfDev b) {
    __shared__ double scratch[8];
    BDF_FOR_CHUNKS(b, b.newton_active[m]) {
        const double c = b.c[m];
        double s0 = 0.0, s1 = 0.0;
        BDF_ELEMS(b, ci, e) {
            const long long g = (long long)m * b.ndof + e;
            const double y = b.y[g], d = b.d[g];
            const double scale = b.atol + b.rtol * fabs(y - d);
            const double r = (c * b.f[g] - b.psi[g] - d) / scale;
            b.V[g] = r;
            s0 += r * r; s1 += y * y;
        }
        const double t0 = occm_block_sum(s0, scratch), t1 = occm_block_sum(s1, scratch);
        if (threadIdx.x == 0) { b.partial[BDF_PIDX(b, m, ci, 0)] = t0; b.partial[BDF_PIDX(b, m, ci, 1)] = t1; }
    }
}

// ---- warp-cooperative block preconditioner: out = M^{-1} (scale .* in), M ~ I - c J, nothing stored between applications
// W = 4, 8 or 16 lanes form a group, lane i < S owns row i of the S x S block in registers; elimination without
// pivoting through warp shuffles (I - c J of a mass-action mechanism is column diagonally dominant).
//   CSTR: one group per point (block Jacobi).
//   PFR : one group per PFR sweeps its points in flow order, x_p = B_p^{-1} (u_p + c a x_{p-1}): the exact inverse of
//         the block lower-bidiagonal operator of the upwind discretisation (pfr_system.py:296) — block ILU(0) == LU here.
//         Inlet nodes (coupled to OTHER PFRs, pfr_system.py:313-314) keep the identity; that coupling is left to GMRES.
template <int W>
__device__ __forceinline__ void bdf_group_rows(const OccmTab& T, int nterm, int nfac, int S, int lane, double cown, double c, double tau,
                                               const double* ksc, double (&row)[W]) {
    // row `lane` of B = (1 - c tau) I - c dr/dc
#pragma unroll
    for (int j = 0; j < W; ++j) row[j] = (j == lane) ? 1.0 - c * tau : 0.0;
    auto conc = [&](int sp) { return __shfl_sync(0xffffffffu, cown, sp, W); };
    const int tb = lane < S ? T.term_ptr[lane] : 0, te = lane < S ? T.term_ptr[lane + 1] : 0;
    // all lanes of the warp must execute the same shuffles: nterm / nfac are the mechanism-wide maxima (table header)
    for (int kk = 0; kk < nterm; ++kk) {
        const int k = tb + kk;
        const bool on = k < te;
        double v0 = on ? T.term_coeff[k] : 0.0;
        if (on && ksc) v0 *= ksc[T.term_rxn[k]];
        const int fb = on ? T.fac_ptr[k] : 0, fe = on ? T.fac_ptr[k + 1] : 0;
        // values of the factors of this term (up to 4 distinct factors are supported in the preconditioner)
        double xv[4]; int xs[4], xo[4];
#pragma unroll
        for (int f = 0; f < 4; ++f) {
            const bool fon = f < nfac && fb + f < fe;
            xs[f] = fon ? T.fac_sp[fb + f] : 0; xo[f] = fon ? T.fac_ord[fb + f] : 0;
            xv[f] = f < nfac ? conc(xs[f]) : 1.0;
        }
#pragma unroll
        for (int f = 0; f < 4; ++f) {
            if (xo[f] == 0) continue;
            double dv = v0 * xo[f];
            for (int q = 1; q < xo[f]; ++q) dv *= xv[f];
#pragma unroll
            for (int g = 0; g < 4; ++g) if (g != f) for (int q = 0; q < xo[g]; ++q) dv *= xv[g];
#pragma unroll
            for (int j = 0; j < W; ++j) row[j] -= (j == xs[f]) ? c * dv : 0.0;
        }
    }
}

// Gauss-Jordan inverse without pivoting: on return inv[] is row `lane` of B^-1 (row[] is destroyed).  At pivot k the
// pivot row of the right half is non-zero only in columns <= k and of the left half only columns > k matter: W shuffles.
template <int W>
__device__ __forceinline__ void bdf_group_invert(int S, int lane, double (&row)[W], double (&inv)[W]) {
#pragma unroll
    for (int j = 0; j < W; ++j) inv[j] = (j == lane) ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < W; ++k) {
        if (k >= S) break;
        const double ipk = 1.0 / __shfl_sync(0xffffffffu, row[k], k, W);
        const bool piv = lane == k;
        const double l = (!piv && lane < S) ? row[k] * ipk : 0.0;
#pragma unroll
        for (int j = 0; j < W; ++j) {
            if (j > k) {
                const double pkj = __shfl_sync(0xffffffffu, row[j], k, W);
                row[j] = piv ? pkj * ipk : row[j] - l * pkj;
            } else {
                const double pij = __shfl_sync(0xffffffffu, inv[j], k, W);
                inv[j] = piv ? pij * ipk : inv[j] - l * pij;
            }
        }
    }
}

template <int W>
__device__ __forceinline__ void bdf_group_solve(const OccmTab& T, int nterm, int nfac, int S, int lane, double cown, double c, double tau,
                                                const double* ksc, double& u) {
    double row[W];
    bdf_group_rows<W>(T, nterm, nfac, S, lane, cown, c, tau, ksc, row);
    // elimination (no pivoting)
#pragma unroll
    for (int k = 0; k < W; ++k) {
        if (k >= S) break;
        const double ipk = 1.0 / __shfl_sync(0xffffffffu, row[k], k, W);
        const double uk = __shfl_sync(0xffffffffu, u, k, W);
        const double l = (lane > k && lane < S) ? row[k] * ipk : 0.0;
#pragma unroll
        for (int j = 0; j < W; ++j) {
            if (j <= k) continue;
            const double pkj = __shfl_sync(0xffffffffu, row[j], k, W);
            row[j] -= l * pkj;
        }
        u -= l * uk;
    }
    // back substitution
#pragma unroll
    for (int k = W - 1; k >= 0; --k) {
        if (k >= S) continue;
        const double xk = __shfl_sync(0xffffffffu, u / row[k], k, W);      // lane k holds x_k = u_k / B_kk
        if (lane == k) u = xk;
        else if (lane < k) u -= row[k] * xk;
    }
}

template <int MODEL, int W>
__global__ void __launch_bounds__(128) bdf_precond_warp(const __grid_constant__ OccmSysDev sys, const OccmBdfDev b, const double* src,
                                                        int src_is_V_col_j, double* __restrict__ dst, const int* __restrict__ mask, int keep_z) {
    extern __shared__ double smem_d[];
    int* tab_s = reinterpret_cast<int*>(smem_d);
    occm_load_tables(sys.tables, sys.tables_bytes, tab_s);
    const OccmTab T = occm_tab_views(tab_s);
    const int S = sys.S, P = sys.P;
    const int nterm = tab_s[OCCM_H_MAX_TERMS], nfac = min(4, tab_s[OCCM_H_MAX_SLOTS]);
    const int lane = threadIdx.x % W, gib = threadIdx.x / W, gpb = 128 / W;
    const long long per_member = MODEL == 0 ? sys.n_points : sys.n_models;
    const long long total = (long long)b.M * per_member;
    const long long rounded = (total + gpb - 1) / gpb * gpb;
    for (long long g = (long long)blockIdx.x * gpb + gib; g < rounded; g += (long long)gridDim.x * gpb) {
        const bool gv = g < total;
        const int m = gv ? (int)(g / per_member) : 0;
        const long long idx = gv ? g - (long long)m * per_member : 0;
        const bool on = gv && (!mask || mask[m]) && lane < S;
        const double c = b.c[m];
        const double qs = b.q_scale ? b.q_scale[m] : 1.0;
        const double* ksc = b.k_scale ? b.k_scale + (long long)m * sys.n_rxn : nullptr;
        const double* in = src + (src_is_V_col_j ? (long long)b.gmres_j[m] * b.vstride : 0) + (long long)m * b.ndof;
        // inside GMRES (keep_z) the source is the not yet normalised Krylov vector: it is scaled by 1 / h_{j+1,j} (b.eps) on the way
        // in and the normalised value is written back — the separate scaling pass over V is gone
        const double va = keep_z ? b.eps[m] : 1.0;
        double* vin = const_cast<double*>(in);
        if (MODEL == 0) {
            const long long g0 = (long long)m * b.ndof + idx * S + lane;
            const double y = on ? b.y[g0] : 0.0;
            const double vw = on ? in[idx * S + lane] * va : 0.0;
            if (on && keep_z) vin[idx * S + lane] = vw;
            double u = on ? b.sc[g0] * vw : 0.0;
            bdf_group_solve<W>(T, nterm, nfac, S, lane, y, c, qs * b.tdiag[gv ? idx : 0], ksc, u);
            if (on) { dst[g0] = u; if (keep_z) b.Z[(long long)b.gmres_j[m] * b.vstride + g0] = u; }
        } else {
            const double ca = c * qs * (-b.tdiag[gv ? idx : 0]);             // c * Q/dV of this PFR
            double xprev = 0.0;
            const long long gbase = (long long)m * b.ndof + idx * P * S + lane;
            const double* inp = in + idx * P * S + lane;
            double* inpm = vin + idx * P * S + lane;
            double y = on ? b.y[gbase] : 0.0, d = on ? b.sc[gbase] : 0.0, w = on ? inp[0] : 0.0;       // d: error weight
            for (int i = 0; i < P; ++i) {
                // operands of the next point are requested before this point's (sequentially dependent) solve
                const bool more = on && i + 1 < P;
                const double yn = more ? b.y[gbase + (long long)(i + 1) * S] : 0.0;
                const double dn = more ? b.sc[gbase + (long long)(i + 1) * S] : 0.0;
                const double wn = more ? inp[(long long)(i + 1) * S] : 0.0;
                const double vw = w * va;
                if (on && keep_z) inpm[(long long)i * S] = vw;
                double u = d * vw;
                if (i > 0) {
                    u += ca * xprev;
                    bdf_group_solve<W>(T, nterm, nfac, S, lane, y, c, -ca / c, ksc, u);
                }
                if (on) { dst[gbase + (long long)i * S] = u; if (keep_z) b.Z[(long long)b.gmres_j[m] * b.vstride + gbase + (long long)i * S] = u; }
                xprev = u;
                y = yn; d = dn; w = wn;
            }
        }
    }
}

// row of a stored block inverse -> registers / registers -> row (128-bit accesses; entries past S are zero)
template <int W>
__device__ __forceinline__ void bdf_binv_raw(const occm_binv_t* __restrict__ row, int RS, bool on, int4 (&q)[W / BDF_BINV_VEC]) {
#pragma unroll
    for (int v = 0; v < W / BDF_BINV_VEC; ++v) {
        q[v] = make_int4(0, 0, 0, 0);                          // zero bits are 0.0 in both storage types
        if (on && v * BDF_BINV_VEC < RS) q[v] = __ldg(reinterpret_cast<const int4*>(row) + v);
    }
}
template <int W>
__device__ __forceinline__ void bdf_binv_unpack(const int4 (&q)[W / BDF_BINV_VEC], double (&r)[W]) {
#pragma unroll
    for (int v = 0; v < W / BDF_BINV_VEC; ++v) {
        const int j = v * BDF_BINV_VEC;
        if (sizeof(occm_binv_t) == 4) {
            r[j] = (double)__int_as_float(q[v].x);
            if (j + 1 < W) r[j + 1] = (double)__int_as_float(q[v].y);
            if (j + 2 < W) r[j + 2] = (double)__int_as_float(q[v].z);
            if (j + 3 < W) r[j + 3] = (double)__int_as_float(q[v].w);
        } else {
            r[j] = __hiloint2double(q[v].y, q[v].x);
            if (j + 1 < W) r[j + 1] = __hiloint2double(q[v].w, q[v].z);
        }
    }
}
template <int W>
__device__ __forceinline__ void bdf_binv_load(const occm_binv_t* __restrict__ row, int RS, bool on, double (&r)[W]) {
    int4 q[W / BDF_BINV_VEC];
    bdf_binv_raw<W>(row, RS, on, q);
    bdf_binv_unpack<W>(q, r);
}
template <int W>
__device__ __forceinline__ void bdf_binv_store(occm_binv_t* __restrict__ row, int S, int RS, const double (&inv)[W]) {
#pragma unroll
    for (int j = 0; j < W; j += BDF_BINV_VEC) {
        if (j >= RS) break;
        int4 raw;
        if (sizeof(occm_binv_t) == 4) {
            raw.x = __float_as_int(j < S ? (float)inv[j] : 0.0f);
            raw.y = __float_as_int(j + 1 < S && j + 1 < W ? (float)inv[j + 1 < W ? j + 1 : 0] : 0.0f);
            raw.z = __float_as_int(j + 2 < S && j + 2 < W ? (float)inv[j + 2 < W ? j + 2 : 0] : 0.0f);
            raw.w = __float_as_int(j + 3 < S && j + 3 < W ? (float)inv[j + 3 < W ? j + 3 : 0] : 0.0f);
        } else {
            const double a = j < S ? inv[j] : 0.0, c = j + 1 < S && j + 1 < W ? inv[j + 1 < W ? j + 1 : 0] : 0.0;
            raw.x = __double2loint(a); raw.y = __double2hiint(a); raw.z = __double2loint(c); raw.w = __double2hiint(c);
        }
        *reinterpret_cast<int4*>(row + j) = raw;
    }
}

// ---- stored variant: factor (all points in parallel, members with need_factor) ...
template <int W>
__global__ void __launch_bounds__(128) bdf_precond_factor(const __grid_constant__ OccmSysDev sys, const OccmBdfDev b) {
    extern __shared__ double smem_d[];
    if (b.counters[3] == 0) return;                   // no member asked for a rebuild in this attempt
    int* tab_s = reinterpret_cast<int*>(smem_d);
    occm_load_tables(sys.tables, sys.tables_bytes, tab_s);
    const OccmTab T = occm_tab_views(tab_s);
    const int S = sys.S, P = sys.P, RS = bdf_binv_row(S);
    const int nterm = tab_s[OCCM_H_MAX_TERMS], nfac = min(4, tab_s[OCCM_H_MAX_SLOTS]);
    const int lane = threadIdx.x % W, gib = threadIdx.x / W, gpb = 128 / W;
    const long long per_member = sys.n_points;
    const long long total = (long long)b.M * per_member;
    const long long rounded = (total + gpb - 1) / gpb * gpb;
    for (long long g = (long long)blockIdx.x * gpb + gib; g < rounded; g += (long long)gridDim.x * gpb) {
        const bool gv = g < total;
        const int m = gv ? (int)(g / per_member) : 0;
        const long long idx = gv ? g - (long long)m * per_member : 0;
        const bool inlet_node = sys.model == OCCM_MODEL_PFR && idx % P == 0;      // keeps the identity, never read
        // whole warps skip together only when both of their groups skip: the shuffles below are group-wide (width W)
        const bool work = gv && b.need_factor[m] && !inlet_node;
        const bool on = work && lane < S;
        const long long e = (long long)m * b.ndof + idx * S;
        const double y = on ? b.y[e + lane] : 0.0;
        const double qs = b.q_scale ? b.q_scale[m] : 1.0;
        const double* ksc = b.k_scale ? b.k_scale + (long long)m * sys.n_rxn : nullptr;
        const long long model = sys.model == OCCM_MODEL_PFR ? idx / P : idx;
        double row[W], inv[W];
        bdf_group_rows<W>(T, nterm, nfac, S, lane, y, b.c[m], qs * b.tdiag[gv ? model : 0], ksc, row);
        bdf_group_invert<W>(S, lane, row, inv);
        if (on) bdf_binv_store<W>(b.Binv + (e + lane) * RS, S, RS, inv);
    }
}

// ... and apply: out = M^-1 (scale .* in) as one S x S mat-vec per point (PFR: inside the flow-order sweep)
template <int MODEL, int W>
__global__ void __launch_bounds__(128) bdf_precond_apply(const OccmBdfDev b, const double* src, int src_is_V_col_j,
                                                         double* __restrict__ dst, const int* __restrict__ mask, int keep_z) {
    const int S = b.S, P = b.P, RS = bdf_binv_row(S);
    const int lane = threadIdx.x % W, gib = threadIdx.x / W, gpb = 128 / W;
    const long long per_member = MODEL == 0 ? b.n_points : b.n_points / P;
    const long long total = (long long)b.M * per_member;
    const long long rounded = (total + gpb - 1) / gpb * gpb;
    for (long long g = (long long)blockIdx.x * gpb + gib; g < rounded; g += (long long)gridDim.x * gpb) {
        const bool gv = g < total;
        const int m = gv ? (int)(g / per_member) : 0;
        const long long idx = gv ? g - (long long)m * per_member : 0;
        const bool on = gv && (!mask || mask[m]) && lane < S;
        const double* in = src + (src_is_V_col_j ? (long long)b.gmres_j[m] * b.vstride : 0) + (long long)m * b.ndof;
        // inside GMRES (keep_z) the source is the not yet normalised Krylov vector: it is scaled by 1 / h_{j+1,j} (b.eps) on the way
        // in and the normalised value is written back — the separate scaling pass over V is gone
        const double va = keep_z ? b.eps[m] : 1.0;
        double* vin = const_cast<double*>(in);
        if (MODEL == 0) {
            const long long g0 = (long long)m * b.ndof + idx * S + lane;
            double r[W];
            bdf_binv_load<W>(b.Binv + g0 * RS, RS, on, r);
            const double vw = on ? in[idx * S + lane] * va : 0.0;
            if (on && keep_z) vin[idx * S + lane] = vw;
            const double u = on ? b.sc[g0] * vw : 0.0;
            double a0 = 0.0, a1 = 0.0;
#pragma unroll
            for (int j = 0; j < W; j += 2) {
                a0 += r[j] * __shfl_sync(0xffffffffu, u, j, W);
                a1 += r[j + 1] * __shfl_sync(0xffffffffu, u, j + 1, W);
            }
            if (on) { dst[g0] = a0 + a1; if (keep_z) b.Z[(long long)b.gmres_j[m] * b.vstride + g0] = a0 + a1; }
        } else {
            const double c = b.c[m];
            const double qs = b.q_scale ? b.q_scale[m] : 1.0;
            const double ca = c * qs * (-b.tdiag[gv ? idx : 0]);             // c * Q/dV of this PFR
            const long long gbase = (long long)m * b.ndof + idx * P * S + lane;
            const double* inp = in + idx * P * S + lane;
            const occm_binv_t* bi = b.Binv + gbase * RS;                       // this lane's row of the PFR's first point
            double* inpm = vin + idx * P * S + lane;
            double xprev = 0.0;
            // The sweep along a PFR is sequential (x_p needs x_{p-1}) and a point's mat-vec takes ~150 cycles, a fraction of the
            // memory latency: the block rows and operands of the next BDF_SWEEP_AHEAD points are kept in flight in a register ring
            // of raw 16-byte vectors (C5, ms per application: 1 ahead 0.406, 2 ahead 0.382 = 5.0 TB/s, 4 ahead 0.477; rows unpacked to
            // doubles one point ahead, the previous form: 0.531).
            constexpr int D = BDF_SWEEP_AHEAD, NV = W / BDF_BINV_VEC;
            int4 q[D][NV];
            double sq[D], wq[D];
#pragma unroll
            for (int s = 0; s < D; ++s) {
                const bool ld = on && s < P;
                bdf_binv_raw<W>(bi + (long long)s * S * RS, RS, ld, q[s]);
                sq[s] = ld ? b.sc[gbase + (long long)s * S] : 0.0;
                wq[s] = ld ? inp[(long long)s * S] : 0.0;
            }
            for (int i0 = 0; i0 < P; i0 += D) {
#pragma unroll
                for (int s = 0; s < D; ++s) {
                    const int i = i0 + s;
                    if (i < P) {                                  // uniform
                        double r[W];
                        bdf_binv_unpack<W>(q[s], r);
                        const double vw = wq[s] * va;
                        if (on && keep_z) inpm[(long long)i * S] = vw;
                        double u = sq[s] * vw;
                        const bool more = on && i + D < P;        // refill the slot with the point D ahead
                        bdf_binv_raw<W>(bi + (long long)(i + D) * S * RS, RS, more, q[s]);
                        sq[s] = more ? b.sc[gbase + (long long)(i + D) * S] : 0.0;
                        wq[s] = more ? inp[(long long)(i + D) * S] : 0.0;
                        if (i > 0) {                              // the inlet node keeps the identity block
                            u += ca * xprev;
                            double a0 = 0.0, a1 = 0.0;
#pragma unroll
                            for (int j = 0; j < W; j += 2) {
                                a0 += r[j] * __shfl_sync(0xffffffffu, u, j, W);
                                a1 += r[j + 1] * __shfl_sync(0xffffffffu, u, j + 1, W);
                            }
                            u = a0 + a1;
                        }
                        if (on) { dst[gbase + (long long)i * S] = u; if (keep_z) b.Z[(long long)b.gmres_j[m] * b.vstride + gbase + (long long)i * S] = u; }
                        xprev = u;
                    }
                }
            }
        }
    }
}

// partial[0] = |v|^2
__global__ void __launch_bounds__(256) bdf_norm2(const OccmBdfDev b, const double* __restrict__ v, const int* __restrict__ mask) {
    __shared__ double scratch[8];
    BDF_FOR_CHUNKS(b, mask[m]) {
        double s0 = 0.0;
        BDF_ELEMS(b, ci, e) { const double x = v[(long long)m * b.ndof + e]; s0 += x * x; }
        const double t0 = occm_block_sum(s0, scratch);
        if (threadIdx.x == 0) b.partial[BDF_PIDX(b, m, ci, 0)] = t0;
    }
}

// w~ = (z - c (f2 - f) / eps) / scale -> V[j+1];  partial[i] = <w~, V[i]>, i = 0..j;  partial[j+1] = |w~|^2
__global__ void __launch_bounds__(256) bdf_arnoldi_w(const OccmBdfDev b) {
    __shared__ double scratch[8];
    BDF_FOR_CHUNKS(b, b.gmres_active[m]) {
        const int j = b.gmres_j[m];
        const double c = b.c[m], inv_eps = 1.0 / b.eps[m];
        double dots[BDF_MAXDOT];
        for (int i = 0; i <= j + 1; ++i) dots[i] = 0.0;
        BDF_ELEMS(b, ci, e) {
            const long long g = (long long)m * b.ndof + e;
            const double w = (b.z[g] - c * ((b.f2[g] - b.f[g]) * inv_eps)) / b.sc[g];
            b.V[(long long)(j + 1) * b.vstride + g] = w;
            for (int i = 0; i <= j; ++i) dots[i] += w * b.V[(long long)i * b.vstride + g];
            dots[j + 1] += w * w;
        }
        for (int i = 0; i <= j + 1; ++i) {
            const double tsum = occm_block_sum(dots[i], scratch);
            if (threadIdx.x == 0) b.partial[BDF_PIDX(b, m, ci, i)] = tsum;
        }
    }
}

// V[j+1] -= sum_i h_i V[i]  (h_i = norms[m][i]);  partial[0] = |V[j+1]|^2, partial[1 + i] = <V[j+1], V[i]> (re-orthogonalisation)
__global__ void __launch_bounds__(256) bdf_orth(const OccmBdfDev b, int second_pass) {
    __shared__ double scratch[8];
    BDF_FOR_CHUNKS(b, b.gmres_active[m] && (!second_pass || b.reorth[m])) {
        const int j = b.gmres_j[m];
        const double* h = b.norms + (long long)m * BDF_MAXDOT + (second_pass ? 1 : 0);
        double s0 = 0.0;
        double dots[BDF_MAXDOT];
        for (int i = 0; i <= j; ++i) dots[i] = 0.0;
        BDF_ELEMS(b, ci, e) {
            const long long g = (long long)m * b.ndof + e;
            double w = b.V[(long long)(j + 1) * b.vstride + g];
            for (int i = 0; i <= j; ++i) w -= h[i] * b.V[(long long)i * b.vstride + g];
            b.V[(long long)(j + 1) * b.vstride + g] = w;
            s0 += w * w;
            if (!second_pass) for (int i = 0; i <= j; ++i) dots[i] += w * b.V[(long long)i * b.vstride + g];
        }
        const double t0 = occm_block_sum(s0, scratch);
        if (threadIdx.x == 0) b.partial[BDF_PIDX(b, m, ci, 0)] = t0;
        if (!second_pass)
            for (int i = 0; i <= j; ++i) {
                const double tsum = occm_block_sum(dots[i], scratch);
                if (threadIdx.x == 0) b.partial[BDF_PIDX(b, m, ci, 1 + i)] = tsum;
            }
    }
}

// dy = z = sum_k ycoef[m][k] Z[k], k < n_iter_gmres[m]: the Newton correction straight from the stored preconditioned Krylov
// vectors z_k = M^-1 S v_k (flexible GMRES) — no preconditioner application after the solve (one of ~3.3 per Newton iteration on C5)
__global__ void __launch_bounds__(256) bdf_lincomb_V(const OccmBdfDev b) {
    BDF_FOR_CHUNKS(b, b.newton_active[m]) {
        const int n = b.gmres_j[m];             // number of Arnoldi vectors used
        const double* yc = b.ycoef + (long long)m * BDF_MAXDOT;
        BDF_ELEMS(b, ci, e) {
            const long long g = (long long)m * b.ndof + e;
            double acc = 0.0;
            for (int k = 0; k < n; ++k) acc += yc[k] * b.Z[(long long)k * b.vstride + g];
            b.z[g] = acc;
        }
    }
}

// dy = z: partial[0] = |dy / scale|^2 (scale from y_predict = y - d);  y += dy;  d += dy      (bdf.py:47, 58-59)
__global__ void __launch_bounds__(256) bdf_newton_update(const OccmBdfDev b) {
    __shared__ double scratch[8];
    BDF_FOR_CHUNKS(b, b.newton_active[m]) {
        const bool zero = b.gmres_j[m] == 0;     // right-hand side already below the linear tolerance: dy = 0
        double s0 = 0.0;
        BDF_ELEMS(b, ci, e) {
            const long long g = (long long)m * b.ndof + e;
            const double y = b.y[g], d = b.d[g];
            const double scale = b.atol + b.rtol * fabs(y - d);
            const double dy = zero ? 0.0 : b.z[g];
            const double q = dy / scale;
            s0 += q * q;
            b.y[g] = y + dy; b.d[g] = d + dy;
        }
        const double t0 = occm_block_sum(s0, scratch);
        if (threadIdx.x == 0) b.partial[BDF_PIDX(b, m, ci, 0)] = t0;
    }
}

// partial[0] = |error_const[order] * d / (atol + rtol |y_new|)|^2        (bdf.py:394-396)
__global__ void __launch_bounds__(256) bdf_error(const OccmBdfDev b) {
    __shared__ double scratch[8];
    BDF_FOR_CHUNKS(b, b.attempt_ok[m]) {
        const double ec = bdf_error_const(b.order[m]);
        double s0 = 0.0;
        BDF_ELEMS(b, ci, e) {
            const long long g = (long long)m * b.ndof + e;
            const double q = ec * b.d[g] / (b.atol + b.rtol * fabs(b.y[g]));
            s0 += q * q;
        }
        const double t0 = occm_block_sum(s0, scratch);
        if (threadIdx.x == 0) b.partial[BDF_PIDX(b, m, ci, 0)] = t0;
    }
}

// accepted step: D[o+2] = d - D[o+1]; D[o+1] = d; D[i] += D[i+1] (i = o..0);                      (bdf.py:419-422)
// partial[0] = |error_const[o-1] D[o] / scale|^2, partial[1] = |error_const[o+1] D[o+2] / scale|^2   (:427-437)
template <int O>
__device__ __forceinline__ void bdf_update_D_elems(const OccmBdfDev& b, int m, int ci, double ecm, double ecp, double& s0, double& s1) {
    BDF_ELEMS(b, ci, e) {
        const long long g = (long long)m * b.ndof + e;
        const double d = b.d[g];
        double Dv[O + 3];
#pragma unroll
        for (int i = 0; i <= O + 1; ++i) Dv[i] = b.D[i * b.vstride + g];
        Dv[O + 2] = d - Dv[O + 1];
        Dv[O + 1] = d;
#pragma unroll
        for (int i = O; i >= 0; --i) Dv[i] += Dv[i + 1];
#pragma unroll
        for (int i = 0; i <= O + 2; ++i) b.D[i * b.vstride + g] = Dv[i];
        const double scale = b.atol + b.rtol * fabs(b.y[g]);
        const double qm = ecm * Dv[O] / scale, qp = ecp * Dv[O + 2] / scale;
        s0 += qm * qm; s1 += qp * qp;
    }
}

__global__ void __launch_bounds__(256) bdf_update_D(const OccmBdfDev b) {
    __shared__ double scratch[8];
    BDF_FOR_CHUNKS(b, b.accepted_now[m]) {
        const int o = b.order[m];
        const double ecm = o > 1 ? bdf_error_const(o - 1) : 0.0, ecp = o < BDF_MAX_ORDER ? bdf_error_const(o + 1) : 0.0;
        double s0 = 0.0, s1 = 0.0;
        switch (o) {
            case 1: bdf_update_D_elems<1>(b, m, ci, ecm, ecp, s0, s1); break;
            case 2: bdf_update_D_elems<2>(b, m, ci, ecm, ecp, s0, s1); break;
            case 3: bdf_update_D_elems<3>(b, m, ci, ecm, ecp, s0, s1); break;
            case 4: bdf_update_D_elems<4>(b, m, ci, ecm, ecp, s0, s1); break;
            default: bdf_update_D_elems<5>(b, m, ci, ecm, ecp, s0, s1); break;
        }
        const double t0 = occm_block_sum(s0, scratch), t1 = occm_block_sum(s1, scratch);
        if (threadIdx.x == 0) { b.partial[BDF_PIDX(b, m, ci, 0)] = t0; b.partial[BDF_PIDX(b, m, ci, 1)] = t1; }
    }
}

// BdfDenseOutput (bdf.py:458-483) at the t_eval points inside the accepted step; "all" mode stores y_new.
__global__ void __launch_bounds__(256) bdf_output(const OccmBdfDev b, int n_save, const int* __restrict__ save_idx,
                                                  const double* __restrict__ t_eval, int n_t_eval, double* __restrict__ out,
                                                  double* __restrict__ out_t) {
    const int cps = (n_save + 255) / 256;
    const long long total = (long long)b.M * cps;
    for (long long cc = blockIdx.x; cc < total; cc += gridDim.x) {
        const int m = (int)(cc / cps);
        if (!b.accepted_now[m]) continue;
        const int lo = b.out_lo[m], hi = t_eval ? b.out_hi[m] : lo + 1;
        if (hi <= lo) continue;
        const int i = (int)(cc - (long long)m * cps) * 256 + threadIdx.x;
        if (i >= n_save) continue;
        const long long e = save_idx ? save_idx[i] : i;
        const long long g = (long long)m * b.ndof + e;
        const double t = b.t[m];
        if (!t_eval) {
            out[((long long)m * n_t_eval + lo) * n_save + i] = b.D[g];
            if (out_t && i == 0) out_t[(long long)m * n_t_eval + lo] = t;
            continue;
        }
        const int o = b.order[m];
        const double h = b.h_abs[m];
        double Dv[6];
        for (int k = 0; k <= o; ++k) Dv[k] = b.D[k * b.vstride + g];
        for (int j = lo; j < hi; ++j) {
            const double te = t_eval[j];
            double p = 1.0, yv = Dv[0];
            for (int k = 0; k < o; ++k) {                  // x_k = (t - (t_n - h k)) / (h (1 + k)); p = cumprod(x)
                p *= (te - (t - h * k)) / (h * (1 + k));
                yv += Dv[k + 1] * p;
            }
            out[((long long)m * n_t_eval + j) * n_save + i] = yv;
            if (out_t && i == 0) out_t[(long long)m * n_t_eval + j] = te;
        }
    }
}

__global__ void bdf_init_D(const OccmBdfDev b, const double* __restrict__ y0) {       // D[0] = y0, D[1] = f h   (bdf.py:257-259)
    BDF_FOR_CHUNKS(b, true) {
        const double h = b.h_abs[m];
        BDF_ELEMS(b, ci, e) {
            const long long g = (long long)m * b.ndof + e;
            b.D[g] = y0[g];
            b.D[b.vstride + g] = b.f[g] * h;
        }
    }
}

__global__ void bdf_write_initial(const OccmBdfDev b, int n_save, const int* __restrict__ save_idx, int cap, double t0,
                                  double* __restrict__ out, double* __restrict__ out_t) {
    const long long total = (long long)b.M * n_save;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
        const int m = (int)(k / n_save);
        const int i = (int)(k - (long long)m * n_save);
        const long long e = save_idx ? save_idx[i] : i;
        out[((long long)m * cap) * n_save + i] = b.D[(long long)m * b.ndof + e];
        if (i == 0) out_t[(long long)m * cap] = t0;
    }
}

// ------------------------------------------------------------------------------------------------ per-member control
__global__ void bdf_ctrl_init(const OccmBdfDev b, double t0, double first_step, int all_mode) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= b.M) return;
    b.t[m] = t0; b.h_abs[m] = first_step; b.order[m] = 1; b.n_equal[m] = 0; b.status[m] = OCCM_MEMBER_RUNNING; b.fresh[m] = 1;
    b.newton_active[m] = 1;            // mask of the initial f(t0, y0) evaluation
    b.t_new[m] = t0;
    b.nfev[m] = 1; b.n_acc[m] = 0; b.n_rej[m] = 0; b.n_newton_fail[m] = 0; b.n_gmres[m] = 0; b.n_precond[m] = 0;
    b.accepted_now[m] = 0; b.out_lo[m] = 0; b.out_hi[m] = 0; b.n_rows[m] = all_mode ? 1 : 0; b.need_change[m] = 0;
    b.c_fact[m] = 0.0; b.need_factor[m] = 0; b.since_factor[m] = 0; b.force_factor[m] = 0; b.n_factor[m] = 0;
    b.crate[m] = 1.0;
}

// start of a step attempt (bdf.py:318-356): step-size clipping, t_new, h, c; may request a change_D
__global__ void bdf_ctrl_begin(const OccmBdfDev b) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= b.M) return;
    b.need_change[m] = 0; b.accepted_now[m] = 0; b.attempt_ok[m] = 0; b.newton_active[m] = 0; b.gmres_active[m] = 0;
    b.need_factor[m] = 0;
    if (b.status[m] != OCCM_MEMBER_RUNNING) return;
    const double t = b.t[m];
    const int order = b.order[m];
    double h_abs = b.h_abs[m];
    const double h_abs0 = h_abs;
    bool clipped = false;
    const double min_step = 10.0 * fabs(bdf_next_up(t) - t);
    if (b.fresh[m] && h_abs < min_step) { h_abs = min_step; b.n_equal[m] = 0; clipped = true; }
    b.fresh[m] = 0;
    if (h_abs < min_step) { b.status[m] = OCCM_MEMBER_STEP_TOO_SMALL; return; }
    double t_new = t + h_abs;
    if (t_new - b.t_bound > 0.0) { t_new = b.t_bound; b.n_equal[m] = 0; clipped = true; }
    const double h = t_new - t;
    h_abs = fabs(h);
    if (clipped) bdf_set_change(b, m, order, h_abs / h_abs0);     // the clip events of bdf.py:321-326 and :348-352 in one rescale
    const double c = h / bdf_alpha(order);
    b.h_abs[m] = h_abs; b.t_new[m] = t_new; b.c[m] = c;
    if (b.store) {
        // when to rebuild the stored block inverses (the setup policy of CVODE's Newton iteration: gamma ratio, age, failure)
        const double cf = b.c_fact[m];
        const int age = b.since_factor[m];
        if (cf == 0.0 || fabs(c / cf - 1.0) > 0.2 || age >= 20 || b.force_factor[m]) {
            b.need_factor[m] = 1; b.c_fact[m] = c; b.since_factor[m] = 0; b.force_factor[m] = 0; b.n_factor[m] += 1;
            b.crate[m] = 1.0;                   // new preconditioner: the carried rate no longer describes the iteration (CVODE cvNlsNewton)
            atomicAdd(&b.counters[3], 1);
        } else b.since_factor[m] = age + 1;
    }
    b.newton_active[m] = 1; b.newton_k[m] = 0; b.converged[m] = 0; b.n_iter[m] = 0; b.dy_norm_old[m] = -1.0;
    atomicAdd(&b.counters[1], 1);
}

// after bdf_newton_rhs + reduce: start GMRES for the member or skip it when the scaled residual is already tiny
__global__ void bdf_ctrl_gmres_start(const OccmBdfDev b) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= b.M) return;
    b.gmres_active[m] = 0;
    if (!b.newton_active[m]) return;
    const double* nr = b.norms + (long long)m * BDF_MAXDOT;
    const double beta = sqrt(nr[0]);
    b.ynorm[m] = sqrt(nr[1]);
    b.gmres_j[m] = 0;
    // non-finite right-hand side (bdf.py:52-53: `if not np.all(np.isfinite(f)): break`): the Newton iteration stops without
    // convergence and the step is halved (bdf_ctrl_newton sees converged == -1)
    if (!isfinite(beta)) { b.converged[m] = -1; return; }
    if (!(beta / sqrt((double)b.ndof) > b.lin_tol)) return;       // residual already below the linear tolerance: dy = 0
    b.gmres_active[m] = 1;
    double* gv = b.gv + (long long)m * BDF_MAXDOT;
    for (int i = 0; i <= b.restart; ++i) gv[i] = 0.0;
    gv[0] = beta;
    b.eps[m] = 1.0 / beta;              // the first preconditioner application scales V0 by it (v0 = b~ / beta)
    atomicAdd(&b.counters[2], 1);
}

// after the preconditioner: step of the directional difference  eps = sqrt(EPS) (1 + |y|) / |z|
__global__ void bdf_ctrl_eps(const OccmBdfDev b) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= b.M || !b.gmres_active[m]) return;
    const double zn = sqrt(b.norms[(long long)m * BDF_MAXDOT]);
    b.eps[m] = zn > 0.0 ? 1.4901161193847656e-08 * (1.0 + b.ynorm[m]) / zn : 1.0;
    b.nfev[m] += 1; b.n_precond[m] += 1;
}

// first Gram-Schmidt pass done: norms[0..j] hold h_i; move them behind slot 0 and remember them in H
__global__ void bdf_ctrl_h1(const OccmBdfDev b) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= b.M || !b.gmres_active[m]) return;
    const int j = b.gmres_j[m];
    double* nr = b.norms + (long long)m * BDF_MAXDOT;
    double* Hc = b.H + ((long long)m * BDF_MAXDOT + j) * BDF_MAXDOT;       // column j of the Hessenberg matrix
    for (int i = 0; i <= j; ++i) Hc[i] = nr[i];
    b.wn2[m] = nr[j + 1];
}

// second pass: norms[0] = |w|^2 after pass 1, norms[1..j+1] = correction coefficients; accumulate into H.
// The pass is only made when the first one cancelled more than half of |w|^2 (Daniel-Gragg-Kaufman-Stewart test).
__global__ void bdf_ctrl_h2(const OccmBdfDev b) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= b.M) return;
    b.reorth[m] = 0;
    if (!b.gmres_active[m]) return;
    const int j = b.gmres_j[m];
    const double* nr = b.norms + (long long)m * BDF_MAXDOT;
    if (!(nr[0] < BDF_REORTH_ETA2 * b.wn2[m])) return;
    b.reorth[m] = 1;
    double* Hc = b.H + ((long long)m * BDF_MAXDOT + j) * BDF_MAXDOT;
    for (int i = 0; i <= j; ++i) Hc[i] += nr[1 + i];
}

// norms[0] = |V[j+1]|^2 after re-orthogonalisation: Givens update, convergence test   (standard GMRES)
__global__ void bdf_ctrl_givens(const OccmBdfDev b) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= b.M || !b.gmres_active[m]) return;
    const int j = b.gmres_j[m];
    double* Hc = b.H + ((long long)m * BDF_MAXDOT + j) * BDF_MAXDOT;
    double* cs = b.cs + (long long)m * BDF_MAXDOT; double* sn = b.sn + (long long)m * BDF_MAXDOT;
    double* gv = b.gv + (long long)m * BDF_MAXDOT;
    const double hn = sqrt(b.norms[(long long)m * BDF_MAXDOT]);
    Hc[j + 1] = hn;
    for (int i = 0; i < j; ++i) {
        const double a = cs[i] * Hc[i] + sn[i] * Hc[i + 1];
        Hc[i + 1] = -sn[i] * Hc[i] + cs[i] * Hc[i + 1];
        Hc[i] = a;
    }
    const double r = hypot(Hc[j], Hc[j + 1]);
    cs[j] = r > 0.0 ? Hc[j] / r : 1.0; sn[j] = r > 0.0 ? Hc[j + 1] / r : 0.0;
    Hc[j] = r; Hc[j + 1] = 0.0;
    gv[j + 1] = -sn[j] * gv[j];
    gv[j] = cs[j] * gv[j];
    b.n_gmres[m] += 1;
    const double res = fabs(gv[j + 1]) / sqrt((double)b.ndof);
    b.eps[m] = hn > 0.0 ? 1.0 / hn : 0.0;           // the next preconditioner application scales V[j+1] by it
    b.gmres_j[m] = j + 1;
    if (!(res > b.lin_tol) || j + 1 >= b.restart || !(hn > 0.0)) {
        // back substitution R y = g
        const int n = j + 1;
        double* yc = b.ycoef + (long long)m * BDF_MAXDOT;
        for (int i = n - 1; i >= 0; --i) {
            double acc = gv[i];
            for (int k = i + 1; k < n; ++k) acc -= b.H[((long long)m * BDF_MAXDOT + k) * BDF_MAXDOT + i] * yc[k];
            yc[i] = acc / b.H[((long long)m * BDF_MAXDOT + i) * BDF_MAXDOT + i];
        }
        b.gmres_active[m] = 0;
        b.converged[m] = res > b.lin_tol ? -1 : 0;      // -1: linear solve stalled -> treated as Newton failure
    } else {
        atomicAdd(&b.counters[2], 1);
    }
}

// bdf.py:47-69: convergence logic of the simplified Newton iteration
__global__ void bdf_ctrl_newton(const OccmBdfDev b) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= b.M || !b.newton_active[m]) return;
    const int k = b.newton_k[m];
    const double dy_norm = sqrt(b.norms[(long long)m * BDF_MAXDOT] / (double)b.ndof);
    const double old = b.dy_norm_old[m];
    const bool have_rate = old >= 0.0;
    const double rate = have_rate ? dy_norm / old : 0.0;
    b.n_iter[m] = k + 1;
    b.nfev[m] += 1;
    bool stop = false, conv = false;
    // Optional (b.carry): CVODE's carried rate, crate = max(0.3 crate, |dy_k| / |dy_k-1|), kept from step to step and reset to 1
    // when the preconditioner is rebuilt or a step fails.  scipy can only declare convergence once it has a rate, i.e. after
    // two iterations (bdf.py:60-66).
    double crate = b.crate[m];
    if (have_rate) { crate = fmax(0.3 * crate, rate); b.crate[m] = crate; }
    if (b.converged[m] == -1 || !isfinite(dy_norm)) { stop = true; }
    else if (have_rate && (rate >= 1.0 || pow(rate, (double)(BDF_NEWTON_MAXITER - k)) / (1.0 - rate) * dy_norm > b.newton_tol)) {
        // scipy breaks BEFORE applying dy; the update kernel already applied it, which is harmless: the attempt is
        // discarded (step halved, predictor recomputed from D)
        stop = true;
    } else if (dy_norm == 0.0 || (have_rate && rate / (1.0 - rate) * dy_norm < b.newton_tol)) {
        stop = true; conv = true;
    } else if (!have_rate && b.carry && crate < 1.0 && crate / (1.0 - crate) * dy_norm < b.newton_tol) {
        stop = true; conv = true;               // first iteration, judged with the carried rate
    }
    b.dy_norm_old[m] = dy_norm;
    if (b.gmres_j[m] >= 6) b.force_factor[m] = 1;        // the stored preconditioner has aged: many Krylov vectors were needed
    b.newton_k[m] = k + 1;
    if (!stop && k + 1 >= BDF_NEWTON_MAXITER) stop = true;
    if (stop) { b.newton_active[m] = 0; b.converged[m] = conv ? 1 : 0; }
    else atomicAdd(&b.counters[1], 1);
}

// after Newton: failure -> halve the step (bdf.py:381-387); success -> safety factor, go on to the error test
__global__ void bdf_ctrl_after_newton(const OccmBdfDev b) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= b.M || b.status[m] != OCCM_MEMBER_RUNNING) return;
    b.need_change[m] = 0;
    if (b.converged[m] == 1) {
        b.attempt_ok[m] = 1;
        b.safety[m] = 0.9 * (2 * BDF_NEWTON_MAXITER + 1) / (double)(2 * BDF_NEWTON_MAXITER + b.n_iter[m]);
    } else {
        b.attempt_ok[m] = 0;
        b.h_abs[m] *= 0.5;
        bdf_set_change(b, m, b.order[m], 0.5);
        b.n_equal[m] = 0;
        b.n_newton_fail[m] += 1; b.n_rej[m] += 1;
        b.force_factor[m] = 1; b.crate[m] = 1.0;
        if (b.max_steps > 0 && b.n_acc[m] + b.n_rej[m] >= b.max_steps) b.status[m] = OCCM_MEMBER_MAX_STEPS;
    }
}

// error test (bdf.py:389-405)
__global__ void bdf_ctrl_error(const OccmBdfDev b, const double* __restrict__ t_eval, int n_t_eval) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= b.M || !b.attempt_ok[m]) return;
    const int order = b.order[m];
    const double error_norm = sqrt(b.norms[(long long)m * BDF_MAXDOT] / (double)b.ndof);
    b.err_norm[m] = error_norm;
    if (error_norm > 1.0 || !isfinite(error_norm)) {
        double factor = b.safety[m] * pow(error_norm, -1.0 / (order + 1));
        factor = factor > BDF_MIN_FACTOR ? factor : BDF_MIN_FACTOR;         // max(MIN_FACTOR, x); NaN -> MIN_FACTOR
        b.h_abs[m] *= factor;
        bdf_set_change(b, m, order, factor);
        b.n_equal[m] = 0;
        b.n_rej[m] += 1;
        if (b.max_steps > 0 && b.n_acc[m] + b.n_rej[m] >= b.max_steps) b.status[m] = OCCM_MEMBER_MAX_STEPS;
        return;
    }
    // accepted
    b.accepted_now[m] = 1;
    b.n_equal[m] += 1;
    b.t_old_out[m] = b.t[m];
    b.t[m] = b.t_new[m];
    b.n_acc[m] += 1;
    b.fresh[m] = 1;
    const double t = b.t[m];
    if (t_eval) {
        int lo = b.out_hi[m], hi = lo;
        while (hi < n_t_eval && t_eval[hi] <= t) ++hi;
        b.out_lo[m] = lo; b.out_hi[m] = hi;
    } else {
        const int row = b.n_rows[m];
        b.out_lo[m] = row; b.n_rows[m] = row + 1;
    }
}

// order / step selection after an accepted step (bdf.py:424-453) and end-of-step status
__global__ void bdf_ctrl_order(const OccmBdfDev b, int all_cap) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= b.M || !b.accepted_now[m]) return;
    b.need_change[m] = 0;
    int order = b.order[m];
    if (b.n_equal[m] >= order + 1) {
        const double* nr = b.norms + (long long)m * BDF_MAXDOT;
        const double n = (double)b.ndof;
        const double em = order > 1 ? sqrt(nr[0] / n) : INFINITY;
        const double ep = order < BDF_MAX_ORDER ? sqrt(nr[1] / n) : INFINITY;
        const double en[3] = {em, b.err_norm[m], ep};
        double fac[3]; int best = 0;
        for (int i = 0; i < 3; ++i) {
            fac[i] = pow(en[i], -1.0 / (order + i));        // error_norms ** (-1 / arange(order, order + 3))
            if (fac[i] > fac[best]) best = i;                // argmax: first maximum
        }
        order += best - 1;
        b.order[m] = order;
        const double f = b.safety[m] * fac[best];
        const double factor = f < BDF_MAX_FACTOR ? f : BDF_MAX_FACTOR;
        b.h_abs[m] *= factor;
        bdf_set_change(b, m, order, factor);
        b.n_equal[m] = 0;
    }
    int status = OCCM_MEMBER_RUNNING;
    if (b.t[m] - b.t_bound >= 0.0) status = OCCM_MEMBER_FINISHED;
    else if (b.max_steps > 0 && b.n_acc[m] + b.n_rej[m] >= b.max_steps) status = OCCM_MEMBER_MAX_STEPS;
    else if (all_cap > 0 && b.n_rows[m] >= all_cap) status = OCCM_MEMBER_PAUSED;
    b.status[m] = status;
}

__global__ void bdf_clear_change(const OccmBdfDev b) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m < b.M) b.need_change[m] = 0;
}

__global__ void bdf_count(const OccmBdfDev b) {
    __shared__ int cnt;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    int local = 0;
    for (int m = threadIdx.x; m < b.M; m += blockDim.x) local += b.status[m] == OCCM_MEMBER_RUNNING;
    atomicAdd(&cnt, local);
    __syncthreads();
    if (threadIdx.x == 0) b.counters[0] = cnt;
}

__global__ void bdf_resume_kernel(const OccmBdfDev b) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= b.M) return;
    b.n_rows[m] = 0;
    if (b.status[m] == OCCM_MEMBER_PAUSED) b.status[m] = OCCM_MEMBER_RUNNING;
}

// ------------------------------------------------------------------------------------------------ host side
static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

struct BdfLayout { size_t total; size_t off[64]; int n; };

static size_t bdf_carve(const occm_system* sys, int M, int restart, int store, OccmBdfDev* d, char* w) {
    size_t o = 0;
    const size_t vec = al256((size_t)M * sys->dev.ndof * sizeof(double));
    auto take = [&](size_t bytes) { size_t r = o; o += al256(bytes); return w ? (void*)(w + r) : (void*)nullptr; };
    const int nchunk = (int)((sys->dev.ndof + BDF_CHUNK - 1) / BDF_CHUNK);
    OccmBdfDev t; memset(&t, 0, sizeof(t));
    OccmBdfDev& b = d ? *d : t;
    b.D = (double*)take(vec * BDF_NVEC_D);
    b.y = (double*)take(vec); b.psi = (double*)take(vec); b.d = (double*)take(vec); b.f = (double*)take(vec);
    b.f2 = (double*)take(vec); b.z = (double*)take(vec); b.w = (double*)take(vec); b.sc = (double*)take(vec);
    b.V = (double*)take(vec * (restart + 1));
    b.Z = (double*)take(vec * restart);
    const size_t md = (size_t)M * 8, mi = (size_t)M * 4, mm = (size_t)M * BDF_MAXDOT * 8;
    b.t = (double*)take(md); b.h_abs = (double*)take(md); b.t_new = (double*)take(md); b.c = (double*)take(md); b.t_old_out = (double*)take(md);
    b.order = (int*)take(mi); b.n_equal = (int*)take(mi); b.status = (int*)take(mi); b.fresh = (int*)take(mi);
    b.newton_active = (int*)take(mi); b.newton_k = (int*)take(mi); b.converged = (int*)take(mi); b.n_iter = (int*)take(mi);
    b.dy_norm_old = (double*)take(md);
    b.gmres_active = (int*)take(mi); b.gmres_j = (int*)take(mi);
    b.H = (double*)take(mm * BDF_MAXDOT); b.cs = (double*)take(mm); b.sn = (double*)take(mm); b.gv = (double*)take(mm);
    b.ycoef = (double*)take(mm); b.eps = (double*)take(md); b.ynorm = (double*)take(md);
    b.RU = (double*)take((size_t)M * 36 * 8); b.need_change = (int*)take(mi); b.attempt_ok = (int*)take(mi);
    b.safety = (double*)take(md); b.err_norm = (double*)take(md);
    b.partial = (double*)take((size_t)M * nchunk * BDF_MAXDOT * 8); b.norms = (double*)take(mm);
    b.accepted_now = (int*)take(mi); b.out_lo = (int*)take(mi); b.out_hi = (int*)take(mi); b.n_rows = (int*)take(mi);
    b.nfev = (long long*)take(md); b.n_acc = (long long*)take(md); b.n_rej = (long long*)take(md);
    b.n_newton_fail = (long long*)take(md); b.n_gmres = (long long*)take(md); b.n_precond = (long long*)take(md);
    b.counters = (int*)take(256);
    b.c_fact = (double*)take(md); b.need_factor = (int*)take(mi); b.since_factor = (int*)take(mi); b.force_factor = (int*)take(mi);
    b.n_factor = (long long*)take(md);
    b.reorth = (int*)take(mi); b.wn2 = (double*)take(md); b.crate = (double*)take(md);
    b.Binv = store ? (occm_binv_t*)take((size_t)M * sys->dev.ndof * bdf_binv_row(sys->dev.S) * sizeof(occm_binv_t)) : nullptr;
    b.nchunk = nchunk;
    return o;
}

extern "C" size_t occm_bdf_workspace_bytes(const occm_system* sys, int32_t M, int32_t restart, int32_t precond_mode) {
    if (!sys || M < 1) return 0;
    if (restart <= 0) restart = 30;
    return bdf_carve(sys, M, restart, precond_mode == 0, nullptr, nullptr);
}

static int bdf_grid(const occm_system* sys, long long chunks) {
    long long g = (long long)sys->sm_count * 8;
    if (g > chunks) g = chunks;
    return (int)(g < 1 ? 1 : g);
}

#define BDF_LAUNCH(name, grid, block, smem, ...)                                   \
    do {                                                                           \
        name<<<grid, block, smem, st>>>(__VA_ARGS__);                              \
        OCCM_LAUNCH_CHECK(#name);                                                  \
    } while (0)

static int bdf_reduce_launch(occm_bdf* h, int nvals, const int* mask, cudaStream_t st) {
    BDF_LAUNCH(bdf_reduce, dim3(h->M, nvals), 512, 0, h->d.partial, h->d.nchunk, nvals, h->d.norms, mask);
    return OCCM_OK;
}

static int bdf_rhs(occm_bdf* h, const int* mask, const double* z, const double* eps, double* out, cudaStream_t st) {
    OccmMemDev md = occm_mem_dev(h->has_mem ? &h->mem : nullptr, 1);
    md.active = mask; md.z = z; md.eps = eps;
    return occm_launch_rhs(h->sys, md, h->d.t_new, 0.0, h->d.y, out, st);
}

template <int MODEL, int W>
static int bdf_precond_launch_w(occm_bdf* h, const double* src, int from_col_j, double* dst, const int* mask, int keep_z, cudaStream_t st) {
    const occm_system* sys = h->sys;
    const size_t smem = sys->smem_tables;
    static OccmLaunchCache cache;
    if (int rc = occm_kernel_prepare(&cache, (const void*)bdf_precond_warp<MODEL, W>, sys->device, 128, smem, nullptr)) return rc;
    const long long groups = (long long)h->M * (MODEL == 0 ? sys->dev.n_points : sys->dev.n_models);
    const int gpb = 128 / W;
    long long blocks = (groups + gpb - 1) / gpb;
    const long long cap = (long long)sys->sm_count * 16;
    if (blocks > cap) blocks = cap;
    BDF_LAUNCH((bdf_precond_warp<MODEL, W>), (int)blocks, 128, smem, sys->dev, h->d, src, from_col_j, dst, mask, keep_z);
    return OCCM_OK;
}

template <int W>
static int bdf_factor_launch_w(occm_bdf* h, cudaStream_t st) {
    const occm_system* sys = h->sys;
    const size_t smem = sys->smem_tables;
    static OccmLaunchCache cache;
    if (int rc = occm_kernel_prepare(&cache, (const void*)bdf_precond_factor<W>, sys->device, 128, smem, nullptr)) return rc;
    const long long groups = (long long)h->M * sys->dev.n_points;
    const int gpb = 128 / W;
    long long blocks = (groups + gpb - 1) / gpb;
    const long long cap = (long long)sys->sm_count * 16;
    if (blocks > cap) blocks = cap;
    BDF_LAUNCH((bdf_precond_factor<W>), (int)blocks, 128, smem, sys->dev, h->d);
    return OCCM_OK;
}
// rebuild the stored block inverses of the members flagged by bdf_ctrl_begin (at the predictor of this attempt)
static int bdf_factor_launch(occm_bdf* h, cudaStream_t st) {
    const int S = h->sys->dev.S;
    return S <= 4 ? bdf_factor_launch_w<4>(h, st) : S <= 8 ? bdf_factor_launch_w<8>(h, st) : S <= 16 ? bdf_factor_launch_w<16>(h, st) : bdf_factor_launch_w<32>(h, st);
}

template <int MODEL, int W>
static int bdf_apply_launch_w(occm_bdf* h, const double* src, int from_col_j, double* dst, const int* mask, int keep_z, cudaStream_t st) {
    const occm_system* sys = h->sys;
    const long long groups = (long long)h->M * (MODEL == 0 ? sys->dev.n_points : sys->dev.n_models);
    const int gpb = 128 / W;
    long long blocks = (groups + gpb - 1) / gpb;
    const long long cap = (long long)sys->sm_count * 16;
    if (blocks > cap) blocks = cap;
    BDF_LAUNCH((bdf_precond_apply<MODEL, W>), (int)blocks, 128, 0, h->d, src, from_col_j, dst, mask, keep_z);
    return OCCM_OK;
}

static int bdf_precond_launch(occm_bdf* h, const double* src, int from_col_j, double* dst, const int* mask, int want_norm, cudaStream_t st) {
    const occm_system* sys = h->sys;
    const int S = sys->dev.S;
    int rc;
    if (h->d.store) {
        if (sys->dev.model == OCCM_MODEL_CSTR)
            rc = S <= 4 ? bdf_apply_launch_w<0, 4>(h, src, from_col_j, dst, mask, from_col_j, st) : S <= 8 ? bdf_apply_launch_w<0, 8>(h, src, from_col_j, dst, mask, from_col_j, st)
                                                                                              : S <= 16 ? bdf_apply_launch_w<0, 16>(h, src, from_col_j, dst, mask, from_col_j, st) : bdf_apply_launch_w<0, 32>(h, src, from_col_j, dst, mask, from_col_j, st);
        else
            rc = S <= 4 ? bdf_apply_launch_w<1, 4>(h, src, from_col_j, dst, mask, from_col_j, st) : S <= 8 ? bdf_apply_launch_w<1, 8>(h, src, from_col_j, dst, mask, from_col_j, st)
                                                                                              : S <= 16 ? bdf_apply_launch_w<1, 16>(h, src, from_col_j, dst, mask, from_col_j, st) : bdf_apply_launch_w<1, 32>(h, src, from_col_j, dst, mask, from_col_j, st);
    } else if (sys->dev.model == OCCM_MODEL_CSTR)
        rc = S <= 4 ? bdf_precond_launch_w<0, 4>(h, src, from_col_j, dst, mask, from_col_j, st) : S <= 8 ? bdf_precond_launch_w<0, 8>(h, src, from_col_j, dst, mask, from_col_j, st)
                                                                                           : S <= 16 ? bdf_precond_launch_w<0, 16>(h, src, from_col_j, dst, mask, from_col_j, st) : bdf_precond_launch_w<0, 32>(h, src, from_col_j, dst, mask, from_col_j, st);
    else
        rc = S <= 4 ? bdf_precond_launch_w<1, 4>(h, src, from_col_j, dst, mask, from_col_j, st) : S <= 8 ? bdf_precond_launch_w<1, 8>(h, src, from_col_j, dst, mask, from_col_j, st)
                                                                                           : S <= 16 ? bdf_precond_launch_w<1, 16>(h, src, from_col_j, dst, mask, from_col_j, st) : bdf_precond_launch_w<1, 32>(h, src, from_col_j, dst, mask, from_col_j, st);
    if (rc) return rc;
    if (want_norm) {
        const int gv = bdf_grid(sys, (long long)h->M * h->d.nchunk);
        BDF_LAUNCH(bdf_norm2, gv, 256, 0, h->d, dst, mask);
    }
    return OCCM_OK;
}

static int bdf_read_counters(occm_bdf* h, cudaStream_t st) {
    OCCM_CUDA_CHECK(cudaMemcpyAsync(h->h_counters, h->d.counters, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
    OCCM_CUDA_CHECK(cudaStreamSynchronize(st));
    return OCCM_OK;
}

// device side of occm_bdf_init: member scalars, f(t0, y0), D[0] = y0, D[1] = f h, first output row
static int bdf_init_device(occm_bdf* h, cudaStream_t st) {
    const occm_system* sys = h->sys;
    const occm_bdf_problem* p = &h->prob;
    OccmBdfDev& b = h->d;
    const int M = h->M;
    const int gm = (M + 63) / 64;
    BDF_LAUNCH(bdf_ctrl_init, gm, 64, 0, b, p->t0, p->first_step, p->t_eval ? 0 : 1);
    OCCM_CUDA_CHECK(cudaMemcpyAsync(b.y, p->y0, (size_t)M * b.ndof * sizeof(double), cudaMemcpyDeviceToDevice, st));
    const int rc = bdf_rhs(h, nullptr, nullptr, nullptr, b.f, st);
    if (rc) return rc;
    const int gv = bdf_grid(sys, (long long)M * b.nchunk);
    BDF_LAUNCH(bdf_init_D, gv, 256, 0, b, p->y0);
    if (!p->t_eval) {
        const int n_save = p->save_idx ? p->n_save : (int)b.ndof;
        BDF_LAUNCH(bdf_write_initial, 148 * 4, 256, 0, b, n_save, p->save_idx, p->n_t_eval, p->t0, p->out, p->out_t);
    }
    return OCCM_OK;
}

extern "C" int occm_bdf_init(occm_bdf** out, const occm_system* sys, const occm_members* mem, const occm_bdf_problem* p,
                             void* workspace, size_t workspace_bytes, void* stream) {
    if (!out || !sys || !p || !workspace || !p->y0 || !p->out || !p->tdiag) { occm_set_error("occm_bdf_init: null argument"); return OCCM_ERR_INVALID; }
    const int M = mem ? mem->n_members : 1;
    if (M < 1) { occm_set_error("n_members must be >= 1"); return OCCM_ERR_INVALID; }
    if (sys->dev.S > 32) { occm_set_error("B200-BDF supports up to 32 species (one lane per row of the S x S preconditioner block), got %d", sys->dev.S); return OCCM_ERR_UNSUPPORTED; }
    if (!(p->t_bound > p->t0)) { occm_set_error("t_bound must be greater than t0"); return OCCM_ERR_INVALID; }
    if (!(p->first_step > 0.0)) { occm_set_error("`first_step` must be positive."); return OCCM_ERR_INVALID; }
    if (p->first_step > fabs(p->t_bound - p->t0)) { occm_set_error("`first_step` exceeds bounds."); return OCCM_ERR_INVALID; }
    if (!(p->rtol > 0.0) || !(p->atol >= 0.0)) { occm_set_error("rtol must be > 0 and atol >= 0"); return OCCM_ERR_INVALID; }
    if (p->n_t_eval < 1 || (!p->t_eval && (p->n_t_eval < 2 || !p->out_t))) { occm_set_error("bad output description"); return OCCM_ERR_INVALID; }
    const int restart = p->gmres_restart > 0 ? p->gmres_restart : 30;
    if (restart + 2 > BDF_MAXDOT) { occm_set_error("gmres_restart must be <= %d", BDF_MAXDOT - 2); return OCCM_ERR_INVALID; }
    if (p->precond_mode != 0 && p->precond_mode != 1) { occm_set_error("precond_mode must be 0 (stored block inverses) or 1 (on the fly)"); return OCCM_ERR_INVALID; }
    const int store = p->precond_mode == 0;
    const size_t need = bdf_carve(sys, M, restart, store, nullptr, nullptr);
    if (workspace_bytes < need) { occm_set_error("workspace too small: %zu < %zu", workspace_bytes, need); return OCCM_ERR_WORKSPACE; }
    if ((uintptr_t)workspace & 255) { occm_set_error("workspace must be 256-byte aligned"); return OCCM_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    occm_bdf* h = (occm_bdf*)calloc(1, sizeof(occm_bdf));
    h->sys = sys; h->M = M; h->prob = *p; h->has_mem = mem != nullptr;
    if (mem) h->mem = *mem; else { memset(&h->mem, 0, sizeof(h->mem)); h->mem.n_members = 1; }
    OccmBdfDev& b = h->d;
    memset(&b, 0, sizeof(b));
    bdf_carve(sys, M, restart, store, &b, (char*)workspace);
    b.store = store;
    b.M = M; b.S = sys->dev.S; b.P = sys->dev.P; b.model = sys->dev.model; b.restart = restart;
    b.ndof = sys->dev.ndof; b.n_points = sys->dev.n_points; b.vstride = (long long)M * sys->dev.ndof;
    b.rtol = p->rtol; b.atol = p->atol; b.t_bound = p->t_bound; b.max_steps = p->max_steps;
    const double EPS = 2.220446049250313e-16;
    b.newton_tol = fmax(10.0 * EPS / p->rtol, fmin(0.03, sqrt(p->rtol)));       // bdf.py:224
    b.lin_tol = (p->lin_tol_factor > 0.0 ? p->lin_tol_factor : 0.05) * b.newton_tol;   // CVODE's eps_lin
    b.tdiag = p->tdiag; b.q_scale = h->mem.q_scale; b.k_scale = h->mem.k_scale;
    // 0 (default): scipy's rule, convergence needs a measured rate, i.e. two iterations; 1 (OCCM_B200_BDF_CARRY_RATE=1): CVODE's
    // carried rate.  Measured on C5: 1.54 instead of 2.00 Newton iterations per attempt but the same number of GMRES iterations
    // (the second Newton iteration is cheap), 45.8 s instead of 47.8 s — and on the mid-size C5 parity case 2.4 instead of 1.8
    // acceptance bands from the tight solution.  4 % of time is not worth a third of the accuracy.
    b.carry = getenv("OCCM_B200_BDF_CARRY_RATE") ? 1 : 0;
    h->h_counters = (int*)occm_pinned_acquire();
    if (!h->h_counters) { occm_set_error("host allocation failed"); free(h); return OCCM_ERR_CUDA; }
    const int rc = bdf_init_device(h, st);
    if (rc) { occm_pinned_release(h->h_counters); free(h); return rc; }        // no handle, no pinned slot left behind
    *out = h;
    return OCCM_OK;
}

// one step attempt of every running member
static int bdf_attempt(occm_bdf* h, cudaStream_t st) {
    OccmBdfDev& b = h->d;
    const occm_system* sys = h->sys;
    const int gm = (h->M + 63) / 64;
    const int gv = bdf_grid(sys, (long long)h->M * b.nchunk);
    int rc;
    OCCM_CUDA_CHECK(cudaMemsetAsync(b.counters, 0, 4 * sizeof(int), st));
    BDF_LAUNCH(bdf_ctrl_begin, gm, 64, 0, b);
    BDF_LAUNCH(bdf_change_D, gv, 256, 0, b);
    BDF_LAUNCH(bdf_clear_change, gm, 64, 0, b);
    BDF_LAUNCH(bdf_predict, gv, 256, 0, b);
    if (b.store && (rc = bdf_factor_launch(h, st))) return rc;
    for (int k = 0; k < BDF_NEWTON_MAXITER; ++k) {
        if ((rc = bdf_rhs(h, b.newton_active, nullptr, nullptr, b.f, st))) return rc;          // f = fun(t_new, y)
        BDF_LAUNCH(bdf_newton_rhs, gv, 256, 0, b);
        if ((rc = bdf_reduce_launch(h, 2, b.newton_active, st))) return rc;
        OCCM_CUDA_CHECK(cudaMemsetAsync(b.counters + 2, 0, sizeof(int), st));
        BDF_LAUNCH(bdf_ctrl_gmres_start, gm, 64, 0, b);
        if ((rc = bdf_read_counters(h, st))) return rc;
        int gm_active = h->h_counters[2];
        for (int j = 0; j < b.restart && gm_active > 0; ++j) {
            if ((rc = bdf_precond_launch(h, b.V, 1, b.z, b.gmres_active, 1, st))) return rc;   // z = M^-1 S v_j
            if ((rc = bdf_reduce_launch(h, 1, b.gmres_active, st))) return rc;
            BDF_LAUNCH(bdf_ctrl_eps, gm, 64, 0, b);
            if ((rc = bdf_rhs(h, b.gmres_active, b.z, b.eps, b.f2, st))) return rc;             // f(y + eps z)
            BDF_LAUNCH(bdf_arnoldi_w, gv, 256, 0, b);
            if ((rc = bdf_reduce_launch(h, j + 2, b.gmres_active, st))) return rc;
            BDF_LAUNCH(bdf_ctrl_h1, gm, 64, 0, b);
            BDF_LAUNCH(bdf_orth, gv, 256, 0, b, 0);
            if ((rc = bdf_reduce_launch(h, j + 2, b.gmres_active, st))) return rc;
            BDF_LAUNCH(bdf_ctrl_h2, gm, 64, 0, b);
            BDF_LAUNCH(bdf_orth, gv, 256, 0, b, 1);                                             // re-orthogonalisation where needed
            if ((rc = bdf_reduce_launch(h, 1, b.reorth, st))) return rc;
            OCCM_CUDA_CHECK(cudaMemsetAsync(b.counters + 2, 0, sizeof(int), st));
            BDF_LAUNCH(bdf_ctrl_givens, gm, 64, 0, b);                // leaves 1 / h_{j+1,j} in b.eps: the next preconditioner
                                                                      // application normalises v_{j+1} while it reads it
            if ((rc = bdf_read_counters(h, st))) return rc;
            gm_active = h->h_counters[2];
        }
        BDF_LAUNCH(bdf_lincomb_V, gv, 256, 0, b);                                               // dy = sum y_k z_k
        BDF_LAUNCH(bdf_newton_update, gv, 256, 0, b);
        if ((rc = bdf_reduce_launch(h, 1, b.newton_active, st))) return rc;
        OCCM_CUDA_CHECK(cudaMemsetAsync(b.counters + 1, 0, sizeof(int), st));
        BDF_LAUNCH(bdf_ctrl_newton, gm, 64, 0, b);
        if ((rc = bdf_read_counters(h, st))) return rc;
        if (h->h_counters[1] == 0) break;
    }
    BDF_LAUNCH(bdf_ctrl_after_newton, gm, 64, 0, b);
    BDF_LAUNCH(bdf_error, gv, 256, 0, b);
    if ((rc = bdf_reduce_launch(h, 1, b.attempt_ok, st))) return rc;
    BDF_LAUNCH(bdf_ctrl_error, gm, 64, 0, b, h->prob.t_eval, h->prob.n_t_eval);
    BDF_LAUNCH(bdf_change_D, gv, 256, 0, b);            // rejected / failed attempts rescale D
    BDF_LAUNCH(bdf_clear_change, gm, 64, 0, b);
    BDF_LAUNCH(bdf_update_D, gv, 256, 0, b);
    if ((rc = bdf_reduce_launch(h, 2, b.accepted_now, st))) return rc;
    BDF_LAUNCH(bdf_ctrl_order, gm, 64, 0, b, h->prob.t_eval ? 0 : h->prob.n_t_eval);
    BDF_LAUNCH(bdf_change_D, gv, 256, 0, b);            // accepted members that changed order / step size
    BDF_LAUNCH(bdf_clear_change, gm, 64, 0, b);
    const int n_save = h->prob.save_idx ? h->prob.n_save : (int)b.ndof;
    BDF_LAUNCH(bdf_output, bdf_grid(sys, (long long)h->M * ((n_save + 255) / 256)), 256, 0, b, n_save, h->prob.save_idx, h->prob.t_eval,
               h->prob.n_t_eval, h->prob.out, h->prob.out_t);
    return OCCM_OK;
}

extern "C" int occm_bdf_run(occm_bdf* h, int64_t max_attempts, int32_t* n_running, void* stream) {
    if (!h) { occm_set_error("occm_bdf_run: null handle"); return OCCM_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    long long done = 0;
    int running = h->M;
    while (true) {
        int rc = bdf_attempt(h, st);
        if (rc) return rc;
        ++done;
        bdf_count<<<1, 256, 0, st>>>(h->d);
        OCCM_LAUNCH_CHECK("bdf_count");
        if ((rc = bdf_read_counters(h, st))) return rc;
        running = h->h_counters[0];
        if (running == 0) break;
        if (max_attempts > 0 && done >= max_attempts) break;
    }
    if (n_running) *n_running = running;
    return OCCM_OK;
}

extern "C" int occm_bdf_resume(occm_bdf* h, void* stream) {
    if (!h) { occm_set_error("occm_bdf_resume: null handle"); return OCCM_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    BDF_LAUNCH(bdf_resume_kernel, (h->M + 63) / 64, 64, 0, h->d);
    return OCCM_OK;
}

extern "C" int occm_bdf_stats(occm_bdf* h, int32_t* status, int64_t* n_accepted, int64_t* n_rejected, int64_t* nfev,
                              int64_t* n_gmres, int64_t* n_newton_fail, double* t, int32_t* n_rows, int32_t* order, void* stream) {
    if (!h) { occm_set_error("occm_bdf_stats: null handle"); return OCCM_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    const int M = h->M;
    const OccmBdfDev& b = h->d;
    auto cp = [&](void* dst, const void* src, size_t n) { return dst ? cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, st) : cudaSuccess; };
    OCCM_CUDA_CHECK(cp(status, b.status, M * 4)); OCCM_CUDA_CHECK(cp(n_accepted, b.n_acc, M * 8)); OCCM_CUDA_CHECK(cp(n_rejected, b.n_rej, M * 8));
    OCCM_CUDA_CHECK(cp(nfev, b.nfev, M * 8)); OCCM_CUDA_CHECK(cp(n_gmres, b.n_gmres, M * 8)); OCCM_CUDA_CHECK(cp(n_newton_fail, b.n_newton_fail, M * 8));
    OCCM_CUDA_CHECK(cp(t, b.t, M * 8)); OCCM_CUDA_CHECK(cp(n_rows, h->prob.t_eval ? b.out_hi : b.n_rows, M * 4)); OCCM_CUDA_CHECK(cp(order, b.order, M * 4));
    OCCM_CUDA_CHECK(cudaStreamSynchronize(st));
    return OCCM_OK;
}

extern "C" int occm_bdf_precond_stats(occm_bdf* h, int64_t* n_apply, int64_t* n_factor, void* stream) {
    if (!h) { occm_set_error("occm_bdf_precond_stats: null handle"); return OCCM_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    if (n_apply) OCCM_CUDA_CHECK(cudaMemcpyAsync(n_apply, h->d.n_precond, h->M * 8, cudaMemcpyDeviceToHost, st));
    if (n_factor) OCCM_CUDA_CHECK(cudaMemcpyAsync(n_factor, h->d.n_factor, h->M * 8, cudaMemcpyDeviceToHost, st));
    OCCM_CUDA_CHECK(cudaStreamSynchronize(st));
    return OCCM_OK;
}

extern "C" int occm_bdf_get_state(occm_bdf* h, double* y_out, void* stream) {
    if (!h || !y_out) { occm_set_error("occm_bdf_get_state: null argument"); return OCCM_ERR_INVALID; }
    OCCM_CUDA_CHECK(cudaMemcpyAsync(y_out, h->d.D, (size_t)h->M * h->d.ndof * sizeof(double), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return OCCM_OK;
}

extern "C" int occm_bdf_time_kernel(occm_bdf* h, int32_t which, int32_t reps, float* ms_per_launch, void* stream) {
    if (!h || !ms_per_launch || reps < 1 || which < 0 || which > 2) { occm_set_error("occm_bdf_time_kernel: bad argument"); return OCCM_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    OccmBdfDev& b = h->d;
    cudaEvent_t e0, e1;
    OCCM_CUDA_CHECK(cudaEventCreate(&e0));
    OCCM_CUDA_CHECK(cudaEventCreate(&e1));
    OCCM_CUDA_CHECK(cudaEventRecord(e0, st));
    int rc = OCCM_OK;
    for (int i = 0; i < reps && rc == OCCM_OK; ++i) {
        if (which == 0) rc = bdf_rhs(h, nullptr, nullptr, nullptr, b.f2, st);              // every member, scratch output
        else if (which == 1) rc = bdf_rhs(h, nullptr, b.z, b.eps, b.f2, st);
        else rc = bdf_precond_launch(h, b.V, 0, b.w, nullptr, 0, st);
    }
    OCCM_CUDA_CHECK(cudaEventRecord(e1, st));
    OCCM_CUDA_CHECK(cudaEventSynchronize(e1));
    float ms = 0.f;
    OCCM_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *ms_per_launch = ms / reps;
    return rc;
}

extern "C" int occm_bdf_free(occm_bdf* h) {
    if (!h) return OCCM_OK;
    occm_pinned_release(h->h_counters);
    free(h);
    return OCCM_OK;
}
