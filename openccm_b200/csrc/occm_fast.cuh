// Fast path of the fused stage kernels (included by occm_solver.cu after the generic kernel).
//
// Preconditions (checked on the host, otherwise the generic occm_stage kernel runs): mechanism
// in register form (fast_ok: <= 4 terms per species; WIDE = 1 keeps longer term lists in shared memory instead), packed
// ELL operator of width K <= 4, ndof and all member offsets 16-byte aligned.
// CSTR networks: dy = qs * sum_k w_k y[src_k] + r(y).  PFR networks use the same kernel body with a one-entry row per
// point, {src = upwind point, w = Q/dV}: dy = -(qs w)(y - y[src]) + r(y) (pfr_system.py:296); inlet nodes are flagged.
//
// Differences from the generic kernel:
//   * one thread owns TWO adjacent species of a point: every own-element access is a 128-bit LDG/STG, every row
//     gather a 128-bit load — half the memory instructions per byte;
//   * the ELL operator is packed as 16-byte entries {src, flag, w} (one LDG.128 per entry; the point flag rides in
//     column 0, no separate flag array);
//   * straight-line code templated on K, 32-bit element offsets against per-member base pointers;
//   * software pipelining: while a trip computes, the own values, epilogue operands and ELL entries of the CTA's
//     NEXT trip are already in flight (registers), so only the L2 latency of the row gathers is exposed;
//   * the staged tile is double buffered: one barrier per trip — a __syncwarp in the light stages, where a warp owns whole
//     points (OccmWarpMap), a __syncthreads in the dense block mapping of stages 3-6.
#pragma once
#include <type_traits>

#ifndef OCCM_FAST_CTAS
#define OCCM_FAST_CTAS 2
#endif

struct __align__(16) OccmEllEntry { int src; int flag; double w; };

struct OccmOps2 { double2 ye, k1, ka, kb, kc, kd; };

// V = values per thread: 2 = two adjacent species (128-bit accesses), 1 = one species (odd species counts: 64-bit accesses, the
// .y lane of every double2 is dead code)
template <int V>
__device__ __forceinline__ double2 occm_ld2(const double* p) {
    if (V == 2) return __ldg(reinterpret_cast<const double2*>(p));
    return make_double2(__ldg(p), 0.0);
}
template <int V>
__device__ __forceinline__ void occm_st2(double* p, double2 v) {
    if (V == 2) *reinterpret_cast<double2*>(p) = v; else *p = v.x;
}
__device__ __forceinline__ double2 occm_fma2(double a, double2 x, double2 y) { return make_double2(fma(a, x.x, y.x), fma(a, x.y, y.y)); }

// Per-member context of a trip.  Only the values that cost a global load are kept (6 registers); the array pointers
// are rebuilt where they are used from kernel parameters (constant bank) + m * ndof.
template <int STAGE>
struct OccmFastCtx {
    double h;        // step (stages 2..7) or eps of the directional difference (stage 1)
    double qs;       // flow scale of the member
    int par;         // parity of the state ping-pong
    int m;
};

template <int STAGE>
__device__ __forceinline__ OccmFastCtx<STAGE> occm_fast_ctx(const OccmSysDev& sys, const OccmMemDev& mem, const OccmRkDev& rk, int m) {
    OccmFastCtx<STAGE> c;
    c.m = m; c.h = 0.0; c.par = 0;
    c.qs = mem.q_scale ? __ldg(mem.q_scale + m) : 1.0;
    if (STAGE == 1) c.h = __ldg(mem.eps + m);
    if (STAGE >= 2) { c.h = rk.h[m]; c.par = rk.parity[m]; }
    return c;
}

// base pointers of member c.m
template <int STAGE>
__device__ __forceinline__ const double* occm_fast_in_a(const OccmSysDev& sys, const OccmRkDev& rk, const double* y_in, const OccmFastCtx<STAGE>& c) {
    const long long base = (long long)c.m * sys.ndof;
    return STAGE <= 1 ? y_in + base : STAGE == 2 ? rk.Y[c.par] + base : STAGE == 3 || STAGE == 5 ? rk.YS[0] + base
         : STAGE == 4 || STAGE == 6 ? rk.YS[1] + base : rk.Y[c.par ^ 1] + base;
}
template <int STAGE>
__device__ __forceinline__ const double* occm_fast_in_b(const OccmSysDev& sys, const OccmMemDev& mem, const OccmRkDev& rk, const OccmFastCtx<STAGE>& c) {
    const long long base = (long long)c.m * sys.ndof;
    return STAGE == 1 ? mem.z + base : rk.KF[c.par] + base;
}

template <int STAGE, int V>
__device__ __forceinline__ double2 occm_fast_in(const double* in_a, const double* in_b, double hb, int e) {
    const double2 a = occm_ld2<V>(in_a + e);
    if (STAGE == 2) { const double2 b = occm_ld2<V>(in_b + e); return make_double2(a.x + (dp::A21 * b.x) * hb, a.y + (dp::A21 * b.y) * hb); }
    if (STAGE == 1) { const double2 b = occm_ld2<V>(in_b + e); return make_double2(a.x + b.x * hb, a.y + b.y * hb); }
    return a;
}

template <int STAGE, int V>
__device__ __forceinline__ OccmOps2 occm_fast_load_ops(const OccmSysDev& sys, const OccmRkDev& rk, const OccmFastCtx<STAGE>& c, int e) {
    OccmOps2 o;
    const double2 z = make_double2(0.0, 0.0);
    o.ye = o.k1 = o.ka = o.kb = o.kc = o.kd = z;
    if (STAGE <= 1) return o;
    const long long g = (long long)c.m * sys.ndof + e;
    o.ye = occm_ld2<V>(rk.Y[c.par] + g);
    if (STAGE != 7) o.k1 = occm_ld2<V>(rk.KF[c.par] + g);
    if (STAGE == 3) { o.ka = occm_ld2<V>(rk.K2 + g); }
    if (STAGE == 4) { o.ka = occm_ld2<V>(rk.K2 + g); o.kb = occm_ld2<V>(rk.K3 + g); }
    if (STAGE == 5) { o.ka = occm_ld2<V>(rk.K2 + g); o.kb = occm_ld2<V>(rk.K3 + g); o.kc = occm_ld2<V>(rk.K4 + g); }
    if (STAGE == 6) { o.ka = occm_ld2<V>(rk.K3 + g); o.kb = occm_ld2<V>(rk.K4 + g); o.kc = occm_ld2<V>(rk.K5 + g); }
    if (STAGE == 7) { o.ka = occm_ld2<V>(rk.K2 + g); }        // partial error combination written by stage 6
    return o;
}

// L2 prefetch of the epilogue operands of the NEXT trip (no registers held): when that trip starts, its operand loads are
// L2 hits instead of DRAM round trips.  One lane in eight covers a 128-byte line.
__device__ __forceinline__ void occm_prefetch_l2(const double* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
template <int STAGE>
__device__ __forceinline__ void occm_fast_prefetch_ops(const OccmSysDev& sys, const OccmRkDev& rk, const OccmFastCtx<STAGE>& c, int e) {
    if (STAGE <= 1 || (threadIdx.x & 7) != 0) return;
    const long long g = (long long)c.m * sys.ndof + e;
    occm_prefetch_l2(rk.Y[c.par] + g);
    if (STAGE != 7) occm_prefetch_l2(rk.KF[c.par] + g);
    if (STAGE == 3) { occm_prefetch_l2(rk.K2 + g); }
    if (STAGE == 4) { occm_prefetch_l2(rk.K2 + g); occm_prefetch_l2(rk.K3 + g); }
    if (STAGE == 5) { occm_prefetch_l2(rk.K2 + g); occm_prefetch_l2(rk.K3 + g); occm_prefetch_l2(rk.K4 + g); }
    if (STAGE == 6) { occm_prefetch_l2(rk.K3 + g); occm_prefetch_l2(rk.K4 + g); occm_prefetch_l2(rk.K5 + g); }
    if (STAGE == 7) { occm_prefetch_l2(rk.K2 + g); }
}

// ---- epilogue operands staged through shared memory with cp.async, one trip ahead (Dormand-Prince stages).
// Every thread copies its own 16-byte operands into its own slots and reads them back itself: no barrier, only
// cp.async.wait_group.  Nothing is held in registers while the trip computes (the register version kept up to 20).
#ifndef OCCM_OPS_ASYNC
#define OCCM_OPS_ASYNC 1
#endif
// Measured per stage (C4, B200): staging helps the stages with few operands (2: -1 %, 3: -1 %, 7: -9 %) and costs 2-5 % in
// stages 4-6, whose operands then pass through 40 KB of shared memory per CTA: those keep the register version + L2 prefetch.
template <int STAGE> struct OccmOpsAsync { static constexpr bool value = OCCM_OPS_ASYNC && (STAGE == 2 || STAGE == 3 || STAGE == 7); };
template <int STAGE> struct OccmWarpMap { static constexpr bool value = STAGE <= 2 || STAGE == 7; };
template <int STAGE> struct OccmNops { static constexpr int value = STAGE == 2 ? 2 : STAGE == 3 ? 3 : STAGE == 4 ? 4 : STAGE == 5 ? 5 : STAGE == 6 ? 5 : STAGE == 7 ? 2 : 0; };
#define OCCM_OPS_STRIDE (OCCM_FAST_BLOCK * 16)          // bytes between two operands of the ring (one 16-byte slot per thread)
template <int V>
__device__ __forceinline__ void occm_cp16(unsigned sa, const double* g) {
    if (V == 2) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sa), "l"(g) : "memory");
    else        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(sa), "l"(g) : "memory");
}
template <unsigned OFF, int V>
__device__ __forceinline__ double2 occm_lds2_at(unsigned sa) {
    double2 v;
    if (V == 2) asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(v.x), "=d"(v.y) : "r"(sa), "n"(OFF));
    else { asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v.x) : "r"(sa), "n"(OFF)); v.y = 0.0; }
    return v;
}

// sa = this thread's slot of operand 0 in the ring buffer that belongs to the trip
template <int STAGE, int V>
__device__ __forceinline__ void occm_fast_issue_ops(const OccmSysDev& sys, const OccmRkDev& rk, const OccmFastCtx<STAGE>& c, int e, unsigned sa) {
    const long long g = (long long)c.m * sys.ndof + e;
    occm_cp16<V>(sa, rk.Y[c.par] + g);
    if (STAGE != 7) occm_cp16<V>(sa + OCCM_OPS_STRIDE, rk.KF[c.par] + g);
    if (STAGE == 3) { occm_cp16<V>(sa + 2 * OCCM_OPS_STRIDE, rk.K2 + g); }
    if (STAGE == 4) { occm_cp16<V>(sa + 2 * OCCM_OPS_STRIDE, rk.K2 + g); occm_cp16<V>(sa + 3 * OCCM_OPS_STRIDE, rk.K3 + g); }
    if (STAGE == 5) { occm_cp16<V>(sa + 2 * OCCM_OPS_STRIDE, rk.K2 + g); occm_cp16<V>(sa + 3 * OCCM_OPS_STRIDE, rk.K3 + g); occm_cp16<V>(sa + 4 * OCCM_OPS_STRIDE, rk.K4 + g); }
    if (STAGE == 6) { occm_cp16<V>(sa + 2 * OCCM_OPS_STRIDE, rk.K3 + g); occm_cp16<V>(sa + 3 * OCCM_OPS_STRIDE, rk.K4 + g); occm_cp16<V>(sa + 4 * OCCM_OPS_STRIDE, rk.K5 + g); }
    if (STAGE == 7) { occm_cp16<V>(sa + OCCM_OPS_STRIDE, rk.K2 + g); }
}
template <int STAGE, unsigned RO, int V>     // RO = byte offset of the trip's ring buffer
__device__ __forceinline__ OccmOps2 occm_fast_read_ops(unsigned sa) {
    OccmOps2 o;
    const double2 z = make_double2(0.0, 0.0);
    o.ye = o.k1 = o.ka = o.kb = o.kc = o.kd = z;
    o.ye = occm_lds2_at<RO, V>(sa);
    if (STAGE != 7) o.k1 = occm_lds2_at<RO + OCCM_OPS_STRIDE, V>(sa);
    if (STAGE == 3) { o.ka = occm_lds2_at<RO + 2 * OCCM_OPS_STRIDE, V>(sa); }
    if (STAGE == 4) { o.ka = occm_lds2_at<RO + 2 * OCCM_OPS_STRIDE, V>(sa); o.kb = occm_lds2_at<RO + 3 * OCCM_OPS_STRIDE, V>(sa); }
    if (STAGE == 5 || STAGE == 6) { o.ka = occm_lds2_at<RO + 2 * OCCM_OPS_STRIDE, V>(sa); o.kb = occm_lds2_at<RO + 3 * OCCM_OPS_STRIDE, V>(sa); o.kc = occm_lds2_at<RO + 4 * OCCM_OPS_STRIDE, V>(sa); }
    if (STAGE == 7) { o.ka = occm_lds2_at<RO + OCCM_OPS_STRIDE, V>(sa); }
    return o;
}

// one species of the epilogue (same arithmetic as occm_finish); returns the squared scaled error in stage 7
template <int STAGE>
__device__ __forceinline__ double occm_fast_comb(const OccmRkDev& rk, double ye, double k1, double ka, double kb, double kc, double kd,
                                                 double f, double h, double yn, double& next) {
    if (STAGE == 2) next = ye + (dp::A31 * k1 + dp::A32 * f) * h;
    else if (STAGE == 3) next = ye + (dp::A41 * k1 + dp::A42 * ka + dp::A43 * f) * h;
    else if (STAGE == 4) next = ye + (dp::A51 * k1 + dp::A52 * ka + dp::A53 * kb + dp::A54 * f) * h;
    else if (STAGE == 5) next = ye + (dp::A61 * k1 + dp::A62 * ka + dp::A63 * kb + dp::A64 * kc + dp::A65 * f) * h;
    else if (STAGE == 6) next = ye + h * (dp::B1 * k1 + dp::B3 * ka + dp::B4 * kb + dp::B5 * kc + dp::B6 * f);
    else if (STAGE == 7) {
        const double err = (ka + dp::E7 * f) * h;          // ka = E1 k1 + E3 k3 + E4 k4 + E5 k5 + E6 k6 from stage 6
        const double scale = rk.atol + fmax(fabs(ye), fabs(yn)) * rk.rtol;
        const double q = err / scale;
        return q * q;
    }
    return 0.0;
}

template <int STAGE, int V>
__device__ __forceinline__ double occm_fast_finish(const OccmSysDev& sys, const OccmRkDev& rk, const OccmFastCtx<STAGE>& c, const OccmOps2& o,
                                                   double2 f, double2 yn, int e, double* __restrict__ dy_out) {
    const long long g = (long long)c.m * sys.ndof + e;
    if (STAGE <= 1) { occm_st2<V>(dy_out + g, f); return 0.0; }
    double2 nx = make_double2(0.0, 0.0);
    const double e0 = occm_fast_comb<STAGE>(rk, o.ye.x, o.k1.x, o.ka.x, o.kb.x, o.kc.x, o.kd.x, f.x, c.h, yn.x, nx.x);
    const double e1 = V == 2 ? occm_fast_comb<STAGE>(rk, o.ye.y, o.k1.y, o.ka.y, o.kb.y, o.kc.y, o.kd.y, f.y, c.h, yn.y, nx.y) : 0.0;
    if (STAGE == 2) { occm_st2<V>(rk.K2 + g, f); occm_st2<V>(rk.YS[0] + g, nx); }
    else if (STAGE == 3) { occm_st2<V>(rk.K3 + g, f); occm_st2<V>(rk.YS[1] + g, nx); }
    else if (STAGE == 4) { occm_st2<V>(rk.K4 + g, f); occm_st2<V>(rk.YS[0] + g, nx); }
    else if (STAGE == 5) { occm_st2<V>(rk.K5 + g, f); occm_st2<V>(rk.YS[1] + g, nx); }
    else if (STAGE == 6) {
        occm_st2<V>(rk.K6 + g, f); occm_st2<V>(rk.Y[c.par ^ 1] + g, nx);
        occm_st2<V>(rk.K2 + g, make_double2(dp::E1 * o.k1.x + dp::E3 * o.ka.x + dp::E4 * o.kb.x + dp::E5 * o.kc.x + dp::E6 * f.x,
                                          dp::E1 * o.k1.y + dp::E3 * o.ka.y + dp::E4 * o.kb.y + dp::E5 * o.kc.y + dp::E6 * f.y));
    }
    else { occm_st2<V>(rk.KF[c.par ^ 1] + g, f); }
    return e0 + e1;
}

// ---- shared memory through explicit 32-bit state-space addresses.  With generic pointers the compiler rebuilt the shared
// window base (S2UR CgaCtaId, ULEA, the table-size arithmetic ...) in front of every LDS of the reaction terms: ~10 extra
// issue slots per term (ncu source page, r01e).  An opaque 32-bit address costs one register and one add per access.
__device__ __forceinline__ unsigned occm_opaque_u32(unsigned v) { asm volatile("" : "+r"(v)); return v; }
__device__ __forceinline__ double occm_lds_f64(unsigned sa) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(sa)); return v; }
__device__ __forceinline__ unsigned occm_lds_u32(unsigned sa) { unsigned v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(sa)); return v; }
// the same with a compile-time byte offset folded into the instruction: the two tile buffers sit OCCM_TILE_BUF_BYTES apart, so
// the per-thread slot addresses are loop invariants (the compiler keeps as many of them in registers as it can afford)
template <unsigned OFF>
__device__ __forceinline__ double occm_lds_f64_at(unsigned sa) { double v; asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(sa), "n"(OFF)); return v; }
template <unsigned OFF, int V>
__device__ __forceinline__ void occm_sts2_at(unsigned sa, double2 v) {
    if (V == 2) asm volatile("st.shared.v2.f64 [%0+%3], {%1, %2};" :: "r"(sa), "d"(v.x), "d"(v.y), "n"(OFF) : "memory");
    else        asm volatile("st.shared.f64 [%0+%2], %1;" :: "r"(sa), "d"(v.x), "n"(OFF) : "memory");
}
#define OCCM_TILE_BUF_BYTES (OCCM_FAST_BLOCK * 32)      // >= tcv * (S + 2) * 8 for every even S >= 2

// Reaction rate of one species from its register-resident term list, straight-line: NTc terms x NFc factor slots are always
// evaluated (padding terms have coefficient 0, padding slots read the 1.0 column), which gives the same bits as stopping at
// the mechanism's own maxima and removes every uniform branch from the hot loop.  row_sa = shared address of the point's row.
template <int NTc, int NFc, unsigned BO>
__device__ __forceinline__ double occm_reaction_sa(const double (&cm)[OCCM_NT], const unsigned (&meta)[OCCM_NT], unsigned row_sa) {
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < NTc; ++i) {
        const unsigned m = meta[i];
        double v = cm[i];                      // explicit roundings: every variant (registers, shared lists, table walk) gives the same bits
        v = __dmul_rn(v, occm_lds_f64_at<BO>(row_sa + ((m >> 5) & 0x7f8u)));
        if (NFc > 1) v = __dmul_rn(v, occm_lds_f64_at<BO>(row_sa + ((m >> 13) & 0x7f8u)));
        if (NFc > 2) v = __dmul_rn(v, occm_lds_f64_at<BO>(row_sa + ((m >> 21) & 0x7f8u)));
        acc = __dadd_rn(acc, v);
    }
    return acc;
}
// term list in shared memory (WIDE): c_sa / mt_sa = shared addresses of the species' scaled coefficients and meta words
template <int NFc, unsigned BO>
__device__ __forceinline__ double occm_reaction_wide_sa(unsigned c_sa, unsigned mt_sa, unsigned row_sa, int nt) {
    double acc = 0.0;
    for (int i = 0; i < nt; ++i) {
        const unsigned m = occm_lds_u32(mt_sa + 4u * i);
        double v = occm_lds_f64(c_sa + 8u * i);
        v = __dmul_rn(v, occm_lds_f64_at<BO>(row_sa + ((m >> 5) & 0x7f8u)));
        if (NFc > 1) v = __dmul_rn(v, occm_lds_f64_at<BO>(row_sa + ((m >> 13) & 0x7f8u)));
        if (NFc > 2) v = __dmul_rn(v, occm_lds_f64_at<BO>(row_sa + ((m >> 21) & 0x7f8u)));
        acc = __dadd_rn(acc, v);
    }
    return acc;
}

template <int STAGE, int K>
struct OccmFastPre {          // everything of a trip that can be loaded before the tile barrier
    double2 own; int4 ell[K]; int pt; int ci;
};

template <int MODEL, int STAGE, int K, int WIDE, int V>
__device__ __forceinline__ void
occm_fast_body(const OccmSysDev& sys, const OccmMemDev& mem, const OccmRkDev& rk,
               const double* __restrict__ t_dev, double t_scalar, const double* __restrict__ y_in, double* __restrict__ dy_out,
               const OccmEllEntry* __restrict__ ell, int cpm, int tcv, int err_stride) {
    extern __shared__ double smem_d[];
    int* tab_s = reinterpret_cast<int*>(smem_d);
    occm_load_tables(sys.tables, sys.tables_bytes, tab_s);
    const int S = sys.S, H = S / V, RS = S + 2;            // staged rows: S values + two 1.0 pads (keeps 16-byte alignment)
    double* tile = smem_d + (((sys.tables_bytes + 15) >> 4) << 1);   // two buffers [tcv][RS], OCCM_TILE_BUF_BYTES apart, 16-byte aligned
    constexpr bool ASYNC = OccmOpsAsync<STAGE>::value;
    constexpr unsigned RING_BYTES = ASYNC ? OccmNops<STAGE>::value * OCCM_OPS_STRIDE : 0;      // one of the two operand buffers
    double* ops_ring = tile + 2 * (OCCM_TILE_BUF_BYTES / 8);
    double* rxc_s = ops_ring + 2 * (RING_BYTES / 8);                    // unscaled coefficients [S][OCCM_NT]  (WIDE: [S][nt_u])
    const int tid = threadIdx.x;
    // Warp-autonomous mapping: a warp owns 32 / H whole points (H = S / 2 threads per point), so the staged rows a thread
    // reads were written by its own warp and the per-trip barrier is a __syncwarp (the block barrier cost 15 % of the issue
    // slots in stalls).  The 32 % H left-over lanes of a warp idle (S = 10: 2 of 32).
    // Measured per stage (C4): the light stages gain (stage 2 -10 %, stage 7 -3 %, plain RHS -3 %), the operand-heavy stages
    // 3-6 lose the 6 % of idle lanes, so those keep the dense block mapping and the block barrier (OccmWarpMap).
    constexpr bool WARPMAP = OccmWarpMap<STAGE>::value && !WIDE;
    const int lane_w = tid & 31, pw = lane_w / H;
    const int ppw = 32 / H;                                   // points per warp and trip
    const bool lane_on = WARPMAP ? pw < ppw : tid < tcv * H;
    const int pl = WARPMAP ? (tid >> 5) * ppw + pw : tid / H;
    const int sp = WARPMAP ? V * (lane_w - pw * H) : V * (tid - pl * H);
    const int nt_u = tab_s[OCCM_H_MAX_TERMS], nf_u = tab_s[OCCM_H_MAX_SLOTS];
    // WIDE: term lists of every species in shared memory, coefficients scaled per member in a double-buffered copy
    const int NW = S * nt_u;
    double* cmw_s = rxc_s + (WIDE ? NW : S * OCCM_NT);      // [2][NW]
    unsigned* mtw_s = reinterpret_cast<unsigned*>(cmw_s + 2 * NW);
    int cmw_m0 = -1, cmw_m1 = -1;
    unsigned meta0[OCCM_NT], meta1[OCCM_NT];
    if (WIDE) {
        const OccmTab T = occm_tab_views(tab_s);
        for (int i = tid; i < NW; i += OCCM_FAST_BLOCK) {
            const int s_ = i / nt_u;
            double c; unsigned mt;
            occm_rxn_term(T, s_, i - s_ * nt_u, S, c, mt);
            rxc_s[i] = c; mtw_s[i] = mt;
        }
#pragma unroll
        for (int q = 0; q < OCCM_NT; ++q) { meta0[q] = 0u; meta1[q] = 0u; }
    } else {
        const OccmTab T = occm_tab_views(tab_s);
        const OccmRxnRegs R0 = occm_rxn_regs(T, lane_on ? sp : 0, S), R1 = occm_rxn_regs(T, (lane_on && V == 2) ? sp + 1 : 0, S);
#pragma unroll
        for (int q = 0; q < OCCM_NT; ++q) {
            meta0[q] = R0.meta[q]; meta1[q] = R1.meta[q];
            if (lane_on && pl == 0) { rxc_s[sp * OCCM_NT + q] = R0.coeff[q]; if (V == 2) rxc_s[(sp + 1) * OCCM_NT + q] = R1.coeff[q]; }   // warp 0, first point
        }
    }
    for (int r = tid; r < 2 * tcv; r += OCCM_FAST_BLOCK) {
        double* row = tile + (r >= tcv ? OCCM_TILE_BUF_BYTES / 8 + (r - tcv) * RS : r * RS);
        row[S] = 1.0; row[S + 1] = 1.0;
    }
    const int n_points = (int)sys.n_points, N = n_points;     // ELL columns are [K][n_points] (CSTR: one point per model)
    const unsigned row0_sa = occm_opaque_u32((unsigned)__cvta_generic_to_shared(tile) + 8u * (unsigned)(pl * RS));   // this thread's row, buffer 0
    const unsigned ops_sa = occm_opaque_u32((unsigned)__cvta_generic_to_shared(ops_ring) + 16u * (unsigned)tid);   // my slot of operand 0, buffer 0
    const unsigned cmw_sa = WIDE ? occm_opaque_u32((unsigned)__cvta_generic_to_shared(cmw_s)) : 0u;
    const unsigned mtw_sa = WIDE ? occm_opaque_u32((unsigned)__cvta_generic_to_shared(mtw_s)) : 0u;
    __syncthreads();

    // ---- cursor over this CTA's (member, chunk) trips, skipping members that are not running
    int m = (int)(blockIdx.x / (unsigned)cpm);
    int ci = (int)(blockIdx.x - (unsigned)m * (unsigned)cpm);
    const int step_m = (int)(gridDim.x / (unsigned)cpm), step_c = (int)(gridDim.x - (unsigned)step_m * (unsigned)cpm);
    auto skip = [&]() {       // move forward until (m, ci) is a trip that has to be done; false when past the end
        while (m < mem.M) {
            const bool run = STAGE >= 2 ? rk.status[m] == OCCM_MEMBER_RUNNING : (!mem.active || mem.active[m]);
            if (run) return true;
            m += step_m; ci += step_c; if (ci >= cpm) { ci -= cpm; ++m; }
        }
        return false;
    };
    auto advance = [&]() { m += step_m; ci += step_c; if (ci >= cpm) { ci -= cpm; ++m; } return skip(); };

    // Two register sets (A, B) alternate between "current trip" and "trip in flight": the loop body is instantiated
    // twice so that no register-to-register copies are needed.
    OccmFastCtx<STAGE> cxA, cxB;
    cxA.m = -1; cxB.m = -1;
    OccmFastPre<STAGE, K> A, B;
    int coeff_m = -1;
    double cm0[OCCM_NT], cm1[OCCM_NT];
#pragma unroll
    for (int q = 0; q < OCCM_NT; ++q) { cm0[q] = 0.0; cm1[q] = 0.0; }

    auto load = [&](OccmFastPre<STAGE, K>& pre, OccmFastCtx<STAGE>& cx, unsigned ring_off) {
        if (cx.m != m) cx = occm_fast_ctx<STAGE>(sys, mem, rk, m);
        pre.ci = ci;
        const int p = ci * tcv + pl;
        pre.pt = (lane_on && p < n_points) ? p : -1;
        const int pp = pre.pt >= 0 ? pre.pt : 0;
        const int e = pp * S + sp;
        pre.own = occm_fast_in<STAGE, V>(occm_fast_in_a<STAGE>(sys, rk, y_in, cx), occm_fast_in_b<STAGE>(sys, mem, rk, cx), cx.h, e);
#pragma unroll
        for (int k = 0; k < K; ++k) pre.ell[k] = __ldg(reinterpret_cast<const int4*>(ell + (size_t)k * N + pp));
        if (ASYNC) { if (pre.pt >= 0) occm_fast_issue_ops<STAGE, V>(sys, rk, cx, e, ops_sa + ring_off); }
        else if (pre.pt >= 0) occm_fast_prefetch_ops<STAGE>(sys, rk, cx, e);
    };

    auto process = [&](const OccmFastPre<STAGE, K>& cur, const OccmFastCtx<STAGE>& cc, OccmFastPre<STAGE, K>& nxt,
                       OccmFastCtx<STAGE>& ncx, bool have_next, auto buf_c) {
        constexpr int buf = decltype(buf_c)::value;
        constexpr unsigned BO = (unsigned)buf * OCCM_TILE_BUF_BYTES;      // byte offset of this trip's tile buffer
        // Plain RHS / J*v kernels have registers to spare: their slot addresses row0_sa + offset stay loop invariants (hoisted,
        // -12 % time).  The Dormand-Prince stages hold up to 20 operand registers: there the hoisting spilled (stage 5 +30 %),
        // so the row address is made opaque per trip and the slot offsets are added in the loop.
        const unsigned row_sa = STAGE <= 1 ? row0_sa : occm_opaque_u32(row0_sa);
        constexpr unsigned RO = (unsigned)buf * RING_BYTES;               // this trip's operand buffer; the next trip's is the other one
        // register version: epilogue operands of THIS trip are requested here, they have the whole trip to arrive
        OccmOps2 ops;
        if (!ASYNC) ops = occm_fast_load_ops<STAGE, V>(sys, rk, cc, (cur.pt >= 0 ? cur.pt : 0) * S + sp);
        if (lane_on) occm_sts2_at<BO, V>(row_sa + 8u * sp, cur.own);
        if (WIDE) {                                     // uniform: this buffer's coefficient copy belongs to another member
            int& have = buf ? cmw_m1 : cmw_m0;
            if (have != cc.m) {
                double* dst = cmw_s + buf * NW;
                for (int i = tid; i < NW; i += OCCM_FAST_BLOCK)
                    dst[i] = rxc_s[i] * (mem.k_scale ? __ldg(mem.k_scale + (long long)cc.m * sys.n_rxn + (mtw_s[i] & 0xffu)) : 1.0);
                have = cc.m;
            }
        }
        if (WARPMAP) __syncwarp();      // the staged rows of a warp's points are written and read by that warp only
        else __syncthreads();
        // ---- row gathers (L2 hits), then the next trip's loads: both fly while the reactions are evaluated
        double2 g[K];
        {
            const double* ga = occm_fast_in_a<STAGE>(sys, rk, y_in, cc);
            const double* gb = occm_fast_in_b<STAGE>(sys, mem, rk, cc);
#pragma unroll
            for (int k = 0; k < K; ++k) g[k] = occm_fast_in<STAGE, V>(ga, gb, cc.h, cur.ell[k].x * S + sp);
        }
        if (have_next) load(nxt, ncx, RING_BYTES - RO);
        if (ASYNC) asm volatile("cp.async.commit_group;" ::: "memory");     // exactly one group per trip (empty after the last one)

        if (!WIDE && coeff_m != cc.m) {                 // uniform: fold the member's rate-constant scales into the coefficients
#pragma unroll
            for (int q = 0; q < OCCM_NT; ++q) {
                const double k0 = mem.k_scale ? __ldg(mem.k_scale + (long long)cc.m * sys.n_rxn + (meta0[q] & 0xffu)) : 1.0;
                const double k1 = mem.k_scale ? __ldg(mem.k_scale + (long long)cc.m * sys.n_rxn + (meta1[q] & 0xffu)) : 1.0;
                cm0[q] = rxc_s[sp * OCCM_NT + q] * k0; cm1[q] = V == 2 ? rxc_s[(sp + 1) * OCCM_NT + q] * k1 : 0.0;
            }
            coeff_m = cc.m;
        }
        double err2 = 0.0;
        if (cur.pt >= 0) {
            if (!cur.ell[0].y) {
                double r0, r1;
                if (WIDE) {
                    const unsigned c0 = cmw_sa + 8u * (unsigned)(buf * NW + sp * nt_u), m0 = mtw_sa + 4u * (unsigned)(sp * nt_u);
                    if (nf_u <= 2) {
                        r0 = occm_reaction_wide_sa<2, BO>(c0, m0, row_sa, nt_u);
                        r1 = V == 2 ? occm_reaction_wide_sa<2, BO>(c0 + 8u * nt_u, m0 + 4u * nt_u, row_sa, nt_u) : 0.0;
                    } else {
                        r0 = occm_reaction_wide_sa<3, BO>(c0, m0, row_sa, nt_u);
                        r1 = V == 2 ? occm_reaction_wide_sa<3, BO>(c0 + 8u * nt_u, m0 + 4u * nt_u, row_sa, nt_u) : 0.0;
                    }
                } else if (nf_u <= 2) {                 // uniform: one of four straight-line variants
                    if (nt_u <= 2) { r0 = occm_reaction_sa<2, 2, BO>(cm0, meta0, row_sa); r1 = V == 2 ? occm_reaction_sa<2, 2, BO>(cm1, meta1, row_sa) : 0.0; }
                    else           { r0 = occm_reaction_sa<4, 2, BO>(cm0, meta0, row_sa); r1 = V == 2 ? occm_reaction_sa<4, 2, BO>(cm1, meta1, row_sa) : 0.0; }
                } else {
                    if (nt_u <= 2) { r0 = occm_reaction_sa<2, 3, BO>(cm0, meta0, row_sa); r1 = V == 2 ? occm_reaction_sa<2, 3, BO>(cm1, meta1, row_sa) : 0.0; }
                    else           { r0 = occm_reaction_sa<4, 3, BO>(cm0, meta0, row_sa); r1 = V == 2 ? occm_reaction_sa<4, 3, BO>(cm1, meta1, row_sa) : 0.0; }
                }
                double2 f;
                if (MODEL == 0) {
                    double2 acc = make_double2(0.0, 0.0);
#pragma unroll
                    for (int k = 0; k < K; ++k) acc = occm_fma2(__hiloint2double(cur.ell[k].w, cur.ell[k].z), g[k], acc);
                    f = make_double2(__dadd_rn(__dmul_rn(cc.qs, acc.x), r0), __dadd_rn(__dmul_rn(cc.qs, acc.y), r1));
                } else {                     // first-order upwind difference along the PFR
                    const double na = -(cc.qs * __hiloint2double(cur.ell[0].w, cur.ell[0].z));
                    f = make_double2(__dadd_rn(__dmul_rn(na, cur.own.x - g[0].x), r0), __dadd_rn(__dmul_rn(na, cur.own.y - g[0].y), r1));
                }
                if (ASYNC) {
                    asm volatile("cp.async.wait_group 1;" ::: "memory");    // everything but the next trip's group has landed
                    ops = occm_fast_read_ops<STAGE, RO, V>(ops_sa);
                }
                err2 = occm_fast_finish<STAGE, V>(sys, rk, cc, ops, f, cur.own, cur.pt * S + sp, dy_out);
            }   // flagged rows (boundary / pulse / overflow entries) are left to occm_fixup, launched right after
        }
        if (STAGE == 7) {          // one partial per warp: no block barrier in the hot loop (summed in fixed order by the controller)
            double v = err2;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
            if ((tid & 31) == 0) rk.err_partial[(long long)cc.m * err_stride + cur.ci * (OCCM_FAST_BLOCK / 32) + (tid >> 5)] = v;
        }
    };

    bool valid = skip();
    if (valid) { load(A, cxA, 0u); if (ASYNC) asm volatile("cp.async.commit_group;" ::: "memory"); }
    while (valid) {
        bool vn = advance();                 // (m, ci) now name the next trip
        process(A, cxA, B, cxB, vn, std::integral_constant<int, 0>{});
        if (!vn) break;
        valid = advance();
        process(B, cxB, A, cxA, valid, std::integral_constant<int, 1>{});
    }
}


template <int STAGE, int K, int WIDE, int V>
__global__ void __launch_bounds__(OCCM_FAST_BLOCK, OCCM_FAST_CTAS)
occm_fast_cstr(const __grid_constant__ OccmSysDev sys, const __grid_constant__ OccmMemDev mem, const __grid_constant__ OccmRkDev rk,
               const double* __restrict__ t_dev, double t_scalar, const double* __restrict__ y_in, double* __restrict__ dy_out,
               const OccmEllEntry* __restrict__ ell, int cpm, int tcv, int err_stride) {
    occm_fast_body<0, STAGE, K, WIDE, V>(sys, mem, rk, t_dev, t_scalar, y_in, dy_out, ell, cpm, tcv, err_stride);
}

template <int STAGE, int WIDE, int V>
__global__ void __launch_bounds__(OCCM_FAST_BLOCK, OCCM_FAST_CTAS)
occm_fast_pfr(const __grid_constant__ OccmSysDev sys, const __grid_constant__ OccmMemDev mem, const __grid_constant__ OccmRkDev rk,
              const double* __restrict__ t_dev, double t_scalar, const double* __restrict__ y_in, double* __restrict__ dy_out,
              const OccmEllEntry* __restrict__ ell, int cpm, int tcv, int err_stride) {
    occm_fast_body<1, STAGE, 1, WIDE, V>(sys, mem, rk, t_dev, t_scalar, y_in, dy_out, ell, cpm, tcv, err_stride);
}

// Flagged points of the fast path: rows with boundary / pulse entries or more entries than the ELL width, PFR inlet
// nodes.  Few points per member; evaluated with the generic CSR / table walk, one thread per (flagged point, species).
// PFR: each flagged point stages [upwind row | own row] so that interior nodes (pulse sources) find their neighbour.
template <int MODEL, int STAGE>
__global__ void __launch_bounds__(OCCM_BLOCK)
occm_fixup(const __grid_constant__ OccmSysDev sys, const __grid_constant__ OccmMemDev mem, const __grid_constant__ OccmRkDev rk,
           const double* __restrict__ t_dev, double t_scalar, const double* __restrict__ y_in, double* __restrict__ dy_out,
           const int* __restrict__ flagged, int n_flagged, int err_cpm, int err_offset) {
    extern __shared__ double smem_d[];
    __shared__ double scratch[8];
    int* tab_s = reinterpret_cast<int*>(smem_d);
    occm_load_tables(sys.tables, sys.tables_bytes, tab_s);
    const int S = sys.S, RS = S + 1, tcf = OCCM_BLOCK / S;
    double* tile = smem_d + (sys.tables_bytes >> 3);
    const int m = blockIdx.y;
    if (STAGE >= 2 && rk.status[m] != OCCM_MEMBER_RUNNING) return;
    if (STAGE <= 1 && mem.active && !mem.active[m]) return;
    const int tid = threadIdx.x, pl = tid / S, s = tid - pl * S;
    const int idx = blockIdx.x * tcf + pl;
    const bool on = pl < tcf && idx < n_flagged;
    const int p = on ? flagged[idx] : 0;
    const long long base = (long long)m * sys.ndof;
    double h = 0.0, t_stage; int par = 0;
    if (STAGE <= 1) { t_stage = t_dev ? __ldg(t_dev + m) : t_scalar; if (STAGE == 1) h = __ldg(mem.eps + m); }
    else {
        const double t0 = rk.t[m]; h = rk.h[m]; par = rk.parity[m];
        t_stage = STAGE == 2 ? t0 + dp::C2 * h : STAGE == 3 ? t0 + dp::C3 * h : STAGE == 4 ? t0 + dp::C4 * h : STAGE == 5 ? t0 + dp::C5 * h : t0 + h;
    }
    OccmStageIn<STAGE> in;
    in.a = STAGE <= 1 ? y_in + base : STAGE == 2 ? rk.Y[par] + base : STAGE == 3 || STAGE == 5 ? rk.YS[0] + base
         : STAGE == 4 || STAGE == 6 ? rk.YS[1] + base : rk.Y[par ^ 1] + base;
    in.b = STAGE == 1 ? mem.z + base : STAGE == 2 ? rk.KF[par] + base : nullptr; in.hb = h;
    const int row = MODEL == 1 ? 2 * pl + 1 : pl;
    if (on) {
        tile[row * RS + s] = in(p * S + s);
        if (MODEL == 1) tile[(row - 1) * RS + s] = p > 0 ? in((p - 1) * S + s) : 0.0;
    }
    __syncthreads();
    double err2 = 0.0;
    if (on) {
        const double* ksc = mem.k_scale ? mem.k_scale + (long long)m * sys.n_rxn : nullptr;
        err2 = occm_cold_element<MODEL, STAGE>(sys, mem, rk, tab_s, y_in, tile + row * RS, p, s, t_stage, m, ksc, dy_out);
    }
    if (STAGE == 7) {
        const double sum = occm_block_sum(err2, scratch);
        if (tid == 0) rk.err_partial[(long long)m * err_cpm + err_offset + blockIdx.x] = sum;
    }
}
