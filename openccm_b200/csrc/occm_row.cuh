// Thread-per-POINT form of the light stage kernels (plain RHS, J*v state, stages 2 and 7) of CSTR and PFR networks, compiled per mechanism
// (spec/occm_spec.cu includes this file after the generated mechanism header; see occm_spec.h).
//
// Why: in occm_pipe.cuh a thread owns a species PAIR, so the species a reaction term multiplies differ from lane to lane and are
// fetched from shared memory per lane — 1.3 shared-memory wavefronts per element: the plain RHS runs the L1 / shared-memory data
// pipe at 90 % of its peak and stops at 0.40 of the HBM roofline (profiles/r02_notes.md).  Here a thread owns a whole point: its
// S concentrations live in registers and the mechanism is straight-line code over them (OCCM_SPEC_TERMS: the term list with the
// species indices as compile-time constants), so reactions cost no memory instruction at all.  A table-driven walk over
// registers was measured first (warp-uniform switches per factor): 2.4x slower than occm_pipe — a register file cannot be
// indexed; the indices have to be constants.
//
// Data movement: the ring of occm_pipe.cuh.  A producer thread streams the chunk's rows (32 * W points) and ELL
// columns into shared memory with bulk copies (mbarrier full / empty pairs).  A neighbour row is read with generic 128-bit loads
// from a pointer that is selected per lane — the staged chunk when the neighbour is inside it, global memory otherwise — so the
// K * S / 2 loads of a point carry no branch and are all in flight together.  Results are staged per warp in shared memory and
// leave with one bulk store per output vector (32 rows are contiguous in global memory): no 16-byte scattered STG.
// A persistent CTA strides through the chunks by the grid size, so nearly every trip belongs to another member: the next trip's
// member status, step, parity, flow and rate scales are requested one trip ahead (consumers) / before the wait for the slot
// (producer), which takes two dependent L2 round trips off every trip.
// Measured and rejected (C4, 512 members): L2 prefetch of the next trip's out-of-chunk rows (plain RHS 1.62 -> 1.84 ms); branching
// between shared-memory and global loads instead of generic loads (+2 %); a warp-cooperative gather of the out-of-chunk rows
// (S / 2 lanes per row, redistributed through shared memory: fewer L1 wavefronts, but the loads of a column then wait on each
// other: plain RHS 1.66 -> 2.31 ms); 128-bit stores from registers instead of the staged
// bulk store (+16 %: a lane's row is 80 contiguous bytes, a warp store instruction scatters over 32 sectors); the largest
// shared-memory carve-out (28 KB of L1 left: 2x slower — the five 16-byte loads of a gathered row hit the line the first one brought).
// Where the remaining gap is (C4, 512 members): with every neighbour taken from the staged chunk (wrong results, timing only) the
// plain RHS runs at 0.97 of the HBM peak, stage 2 at 0.93, stage 7 at 0.79 — the one far (out-of-chunk) neighbour row per point
// costs 0.34 / 1.45 / 0.63 ms.  The ELL columns are ordered nearest source first, so the far sources share the last column; copying
// the next trip's far rows into per-lane staging rows with cp.async while the current trip computes was 50 % SLOWER (plain RHS
// 1.70 -> 2.55 ms, with or without waiting for the next slot's ELL columns), like every other form of prefetch tried here.
// Same arithmetic and roundings as the other kernel forms (bit-identical outputs); flagged rows are left to occm_fixup.
#pragma once
#include "occm_spec.h"

#ifndef OCCM_ROW_BULK_STORE
#define OCCM_ROW_BULK_STORE 1          // 0: 128-bit stores straight from registers (A/B)
#endif

template <int STAGE> struct OccmRowOut { static constexpr int value = STAGE == 2 ? 2 : 1; };      // output vectors of a stage

__device__ __forceinline__ void occm_bulk_s2g(void* dst, unsigned src_sa, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst), "r"(src_sa), "r"(bytes) : "memory");
}
__device__ __forceinline__ void occm_sts2(unsigned sa, double2 v) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" :: "r"(sa), "d"(v.x), "d"(v.y) : "memory");
}

// two adjacent values of the stage state at byte offset `off` of the slot's stream 0 (stage 2 combines both staged summands
// with the roundings of occm_fast_in)
template <int STAGE>
__device__ __forceinline__ double2 occm_row_state(unsigned slot_sa, unsigned off, unsigned stream_bytes, double hb) {
    const double2 a = occm_lds2<2>(slot_sa + off);
    if (STAGE == 2) { const double2 b = occm_lds2<2>(slot_sa + stream_bytes + off); return make_double2(a.x + (dp::A21 * b.x) * hb, a.y + (dp::A21 * b.y) * hb); }
    if (STAGE == 1) { const double2 b = occm_lds2<2>(slot_sa + stream_bytes + off); return make_double2(a.x + b.x * hb, a.y + b.y * hb); }
    return a;
}

// a row of the stage state through generic pointers (shared or global window)
template <int STAGE, int S>
__device__ __forceinline__ void occm_row_load(double (&c)[S], const double* pa, const double* pb, double hb) {
    double2 a[S / 2], b[S / 2];
#pragma unroll
    for (int j = 0; j < S / 2; ++j) a[j] = reinterpret_cast<const double2*>(pa)[j];
    if (STAGE == 1 || STAGE == 2) {
#pragma unroll
        for (int j = 0; j < S / 2; ++j) b[j] = reinterpret_cast<const double2*>(pb)[j];
    }
#pragma unroll
    for (int j = 0; j < S / 2; ++j) {
        if (STAGE == 2) { c[2 * j] = a[j].x + (dp::A21 * b[j].x) * hb; c[2 * j + 1] = a[j].y + (dp::A21 * b[j].y) * hb; }
        else if (STAGE == 1) { c[2 * j] = a[j].x + b[j].x * hb; c[2 * j + 1] = a[j].y + b[j].y * hb; }
        else { c[2 * j] = a[j].x; c[2 * j + 1] = a[j].y; }
    }
}

template <int STAGE, int W>                       // W consumer warps (32 points each) + one producer warp
__device__ __forceinline__ void occm_row_body(const OccmSysDev& sys, const OccmMemDev& mem, const OccmRkDev& rk,
                                              const double* __restrict__ y_in, double* __restrict__ dy_out,
                                              const OccmEllEntry* __restrict__ ell, const OccmRowGeom& g) {
    constexpr int S = OCCM_SPEC_S, K = OCCM_SPEC_K, NT = OCCM_SPEC_NTERM;
    static_assert(S % 2 == 0 && S <= OCCM_SPEC_MAX_S, "even species counts up to 16");
    static_assert(STAGE == 0 || STAGE == 1 || STAGE == 2 || STAGE == 7, "light stages");
    constexpr int MODEL = OCCM_SPEC_MODEL;           // 0: CSTR rows (K ELL entries), 1: PFR (one upwind entry per point)
    static_assert(MODEL == 0 || K == 1, "a PFR point has one upwind neighbour");
    extern __shared__ __align__(128) unsigned char smem_p[];
    constexpr int NS = OccmPipeStreams<STAGE>::value, NOUT = OccmRowOut<STAGE>::value;
    constexpr int POINTS = 32 * W;
    constexpr unsigned ROW_BYTES = S * 8u, WARP_BYTES = 32u * ROW_BYTES, ELL_COL = POINTS * 16u;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_points = (int)sys.n_points;
    const unsigned smem_sa = (unsigned)__cvta_generic_to_shared(smem_p);
    const unsigned bar_sa = smem_sa + (unsigned)g.bar_off, ring_sa = smem_sa + (unsigned)g.ring_off;
    if (tid == 0) {
        for (int i = 0; i < g.depth; ++i) { occm_mbar_init(bar_sa + 8u * i, 1u); occm_mbar_init(bar_sa + 8u * (g.depth + i), W); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int cpm = g.cpm;
    int m = (int)(blockIdx.x / (unsigned)cpm);
    int ci = (int)(blockIdx.x - (unsigned)m * (unsigned)cpm);
    const int step_m = (int)(gridDim.x / (unsigned)cpm), step_c = (int)(gridDim.x - (unsigned)step_m * (unsigned)cpm);
    auto skip = [&]() {
        while (m < mem.M) {
            const bool run = STAGE >= 2 ? rk.status[m] == OCCM_MEMBER_RUNNING : (!mem.active || mem.active[m]);
            if (run) return true;
            m += step_m; ci += step_c; if (ci >= cpm) { ci -= cpm; ++m; }
        }
        return false;
    };
    auto advance = [&]() { m += step_m; ci += step_c; if (ci >= cpm) { ci -= cpm; ++m; } return skip(); };
    int slot = 0; unsigned phase = 0;
    const unsigned stream_bytes = (unsigned)g.stream_bytes;

    if (warp == W) {
        // ================================================================ producer: one thread streams the trips into the ring
        if (lane != 0) return;
        // the position (member status) and parity of a trip are looked up right after the previous trip's copies were issued,
        // i.e. before the wait for its slot: their latency is off the critical path of a shallow ring
        int cur_m = -1, par = 0;
        bool valid = skip();
        if (valid) { cur_m = m; par = STAGE >= 2 ? rk.parity[m] : 0; }
        while (valid) {
            const int p0 = ci * POINTS;
            const int npts = min(POINTS, n_points - p0);
            const unsigned vbytes = (unsigned)(npts * S) * 8u, ebytes = (unsigned)npts * 16u;
            const unsigned full = bar_sa + 8u * slot, empty = bar_sa + 8u * (g.depth + slot);
            const unsigned dst = ring_sa + (unsigned)slot * (unsigned)g.slot_bytes;
            const long long e0 = (long long)m * sys.ndof + (long long)p0 * S;
            const int par_now = par;
            valid = advance();
            if (valid && cur_m != m) { cur_m = m; par = STAGE >= 2 ? rk.parity[m] : 0; }
            occm_mbar_wait(empty, phase ^ 1u);
            occm_mbar_expect_tx(full, NS * vbytes + K * ebytes);
#pragma unroll
            for (int j = 0; j < NS; ++j)
                occm_bulk_g2s(dst + j * stream_bytes, occm_pipe_src<STAGE>(mem, rk, y_in, par_now, j) + e0, vbytes, full);
#pragma unroll
            for (int k = 0; k < K; ++k)
                occm_bulk_g2s(dst + NS * stream_bytes + k * ELL_COL, ell + (size_t)k * n_points + p0, ebytes, full);
            if (++slot == g.depth) { slot = 0; phase ^= 1u; }
        }
        return;
    }

    // ==================================================================== consumers: one point per thread
    const int pl = warp * 32 + lane;
    double* my_scaled = reinterpret_cast<double*>(smem_p + g.coef_off) + warp * NT;
    const unsigned out_sa = smem_sa + (unsigned)g.out_off + (unsigned)warp * (NOUT * WARP_BYTES);
    const unsigned own_off = (unsigned)pl * ROW_BYTES;
    OccmFastCtx<STAGE> cc, nx;
    cc.m = -1; cc.h = 0.0; cc.qs = 1.0; cc.par = 0;
    nx = cc;
    constexpr int NTL = (NT + 31) / 32;          // coefficients a lane scales per member
    double nx_ks[NTL];
    bool nx_run = false;
#pragma unroll
    for (int j = 0; j < NTL; ++j) nx_ks[j] = 1.0;
    for (bool valid = skip(); valid; ) {
        if (cc.m != m) {                         // uniform: new member -> step, flow scale, scaled coefficients of this warp
            __syncwarp();
            if (nx.m == m) {                     // looked up one trip ahead (below)
                cc = nx;
#pragma unroll
                for (int j = 0; j < NTL; ++j) { const int i = lane + 32 * j; if (i < NT) my_scaled[i] = occm_spec_coef[i] * nx_ks[j]; }
            } else {
                cc = occm_fast_ctx<STAGE>(sys, mem, rk, m);
                for (int i = lane; i < NT; i += 32)
                    my_scaled[i] = occm_spec_coef[i] * (mem.k_scale ? __ldg(mem.k_scale + (long long)m * sys.n_rxn + occm_spec_rxn[i]) : 1.0);
            }
            __syncwarp();
        }
        // One trip ahead: a persistent CTA strides through the chunks by the grid size, so nearly every trip belongs to another
        // member; its status, context and rate scales are requested now and consumed at the next trip, off the critical path.
        int m2 = m + step_m, ci2 = ci + step_c;
        if (ci2 >= cpm) { ci2 -= cpm; ++m2; }
        nx.m = -1;
        if (m2 != m && m2 < mem.M) {
            nx_run = STAGE >= 2 ? rk.status[m2] == OCCM_MEMBER_RUNNING : (!mem.active || mem.active[m2]);
            nx = occm_fast_ctx<STAGE>(sys, mem, rk, m2);
#pragma unroll
            for (int j = 0; j < NTL; ++j) {
                const int i = lane + 32 * j;
                nx_ks[j] = (i < NT && mem.k_scale) ? __ldg(mem.k_scale + (long long)m2 * sys.n_rxn + occm_spec_rxn[i]) : 1.0;
            }
        }
        const int p0 = ci * POINTS;
        const int npts = min(POINTS, n_points - p0);
        const int p = p0 + pl;
        const unsigned slot_off = (unsigned)slot * (unsigned)g.slot_bytes;
        const unsigned slot_sa = ring_sa + slot_off;
        const unsigned char* slot_gp = smem_p + g.ring_off + slot_off;           // the same slot as a generic pointer
        const long long mbase = (long long)m * sys.ndof;
        occm_mbar_wait(bar_sa + 8u * slot, phase);
        double err2 = 0.0;
        int4 e4[K];
        bool work = p < n_points;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (work)
                asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(e4[k].x), "=r"(e4[k].y), "=r"(e4[k].z), "=r"(e4[k].w)
                             : "r"(slot_sa + NS * stream_bytes + k * ELL_COL + 16u * (unsigned)pl));
            else e4[k] = make_int4(n_points - 1, 1, 0, 0);           // past the end: a valid row to point at, weight 0, flagged
        }
        work = !e4[0].y;
#if OCCM_ROW_BULK_STORE
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the previous trip's stores have left the staging rows
        __syncwarp();
#endif
        double f[S];
#pragma unroll
        for (int s = 0; s < S; ++s) f[s] = 0.0;
        double c[S];                                   // own stage state (stage 7: y_new)
#pragma unroll
        for (int s = 0; s < S; ++s) c[s] = 0.0;
        {
            const double* ga = occm_fast_in_a<STAGE>(sys, rk, y_in, cc);
            const double* gb = occm_fast_in_b<STAGE>(sys, mem, rk, cc);
            // ---- transport: acc = sum_k w_k * row(src_k)
            double acc[S];
#pragma unroll
            for (int s = 0; s < S; ++s) acc[s] = 0.0;
            const double* pa[K]; const double* pb[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const unsigned rel = (unsigned)(e4[k].x - p0);
                const bool in = rel < (unsigned)npts;
                pa[k] = in ? reinterpret_cast<const double*>(slot_gp + rel * ROW_BYTES) : ga + (long long)e4[k].x * S;
                pb[k] = in ? reinterpret_cast<const double*>(slot_gp + stream_bytes + rel * ROW_BYTES) : gb + (long long)e4[k].x * S;
            }
#pragma unroll
            for (int k = 0; k < K; ++k) {
                double nb[S];
                occm_row_load<STAGE, S>(nb, pa[k], pb[k], cc.h);
                const double w = __hiloint2double(e4[k].w, e4[k].z);
#pragma unroll
                for (int s = 0; s < S; ++s) acc[s] = MODEL == 0 ? fma(w, nb[s], acc[s]) : nb[s];      // PFR: keeps the upwind row
            }
            // ---- reactions: straight-line code over the registers of the point (generated per mechanism)
#pragma unroll
            for (int j = 0; j < S / 2; ++j) {
                const double2 v = occm_row_state<STAGE>(slot_sa, own_off + 16u * j, stream_bytes, cc.h);
                c[2 * j] = v.x; c[2 * j + 1] = v.y;
            }
            double r[S];
#pragma unroll
            for (int s = 0; s < S; ++s) r[s] = 0.0;
            OCCM_SPEC_TERMS(c, my_scaled, r)
#pragma unroll
            for (int s = 0; s < S; ++s) {
                if (MODEL == 0) f[s] = __dadd_rn(__dmul_rn(cc.qs, acc[s]), r[s]);
                else {                   // first-order upwind difference along the PFR (pfr_system.py:296), roundings of occm_pipe
                    const double na = -(cc.qs * __hiloint2double(e4[0].w, e4[0].z));
                    f[s] = __dadd_rn(__dmul_rn(na, c[s] - acc[s]), r[s]);
                }
                if (!work) f[s] = 0.0;                   // flagged rows and the tail past the last point: rewritten by occm_fixup / not stored
            }
        }
        // ---- epilogue of the stage; rows of flagged points are rewritten by occm_fixup afterwards (zeros here)
        double2 o0[S / 2], o1[S / 2];
#pragma unroll
        for (int j = 0; j < S / 2; ++j) {
            o0[j] = make_double2(f[2 * j], f[2 * j + 1]);
            o1[j] = make_double2(0.0, 0.0);
            if (STAGE == 2 && work) {
                const double2 ye = occm_lds2<2>(slot_sa + own_off + 16u * j), k1 = occm_lds2<2>(slot_sa + stream_bytes + own_off + 16u * j);
                o1[j] = make_double2(ye.x + (dp::A31 * k1.x + dp::A32 * f[2 * j]) * cc.h, ye.y + (dp::A31 * k1.y + dp::A32 * f[2 * j + 1]) * cc.h);
            }
            if (STAGE == 7 && work) {
                const double2 ye = occm_lds2<2>(slot_sa + stream_bytes + own_off + 16u * j), ka = occm_lds2<2>(slot_sa + 2 * stream_bytes + own_off + 16u * j);
                const double ex = (ka.x + dp::E7 * f[2 * j]) * cc.h, ey = (ka.y + dp::E7 * f[2 * j + 1]) * cc.h;
                const double sx = rk.atol + fmax(fabs(ye.x), fabs(c[2 * j])) * rk.rtol, sy = rk.atol + fmax(fabs(ye.y), fabs(c[2 * j + 1])) * rk.rtol;
                const double qx = ex / sx, qy = ey / sy;
                err2 += qx * qx + qy * qy;
            }
        }
        // the slot is not read any more: hand it back before the stores
        __syncwarp();
        if (lane == 0) occm_mbar_arrive(bar_sa + 8u * (g.depth + slot));
        if (++slot == g.depth) { slot = 0; phase ^= 1u; }

        double* out0 = (STAGE <= 1 ? dy_out : STAGE == 2 ? rk.K2 : rk.KF[cc.par ^ 1]) + mbase;
        double* out1 = STAGE == 2 ? rk.YS[0] + mbase : nullptr;
#if OCCM_ROW_BULK_STORE
        const int w0 = p0 + warp * 32;                                  // first point of this warp
        const int wn = min(32, n_points - w0);                          // its valid points (<= 0: nothing to store)
#pragma unroll
        for (int j = 0; j < S / 2; ++j) {
            occm_sts2(out_sa + (unsigned)lane * ROW_BYTES + 16u * j, o0[j]);
            if (NOUT == 2) occm_sts2(out_sa + WARP_BYTES + (unsigned)lane * ROW_BYTES + 16u * j, o1[j]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0 && wn > 0) {
            occm_bulk_s2g(out0 + (long long)w0 * S, out_sa, (unsigned)wn * ROW_BYTES);
            if (NOUT == 2) occm_bulk_s2g(out1 + (long long)w0 * S, out_sa + WARP_BYTES, (unsigned)wn * ROW_BYTES);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
#else
        if (work) {
            const long long ge = (long long)p * S;
#pragma unroll
            for (int j = 0; j < S / 2; ++j) {
                reinterpret_cast<double2*>(out0 + ge)[j] = o0[j];
                if (NOUT == 2) reinterpret_cast<double2*>(out1 + ge)[j] = o1[j];
            }
        }
#endif
        if (STAGE == 7) {
            double v = err2;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
            if (lane == 0) rk.err_partial[(long long)cc.m * g.err_stride + ci * W + warp] = v;
        }
        // advance() with the looked-up status: same sequence of trips as the producer's
        const int m_prev = m;
        m = m2; ci = ci2;
        if (m == m_prev) valid = true;
        else if (m >= mem.M) valid = false;
        else if (nx_run) valid = true;
        else { nx.m = -1; valid = skip(); }
    }
#if OCCM_ROW_BULK_STORE
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
#endif
}
