// TMA-pipelined form of the fused stage kernels (included by occm_solver.cu after occm_fast.cuh).
//
// Same arithmetic, same thread mapping and the same preconditions as the fast kernels of occm_fast.cuh (packed ELL operator,
// polynomial mechanism, flagged rows left to occm_fixup) — what changes is how the bytes reach the SM.  Every input of a
// trip is a CONTIGUOUS range of global memory: the chunk's rows of the stage state, of y, K1..K5 (species innermost, points
// consecutive) and of the K packed ELL columns.  A dedicated producer warp streams them into a ring of shared-memory slots
// with 1-D bulk copies (cp.async.bulk -> UBLKCP, completion on an mbarrier) several trips ahead; seven consumer warps wait
// on the slot's "full" barrier, compute, and hand the slot back through its "empty" barrier.  Nothing of a future trip is
// held in registers, so the bytes in flight per SM are set by the ring depth (tens of KB) instead of by the register file
// (the register-prefetching fast kernels had 8-16 KB in flight and sat at 25 % occupancy: latency bound).
//
// Neighbour rows whose source point lies inside the chunk (the diagonal entry, ring / upwind neighbours, any locality of
// the compartment numbering) are read from the staged rows in shared memory instead of being gathered from L2; in stage 2
// and in the J*v kernel both staged streams are combined on the fly, which removes the doubled gathers of those stages.
//
// Warps are autonomous: a warp owns 32 / H whole points of the chunk (H = threads per point), stages their stage state in
// its private padded tile for the reaction factors (__syncwarp only) and releases the slot with one mbarrier arrival.
#pragma once

#define OCCM_PIPE_WARPS 7                                   // consumer warps per CTA
#define OCCM_PIPE_THREADS (32 * (OCCM_PIPE_WARPS + 1))      // + one producer warp
#ifndef OCCM_PIPE_CTAS
#define OCCM_PIPE_CTAS 2
#endif
#ifndef OCCM_PIPE_CTAS_LIGHT
#define OCCM_PIPE_CTAS_LIGHT 2                              // stages with few streams (plain RHS, J*v, 2, 7): 3 CTAs (80 registers) spilled and measured slower
#endif
#define OCCM_PIPE_SUB_BYTES 4096u                           // a sub-chunk's rows of one vector: 7 * (32 / H) points * S * 8 B <= 4096
#define OCCM_PIPE_MAX_DEPTH 8
#ifndef OCCM_PIPE_L2_AHEAD
#define OCCM_PIPE_L2_AHEAD 0                                // experiment (measured slower, see below): NEXT member's rows of the gathered vector(s) -> L2
#endif
#ifndef OCCM_PIPE_E_LIGHT
#define OCCM_PIPE_E_LIGHT 1                                 // sub-chunks per trip in the light stages (see OccmPipeE)
#endif

// input streams of a stage: 0 = the stage state (stages 1 / 2: its two summands), then the epilogue operands
template <int STAGE> struct OccmPipeStreams { static constexpr int value = STAGE == 0 ? 1 : STAGE == 1 ? 2 : STAGE == 2 ? 2 : STAGE == 3 ? 4 : STAGE == 4 ? 5 : STAGE == 7 ? 3 : 6; };
template <int STAGE> struct OccmPipeLight { static constexpr bool value = STAGE <= 2 || STAGE == 7; };
// Sub-chunks per trip: a consumer thread works on E points of a trip (one per sub-chunk), issues the loads of all of them —
// among them the out-of-chunk row gathers, whose latency is the top stall of the light stages — and only then computes.
// Measured on C4 (512 members, profiles/r02_notes.md): E = 2 made the plain RHS 13 % and stage 2 34 % SLOWER (stage 7 4 % faster),
// E = 3 worse still — so E = 1 everywhere; the code path stays for the experiment (-DOCCM_PIPE_E_LIGHT=2).
template <int STAGE> struct OccmPipeE { static constexpr int value = OccmPipeLight<STAGE>::value ? OCCM_PIPE_E_LIGHT : 1; };
template <int STAGE> struct OccmPipeStreamBytes { static constexpr unsigned value = OccmPipeE<STAGE>::value * OCCM_PIPE_SUB_BYTES; };

struct OccmPipeGeom {
    int pc;            // points per chunk (E sub-chunks of psub points)
    int psub;          // points per sub-chunk = consumer warps * (32 / H)
    int cpm;           // chunks per member
    int depth;         // ring slots
    int slot_bytes;    // bytes of one slot: streams, then the ELL columns
    int ell_stride;    // bytes between two ELL columns of a slot (pc * 16)
    int err_stride;    // error partials per member (stage 7)
    int tile_off;      // byte offsets into dynamic shared memory
    int coef_off;
    int bar_off;
    int ring_off;
    int wide_stride;   // WIDE: bytes between the per-warp scaled coefficient copies
};

// ---- mbarrier / bulk-copy primitives (PTX ISA 8.x, sm_90+)
__device__ __forceinline__ void occm_mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void occm_mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void occm_mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void occm_mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" :: "r"(bar), "r"(parity) : "memory");
}
// global -> shared, completion counted in bytes on `bar`; dst, src and bytes are multiples of 16
// contiguous range -> L2 only (no shared memory, no completion): src and bytes are multiples of 16
__device__ __forceinline__ void occm_bulk_prefetch_l2(const void* src, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void occm_bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// base pointer (member 0) of input stream j of a stage
template <int STAGE>
__device__ __forceinline__ const double* occm_pipe_src(const OccmMemDev& mem, const OccmRkDev& rk, const double* y_in, int par, int j) {
    if (STAGE == 0) return y_in;
    if (STAGE == 1) return j == 0 ? y_in : mem.z;
    if (STAGE == 2) return j == 0 ? rk.Y[par] : rk.KF[par];
    if (j == 0) return STAGE == 3 || STAGE == 5 ? rk.YS[0] : STAGE == 4 || STAGE == 6 ? rk.YS[1] : rk.Y[par ^ 1];
    if (j == 1) return rk.Y[par];
    if (STAGE == 7) return rk.K2;
    if (j == 2) return rk.KF[par];
    if (STAGE == 6) return j == 3 ? rk.K3 : j == 4 ? rk.K4 : rk.K5;
    return j == 3 ? rk.K2 : j == 4 ? rk.K3 : rk.K4;
}

template <int V>
__device__ __forceinline__ double2 occm_lds2(unsigned sa) {
    double2 v;
    if (V == 2) asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(sa));
    else { asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v.x) : "r"(sa)); v.y = 0.0; }
    return v;
}

// stage state at byte offset `off` of the slot's stream 0 (stages 1 / 2: combined from both staged summands with the same
// roundings as occm_fast_in)
template <int STAGE, int V>
__device__ __forceinline__ double2 occm_pipe_state(unsigned slot_sa, unsigned off, double hb) {
    constexpr unsigned SB = OccmPipeStreamBytes<STAGE>::value;
    const double2 a = occm_lds2<V>(slot_sa + off);
    if (STAGE == 2) { const double2 b = occm_lds2<V>(slot_sa + SB + off); return make_double2(a.x + (dp::A21 * b.x) * hb, a.y + (dp::A21 * b.y) * hb); }
    if (STAGE == 1) { const double2 b = occm_lds2<V>(slot_sa + SB + off); return make_double2(a.x + b.x * hb, a.y + b.y * hb); }
    return a;
}

template <int STAGE, int V>
__device__ __forceinline__ OccmOps2 occm_pipe_read_ops(unsigned sa) {      // sa = slot + this thread's element offset
    OccmOps2 o;
    const double2 z = make_double2(0.0, 0.0);
    o.ye = o.k1 = o.ka = o.kb = o.kc = o.kd = z;
    constexpr unsigned SB = OccmPipeStreamBytes<STAGE>::value;
    if (STAGE <= 1) return o;
    if (STAGE == 2) { o.ye = occm_lds2_at<0u, V>(sa); o.k1 = occm_lds2_at<SB, V>(sa); return o; }
    o.ye = occm_lds2_at<SB, V>(sa);
    if (STAGE == 7) { o.ka = occm_lds2_at<2 * SB, V>(sa); return o; }
    o.k1 = occm_lds2_at<2 * SB, V>(sa);
    o.ka = occm_lds2_at<3 * SB, V>(sa);
    if (STAGE >= 4) o.kb = occm_lds2_at<4 * SB, V>(sa);
    if (STAGE >= 5) o.kc = occm_lds2_at<5 * SB, V>(sa);
    return o;
}

template <int MODEL, int STAGE, int K, int WIDE, int V>
__device__ __forceinline__ void
occm_pipe_body(const OccmSysDev& sys, const OccmMemDev& mem, const OccmRkDev& rk, const double* __restrict__ y_in,
               double* __restrict__ dy_out, const OccmEllEntry* __restrict__ ell, const OccmPipeGeom& g) {
    extern __shared__ __align__(128) unsigned char smem_p[];
    int* tab_s = reinterpret_cast<int*>(smem_p);
    occm_load_tables(sys.tables, sys.tables_bytes, tab_s);               // ends with __syncthreads (all 9 warps)
    constexpr int NS = OccmPipeStreams<STAGE>::value;
    constexpr int E = OccmPipeE<STAGE>::value;
    constexpr unsigned SB = OccmPipeStreamBytes<STAGE>::value;
    const int S = sys.S, H = S / V, RS = S + 2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ppw = 32 / H;                                                // points per warp and trip
    const int n_points = (int)sys.n_points;
    double* rxc_s = reinterpret_cast<double*>(smem_p + g.coef_off);      // unscaled coefficients [S][OCCM_NT]  (WIDE: [S][nt_u])
    double* tile = reinterpret_cast<double*>(smem_p + g.tile_off);       // [E][OCCM_PIPE_WARPS][ppw][RS]
    const unsigned smem_sa = (unsigned)__cvta_generic_to_shared(smem_p);
    const unsigned bar_sa = smem_sa + (unsigned)g.bar_off;                // full[depth], empty[depth]
    const unsigned ring_sa = smem_sa + (unsigned)g.ring_off;
    const int nt_u = tab_s[OCCM_H_MAX_TERMS], nf_u = tab_s[OCCM_H_MAX_SLOTS];
    const int NW = S * nt_u;

    // ---- one-time setup by all threads: barriers, coefficient tables, the 1.0 pad columns of the tiles
    if (tid == 0) {
        for (int i = 0; i < g.depth; ++i) { occm_mbar_init(bar_sa + 8u * i, 1u); occm_mbar_init(bar_sa + 8u * (g.depth + i), OCCM_PIPE_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    unsigned* mtw_s = reinterpret_cast<unsigned*>(rxc_s + NW);           // WIDE: meta words [S][nt_u], then the per-warp scaled copies
    {
        const OccmTab T = occm_tab_views(tab_s);
        if (WIDE) {
            for (int i = tid; i < NW; i += OCCM_PIPE_THREADS) {
                const int s_ = i / nt_u;
                double c; unsigned mt;
                occm_rxn_term(T, s_, i - s_ * nt_u, S, c, mt);
                rxc_s[i] = c; mtw_s[i] = mt;
            }
        } else {
            for (int i = tid; i < S * OCCM_NT; i += OCCM_PIPE_THREADS) {
                double c; unsigned mt;
                occm_rxn_term(T, i / OCCM_NT, i % OCCM_NT, S, c, mt);
                rxc_s[i] = c;
            }
        }
    }
    for (int r = tid; r < E * OCCM_PIPE_WARPS * ppw; r += OCCM_PIPE_THREADS) { tile[r * RS + S] = 1.0; tile[r * RS + S + 1] = 1.0; }
    __syncthreads();

    // ---- cursor over this CTA's (member, chunk) trips, skipping members that are not running (identical in every warp)
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

    if (warp == OCCM_PIPE_WARPS) {
        // ================================================================ producer: one thread streams the trips into the ring
        // (Measured and dropped: the idle lanes of this warp pulling the out-of-chunk neighbour rows into L2 ahead of the
        //  consumers.  prefetch.global.L2 fetches whole lines for rows of which the gathers touch 2-3 sectors, and it runs ahead
        //  of the streaming pass that would have brought half of them in anyway: DRAM traffic up, every kernel 10-90 % slower.)
        if (lane != 0) return;
        int cur_m = -1, par = 0, ahead_m = -1, ahead_par = 0;
        for (bool valid = skip(); valid; valid = advance()) {
            if (cur_m != m) { cur_m = m; par = STAGE >= 2 ? rk.parity[m] : 0; }
            const int p0 = ci * g.pc;
            const int npts = min(g.pc, n_points - p0);
            const unsigned vbytes = (unsigned)(npts * S) * 8u, ebytes = (unsigned)npts * 16u;
            const unsigned full = bar_sa + 8u * slot, empty = bar_sa + 8u * (g.depth + slot);
            const unsigned dst = ring_sa + (unsigned)slot * (unsigned)g.slot_bytes;
            occm_mbar_wait(empty, phase ^ 1u);                             // every consumer warp has released the slot
            occm_mbar_expect_tx(full, NS * vbytes + K * ebytes);
            const long long e0 = (long long)m * sys.ndof + (long long)p0 * S;
#pragma unroll
            for (int j = 0; j < NS; ++j)
                occm_bulk_g2s(dst + j * SB, occm_pipe_src<STAGE>(mem, rk, y_in, par, j) + e0, vbytes, full);
#pragma unroll
            for (int k = 0; k < K; ++k)
                occm_bulk_g2s(dst + NS * SB + (unsigned)(k * g.ell_stride), ell + (size_t)k * n_points + p0, ebytes, full);
            // ---- experiment, off by default: one member ahead, the same chunk of the NEXT member's stage-state vector(s) goes to
            //      L2 (cp.async.bulk.prefetch.L2), so that the out-of-chunk row gathers — half of which point ahead of the streaming
            //      front and miss L2 — would hit, with every line fetched once and whole.  Measured (C4, 512 members): every kernel
            //      20-45 % slower (plain RHS 3.19 -> 3.96 ms, stage 5 5.56 -> 7.00 ms): the prefetched lines do not survive the
            //      65 us until their member is streamed and are read from DRAM twice.
            if (OCCM_PIPE_L2_AHEAD && m + 1 < mem.M) {
                if (ahead_m != m + 1) { ahead_m = m + 1; ahead_par = STAGE >= 2 ? rk.parity[m + 1] : 0; }
                const long long e1 = (long long)(m + 1) * sys.ndof + (long long)p0 * S;
                occm_bulk_prefetch_l2(occm_pipe_src<STAGE>(mem, rk, y_in, ahead_par, 0) + e1, vbytes);
                if (STAGE == 1 || STAGE == 2) occm_bulk_prefetch_l2(occm_pipe_src<STAGE>(mem, rk, y_in, ahead_par, 1) + e1, vbytes);
            }
            if (++slot == g.depth) { slot = 0; phase ^= 1u; }
        }
        return;
    }

    // ==================================================================== consumers
    const int pw = lane / H;
    const bool lane_on = pw < ppw;
    const int pl = warp * ppw + pw;                                        // point of a sub-chunk this thread works on
    const int sp = V * (lane - pw * H);                                    // its first species
    const unsigned eo = 8u * (unsigned)(pl * S + sp);                      // byte offset of its element inside a staged sub-chunk
    const unsigned sub_bytes = 8u * (unsigned)(g.psub * S);                // bytes between the sub-chunks of a staged stream
    const unsigned row_sa = occm_opaque_u32((unsigned)__cvta_generic_to_shared(tile) + 8u * (unsigned)((warp * ppw + pw) * RS));
    const unsigned tile_sub = 8u * (unsigned)(OCCM_PIPE_WARPS * ppw * RS); // bytes between the tiles of two sub-chunks
    const unsigned cmw_sa = WIDE ? occm_opaque_u32((unsigned)__cvta_generic_to_shared(mtw_s) + 4u * (unsigned)(NW + (NW & 1)) + (unsigned)(warp * g.wide_stride)) : 0u;
    const unsigned mtw_sa = WIDE ? occm_opaque_u32((unsigned)__cvta_generic_to_shared(mtw_s)) : 0u;
    unsigned meta0[OCCM_NT], meta1[OCCM_NT];
    double cm0[OCCM_NT], cm1[OCCM_NT];
    {
        const OccmTab T = occm_tab_views(tab_s);
        const OccmRxnRegs R0 = occm_rxn_regs(T, (!WIDE && lane_on) ? sp : 0, S), R1 = occm_rxn_regs(T, (!WIDE && lane_on && V == 2) ? sp + 1 : 0, S);
#pragma unroll
        for (int q = 0; q < OCCM_NT; ++q) { meta0[q] = R0.meta[q]; meta1[q] = R1.meta[q]; cm0[q] = 0.0; cm1[q] = 0.0; }
    }
    OccmFastCtx<STAGE> cc;
    cc.m = -1; cc.h = 0.0; cc.qs = 1.0; cc.par = 0;

    for (bool valid = skip(); valid; valid = advance()) {
        if (cc.m != m) {                         // uniform: new member -> step, flow scale, rate-constant scales folded into the coefficients
            cc = occm_fast_ctx<STAGE>(sys, mem, rk, m);
            if (WIDE) {
                __syncwarp();
                for (int i = lane; i < NW; i += 32) {
                    const double sc = mem.k_scale ? __ldg(mem.k_scale + (long long)m * sys.n_rxn + (mtw_s[i] & 0xffu)) : 1.0;
                    asm volatile("st.shared.f64 [%0], %1;" :: "r"(cmw_sa + 8u * i), "d"(rxc_s[i] * sc) : "memory");
                }
                __syncwarp();
            } else {
#pragma unroll
                for (int q = 0; q < OCCM_NT; ++q) {
                    const double k0 = mem.k_scale ? __ldg(mem.k_scale + (long long)m * sys.n_rxn + (meta0[q] & 0xffu)) : 1.0;
                    const double k1 = mem.k_scale ? __ldg(mem.k_scale + (long long)m * sys.n_rxn + (meta1[q] & 0xffu)) : 1.0;
                    cm0[q] = rxc_s[sp * OCCM_NT + q] * k0; cm1[q] = V == 2 ? rxc_s[(sp + 1) * OCCM_NT + q] * k1 : 0.0;
                }
            }
        }
        const int p0 = ci * g.pc;
        const int npts = min(g.pc, n_points - p0);
        const unsigned slot_sa = ring_sa + (unsigned)slot * (unsigned)g.slot_bytes;
        const unsigned ell_sa = slot_sa + NS * SB;
        const double* ga = occm_fast_in_a<STAGE>(sys, rk, y_in, cc);
        const double* gb = occm_fast_in_b<STAGE>(sys, mem, rk, cc);
        occm_mbar_wait(bar_sa + 8u * slot, phase);                         // the trip's bytes have landed
        // ---- phase 1, all E points of this thread: own stage state into the warp's tile, ELL sources, neighbour rows — from the
        //      staged chunk when the source point is inside it, else a gather (L2); every gather of the trip is in flight together
        int pt[E];
        double2 own[E], gv[E][K];
        unsigned skipmask = 0;                   // bit i: point i is flagged (left to occm_fixup) or past the end
#pragma unroll
        for (int i = 0; i < E; ++i) {
            const int pli = i * g.psub + pl;
            pt[i] = (lane_on && p0 + pli < n_points) ? p0 + pli : -1;
        }
        // (the gathers first: their latency then also covers the tile store and the warp barrier)
#pragma unroll
        for (int i = 0; i < E; ++i) {
#pragma unroll
            for (int k = 0; k < K; ++k) gv[i][k] = make_double2(0.0, 0.0);
            if (pt[i] < 0) { skipmask |= 1u << i; continue; }
            const unsigned esa = ell_sa + 16u * (unsigned)(i * g.psub + pl);
            int src0, flag;
            asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(src0), "=r"(flag) : "r"(esa));
            if (flag) { skipmask |= 1u << i; continue; }
#pragma unroll
            for (int k = 0; k < K; ++k) {
                int src = src0;
                if (k > 0) asm volatile("ld.shared.s32 %0, [%1];" : "=r"(src) : "r"(esa + (unsigned)(k * g.ell_stride)));
                const unsigned rel = (unsigned)(src - p0);
                if (rel < (unsigned)npts) gv[i][k] = occm_pipe_state<STAGE, V>(slot_sa, 8u * (rel * (unsigned)S + (unsigned)sp), cc.h);
                else gv[i][k] = occm_fast_in<STAGE, V>(ga, gb, cc.h, src * S + sp);
            }
        }
#pragma unroll
        for (int i = 0; i < E; ++i) {
            own[i] = make_double2(0.0, 0.0);
            if (pt[i] >= 0) {
                own[i] = occm_pipe_state<STAGE, V>(slot_sa, eo + i * sub_bytes, cc.h);
                occm_sts2_at<0u, V>(row_sa + i * tile_sub + 8u * sp, own[i]);
            }
        }
        __syncwarp();                            // the staged rows of a warp's points are written and read by that warp only
        // ---- phase 2: reactions (tile), transport (weights re-read from the staged ELL columns), epilogue
        double err2 = 0.0;
#pragma unroll
        for (int i = 0; i < E; ++i) {
            if (skipmask >> i & 1u) continue;
            const unsigned rsa = row_sa + i * tile_sub;
            double r0, r1;
            if (WIDE) {
                const unsigned c0 = cmw_sa + 8u * (unsigned)(sp * nt_u), m0 = mtw_sa + 4u * (unsigned)(sp * nt_u);
                if (nf_u <= 2) {
                    r0 = occm_reaction_wide_sa<2, 0u>(c0, m0, rsa, nt_u);
                    r1 = V == 2 ? occm_reaction_wide_sa<2, 0u>(c0 + 8u * nt_u, m0 + 4u * nt_u, rsa, nt_u) : 0.0;
                } else {
                    r0 = occm_reaction_wide_sa<3, 0u>(c0, m0, rsa, nt_u);
                    r1 = V == 2 ? occm_reaction_wide_sa<3, 0u>(c0 + 8u * nt_u, m0 + 4u * nt_u, rsa, nt_u) : 0.0;
                }
            } else if (nf_u <= 2) {                 // uniform: one of four straight-line variants
                if (nt_u <= 2) { r0 = occm_reaction_sa<2, 2, 0u>(cm0, meta0, rsa); r1 = V == 2 ? occm_reaction_sa<2, 2, 0u>(cm1, meta1, rsa) : 0.0; }
                else           { r0 = occm_reaction_sa<4, 2, 0u>(cm0, meta0, rsa); r1 = V == 2 ? occm_reaction_sa<4, 2, 0u>(cm1, meta1, rsa) : 0.0; }
            } else {
                if (nt_u <= 2) { r0 = occm_reaction_sa<2, 3, 0u>(cm0, meta0, rsa); r1 = V == 2 ? occm_reaction_sa<2, 3, 0u>(cm1, meta1, rsa) : 0.0; }
                else           { r0 = occm_reaction_sa<4, 3, 0u>(cm0, meta0, rsa); r1 = V == 2 ? occm_reaction_sa<4, 3, 0u>(cm1, meta1, rsa) : 0.0; }
            }
            const unsigned esa = ell_sa + 16u * (unsigned)(i * g.psub + pl) + 8u;       // the weights of this point's ELL entries
            double2 f;
            if (MODEL == 0) {
                double2 acc = make_double2(0.0, 0.0);
#pragma unroll
                for (int k = 0; k < K; ++k) acc = occm_fma2(occm_lds_f64(esa + (unsigned)(k * g.ell_stride)), gv[i][k], acc);
                f = make_double2(__dadd_rn(__dmul_rn(cc.qs, acc.x), r0), __dadd_rn(__dmul_rn(cc.qs, acc.y), r1));
            } else {                     // first-order upwind difference along the PFR
                const double na = -(cc.qs * occm_lds_f64(esa));
                f = make_double2(__dadd_rn(__dmul_rn(na, own[i].x - gv[i][0].x), r0), __dadd_rn(__dmul_rn(na, own[i].y - gv[i][0].y), r1));
            }
            const OccmOps2 ops = occm_pipe_read_ops<STAGE, V>(slot_sa + eo + i * sub_bytes);
            err2 += occm_fast_finish<STAGE, V>(sys, rk, cc, ops, f, own[i], pt[i] * S + sp, dy_out);
        }   // flagged rows (boundary / pulse / overflow entries, PFR inlet nodes) are left to occm_fixup, launched right after
        if (STAGE == 7) {          // one partial per warp, summed in fixed order by the controller
            double v = err2;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
            if (lane == 0) rk.err_partial[(long long)cc.m * g.err_stride + ci * OCCM_PIPE_WARPS + warp] = v;
        }
        __syncwarp();                            // every lane is done with the slot and with the warp's tile
        if (lane == 0) occm_mbar_arrive(bar_sa + 8u * (g.depth + slot));
        if (++slot == g.depth) { slot = 0; phase ^= 1u; }
    }
}

template <int MODEL, int STAGE, int K, int WIDE, int V>
__global__ void __launch_bounds__(OCCM_PIPE_THREADS, (OccmPipeLight<STAGE>::value ? OCCM_PIPE_CTAS_LIGHT : OCCM_PIPE_CTAS))
occm_pipe(const __grid_constant__ OccmSysDev sys, const __grid_constant__ OccmMemDev mem, const __grid_constant__ OccmRkDev rk,
          const double* __restrict__ y_in, double* __restrict__ dy_out, const OccmEllEntry* __restrict__ ell,
          const __grid_constant__ OccmPipeGeom g) {
    occm_pipe_body<MODEL, STAGE, K, WIDE, V>(sys, mem, rk, y_in, dy_out, ell, g);
}
