// Persistent RK45 for networks that fit in shared memory (included by the main translation unit of occm_solver.cu).
//
// The batched kernels of occm_pipe.cuh / occm_fast.cuh are built for 1e6-DOF members: a step attempt is eight dependent
// launches, which on a 300- or 900-DOF network (BASELINE configs C1-C3: one CSTR, the 76-CSTR and the 444-PFR networks of the
// OpenFOAM example) is 83 us of pure launch-to-launch latency — the 134 861-step RK45 solve of the 444-PFR network took 11.2 s.
// Here ONE CTA owns ONE ensemble member for the whole integration: state, the seven Dormand-Prince slopes, the stage state
// and the step controller live in shared memory; stages, error norm, accept / reject, step-size control and the dense
// output at the t_eval points inside an accepted step all happen between block barriers, with no launch in between.
// Same arithmetic as the batched path: the stage combinations of occm_fast_comb, the generic CSR / table-walk right-hand side
// (occm_point_generic: every row, boundary and pulse entries included), the controller of occm_rk45_control and the dense
// output of occm_rk45_output (scipy rk.py:14-71, 111-167, 575-601).
// The kernel starts from and returns to the per-member workspace state (Y[parity], KF[parity] = K1, t, h, counters), so
// occm_rk45_stats / get_state / a later run() see no difference.
#pragma once

#define OCCM_SMALL_MAX_NDOF 2560      // (element plan: 16-bit element indices)

struct OccmSmallIn {        // stage state of this member in shared memory, indexed by the flat element like the global one
    const double* a;
    __device__ __forceinline__ double operator()(int e) const { return a[e]; }
};

struct OccmSmallCtl {       // per-member controller state (shared memory copy of the OccmRkDev scalars)
    double t, t_old, t_new, h, h_abs, err_sum;
    long long n_acc, n_rej, nfev;
    int status, rejected, out_lo, out_hi, accepted;
};

// Element plan of a member's CTA (built once per launch): the elements ordered by how their slope is evaluated, so that warps
// are homogeneous — a warp that mixed classes would execute every path.  Class A / B are transport-only elements (species without
// reaction terms, models without pulse sources — every element of a tracer / RTD run): a handful of shared-memory loads and FMAs,
// written to give the bits of the generic path.  Boundary entries of their rows take the boundary value g_b,s(t_stage) from a
// small table that is filled once per step for all six stage times (the byte-code evaluation of a boundary function was the
// critical path of every stage: one warp walked it while fifteen waited at the barrier).  Everything else calls
// occm_point_generic (ncu, profiles/r02m_small_ncu_raw.csv: 70 % of the instructions of a step sat there, ~350 per element).
#ifndef OCCM_SMALL_PLAN
#define OCCM_SMALL_PLAN 1            // 0: every element through occm_point_generic (A/B, debugging)
#endif
#define OCCM_SMALL_BC_SLOTS 128      // boundary groups x species the per-step table holds (per stage); beyond: generic path
struct OccmSmallPlan {
    const unsigned short* el;     // [ndof] flat element index, class A first, then B, then the generic rest
    const unsigned short* aux;    // [ndof] PFR: model index k of the element; CSTR: its point
    int n_a, n_b;                 // PFR: A = interior nodes, B = inlet nodes; CSTR: A = rows, no B
    const int* row_ptr; const int* row_src; const double* row_w; const double* a;      // the shared-memory operator
    const double* bcv;            // [6][n_bc * S] boundary values at the stage times of this step
    int n_bcs;                    // n_bc * S
};

// f = ddt(t_stage, ys) for every element of the member (block-wide), followed by `post(e, f)` on the same element — the
// combination that only needs the element's own values (next stage state, y_new, error term) rides in the same pass, so a stage
// is ONE pass and ONE barrier.  `post` must not write anything another thread reads during this pass (the stage states ping-pong).
template <int MODEL, class Post>
__device__ __forceinline__ void occm_small_stage(const OccmSysDev& sys, const int* tab_s, const OccmSmallPlan& pl, int stage, const double* ys,
                                                 double* f, double t_stage, double qs, double bs, const double* ksc, const Post& post) {
    const int S = sys.S, ndof = (int)sys.ndof, P = sys.P;
    const int n_a = pl.n_a, n_ab = pl.n_a + pl.n_b;
    const double* bcv = pl.bcv + stage * pl.n_bcs;
    int i = threadIdx.x;
    for (; i < n_a; i += blockDim.x) {
        const int e = pl.el[i];
        double v;
        if (MODEL == 1) {                          // interior node: upwind difference (occm_point_generic, node > 0; pulses = reactions = 0)
            const double conv = __dmul_rn(-(qs * pl.a[pl.aux[i]]), ys[e] - ys[e - S]);
            v = __dadd_rn(__dadd_rn(conv, 0.0), 0.0);
        } else {                                   // CSTR row
            const int p = pl.aux[i], s = e - p * S;
            double acc = 0.0;
            for (int k = pl.row_ptr[p]; k < pl.row_ptr[p + 1]; ++k) {
                const int src = pl.row_src[k];
                const double w = pl.row_w[k];
                if (src >= 0) acc += w * ys[src * S + s];
                else          acc += w * (bs * bcv[(-1 - src) * S + s]);
            }
            v = __dadd_rn(__dadd_rn(__dmul_rn(qs, acc), 0.0), 0.0);
        }
        f[e] = v;
        post(e, v);
    }
    if (MODEL == 1) {
        for (; i < n_ab; i += blockDim.x) {        // inlet node: weighted outlet slopes of the upstream PFRs, boundary inflow (pfr_system.py:313-314)
            const int e = pl.el[i];
            const int kk = pl.aux[i], s = e - kk * P * S;
            double acc = 0.0;
            for (int q = pl.row_ptr[kk]; q < pl.row_ptr[kk + 1]; ++q) {
                const int src = pl.row_src[q];
                const double w = pl.row_w[q];
                if (src >= 0) {
                    const int o = P * (src + 1) - 1;
                    double d = -(qs * pl.a[src]) * (ys[o * S + s] - ys[(o - 1) * S + s]);
                    d += 0.0;
                    d += 0.0;
                    acc += w * d;
                } else {
                    acc += w * (bs * bcv[(-1 - src) * S + s]);
                }
            }
            f[e] = acc;
            post(e, acc);
        }
    }
    if (n_ab < ndof) {                             // uniform: reacting species, pulse sources
        OccmSmallIn in; in.a = ys;
        for (; i < ndof; i += blockDim.x) {
            const int e = pl.el[i];
            const int p = e / S, s = e - p * S;
            const double* myrow = ys + p * S;
            const double v = occm_point_generic<MODEL, OccmSmallIn, true>(sys, tab_s, in, myrow, myrow - S, p, s, t_stage, qs, bs, ksc);
            f[e] = v;
            post(e, v);
        }
    }
    __syncthreads();
}

template <int MODEL>
__global__ void __launch_bounds__(512, 1)
occm_rk45_small(const __grid_constant__ OccmSysDev sys_g, const __grid_constant__ OccmMemDev mem, const __grid_constant__ OccmRkDev rk,
                const double* __restrict__ t_eval, int n_t_eval, const int* __restrict__ save_idx, int n_save,
                double* __restrict__ out, double* __restrict__ out_t, long long max_attempts, int nnz, int pulse_nnz) {
    extern __shared__ double smem_d[];
    __shared__ OccmSmallCtl ctl;
    __shared__ double scratch[32];
    const int m = blockIdx.x;
    if (rk.status[m] != OCCM_MEMBER_RUNNING) return;
    int* tab_s = reinterpret_cast<int*>(smem_d);
    occm_load_tables(sys_g.tables, sys_g.tables_bytes, tab_s);
    const int ndof = (int)sys_g.ndof, tid = threadIdx.x, nthr = blockDim.x;
    double* y = smem_d + (sys_g.tables_bytes >> 3);
    double* yn = y + ndof;            // y_new (stage-7 state); swaps with y on an accepted step
    double* ysA = yn + ndof;          // stage states ping-pong: a stage reads one while its pass writes the next into the other
    double* ysB = ysA + ndof;
    double* Kb = ysB + ndof;          // K1..K7: slot i at Kb + i * ndof; K1 and K7 swap roles on an accepted step (FSAL)
    // ---- the operator itself (CSR rows, Q/dV per PFR, pulse lists) also lives in shared memory: the generic row walk is a
    //      chain of dependent index / weight loads, 300+ cycles each from L2 (29 us per step on the 444-PFR network before)
    __shared__ OccmSysDev sys_sh;                                    // (a local copy would live in local memory: it is passed by reference)
    __shared__ int plan_count[4];
    __shared__ double bcv_s[6 * OCCM_SMALL_BC_SLOTS];
    const OccmSysDev& sys = sys_sh;
    unsigned short* plan_el; unsigned short* plan_aux;
    const int* op_row_ptr; const int* op_row_src; const double* op_row_w; const double* op_a; const int* op_pulse_ptr;
    {
        const int nm = sys_g.n_models;
        double* wq = Kb + 7 * ndof;                                  // doubles first, then the int arrays
        double* row_w_s = wq; wq += nnz;
        double* a_s = wq; wq += (MODEL == 1 ? nm : 0);
        double* pulse_w_s = wq; wq += pulse_nnz;
        int* iq = reinterpret_cast<int*>(wq);
        int* row_ptr_s = iq; iq += nm + 1;
        int* row_src_s = iq; iq += nnz;
        int* pulse_ptr_s = iq; iq += (sys_g.pulse_ptr ? nm + 1 : 0);
        int* pulse_fn_s = iq; iq += pulse_nnz;
        plan_el = reinterpret_cast<unsigned short*>(iq);
        plan_aux = plan_el + ndof;
        op_row_ptr = row_ptr_s; op_row_src = row_src_s; op_row_w = row_w_s; op_a = a_s;
        op_pulse_ptr = sys_g.pulse_ptr ? pulse_ptr_s : nullptr;
        for (int i = tid; i < nnz; i += nthr) { row_w_s[i] = sys_g.row_w[i]; row_src_s[i] = sys_g.row_src[i]; }
        for (int i = tid; i <= nm; i += nthr) { row_ptr_s[i] = sys_g.row_ptr[i]; if (sys_g.pulse_ptr) pulse_ptr_s[i] = sys_g.pulse_ptr[i]; }
        if (MODEL == 1) for (int i = tid; i < nm; i += nthr) a_s[i] = sys_g.a[i];
        for (int i = tid; i < pulse_nnz; i += nthr) { pulse_w_s[i] = sys_g.pulse_w[i]; pulse_fn_s[i] = sys_g.pulse_fn[i]; }
        if (tid == 0) {
            sys_sh = sys_g;
            sys_sh.row_ptr = row_ptr_s; sys_sh.row_src = row_src_s; sys_sh.row_w = row_w_s;
            if (MODEL == 1) sys_sh.a = a_s;
            if (sys_g.pulse_ptr) { sys_sh.pulse_ptr = pulse_ptr_s; sys_sh.pulse_fn = pulse_fn_s; sys_sh.pulse_w = pulse_w_s; }
        }
    }
    int i1 = 0, i7 = 6;
    const long long base = (long long)m * ndof;
    const double qs = mem.q_scale ? mem.q_scale[m] : 1.0, bs = mem.bc_scale ? mem.bc_scale[m] : 1.0;
    const double* ksc = mem.k_scale ? mem.k_scale + (long long)m * sys_g.n_rxn : nullptr;     // (sys_sh is not complete before the barrier below)
    {
        const int par = rk.parity[m];
        for (int e = tid; e < ndof; e += nthr) { y[e] = rk.Y[par][base + e]; Kb[e] = rk.KF[par][base + e]; }
        if (tid == 0) {
            ctl.t = rk.t[m]; ctl.t_old = rk.t_old[m]; ctl.t_new = rk.t_new[m]; ctl.h = rk.h[m]; ctl.h_abs = rk.h_abs[m];
            ctl.n_acc = rk.n_acc[m]; ctl.n_rej = rk.n_rej[m]; ctl.nfev = rk.nfev[m];
            ctl.status = OCCM_MEMBER_RUNNING; ctl.rejected = rk.rejected[m]; ctl.out_lo = rk.out_lo[m]; ctl.out_hi = rk.out_hi[m];
            ctl.accepted = 0;
        }
    }
    if (tid < 4) plan_count[tid] = 0;
    __syncthreads();
    // ---- element plan: class of every element (0 = A, 1 = B, 2 = generic), counted, then placed class by class
    OccmSmallPlan plan;
    {
        const OccmTab T = occm_tab_views(tab_s);
        const int S = sys_g.S, P = sys_g.P;
        auto no_pulse = [&](int model) { return !op_pulse_ptr || op_pulse_ptr[model + 1] == op_pulse_ptr[model]; };
        const bool bc_table = T.n_bc * S <= OCCM_SMALL_BC_SLOTS;
        auto classify = [&](int e, int& aux) {
            const int p = e / S, s = e - p * S;
            aux = MODEL == 1 ? p / P : p;
            if (!OCCM_SMALL_PLAN || T.term_ptr[s + 1] > T.term_ptr[s]) return 2;   // the species reacts
            if (MODEL == 1) {
                const int k = aux;
                if (p - k * P > 0) return no_pulse(k) ? 0 : 2;
                for (int q = op_row_ptr[k]; q < op_row_ptr[k + 1]; ++q) {
                    const int src = op_row_src[q];
                    if (src < 0 ? !bc_table : !no_pulse(src)) return 2;            // boundary values past the table / upstream PFR with a pulse source
                }
                return 1;
            }
            if (!no_pulse(p)) return 2;
            for (int q = op_row_ptr[p]; q < op_row_ptr[p + 1]; ++q)
                if (op_row_src[q] < 0 && !bc_table) return 2;
            return 0;
        };
        for (int e = tid; e < ndof; e += nthr) { int aux; atomicAdd(&plan_count[classify(e, aux)], 1); }
        __syncthreads();
        const int n_a = plan_count[0], n_b = plan_count[1];
        __syncthreads();
        if (tid == 0) { plan_count[0] = 0; plan_count[1] = n_a; plan_count[2] = n_a + n_b; }
        __syncthreads();
        for (int e = tid; e < ndof; e += nthr) {
            int aux;
            const int c = classify(e, aux);
            const int at = atomicAdd(&plan_count[c], 1);
            plan_el[at] = (unsigned short)e; plan_aux[at] = (unsigned short)aux;
        }
        plan.el = plan_el; plan.aux = plan_aux; plan.n_a = n_a; plan.n_b = n_b;
        plan.row_ptr = op_row_ptr; plan.row_src = op_row_src; plan.row_w = op_row_w; plan.a = op_a;
        plan.bcv = bcv_s; plan.n_bcs = bc_table ? T.n_bc * S : 0;
    }
    __syncthreads();

    for (long long attempt = 0; attempt < max_attempts && ctl.status == OCCM_MEMBER_RUNNING; ++attempt) {
        const double t = ctl.t, h = ctl.h;
        double* K1 = Kb + i1 * ndof; double* K2 = Kb + 1 * ndof; double* K3 = Kb + 2 * ndof; double* K4 = Kb + 3 * ndof;
        double* K5 = Kb + 4 * ndof; double* K6 = Kb + 5 * ndof; double* K7 = Kb + i7 * ndof;
        // ---- stages 2..7 (rk.py:14-71, 64-69, 157-162): each pass evaluates a slope and forms what the next pass differentiates
        for (int e = tid; e < ndof; e += nthr) ysA[e] = y[e] + (dp::A21 * K1[e]) * h;
        for (int q = tid; q < 6 * plan.n_bcs; q += nthr) {       // boundary values at the six stage times (the t expressions of the stage calls below)
            const OccmTab T = occm_tab_views(tab_s);
            const int st = q / plan.n_bcs, j = q - st * plan.n_bcs;
            const double ts = st == 0 ? t + dp::C2 * h : st == 1 ? t + dp::C3 * h : st == 2 ? t + dp::C4 * h : st == 3 ? t + dp::C5 * h : t + h;
            bcv_s[st * plan.n_bcs + j] = occm_bc_value(T, j / sys_g.S, j - (j / sys_g.S) * sys_g.S, ts);
        }
        __syncthreads();
        occm_small_stage<MODEL>(sys, tab_s, plan, 0, ysA, K2, t + dp::C2 * h, qs, bs, ksc,
                                [&](int e, double f) { ysB[e] = y[e] + (dp::A31 * K1[e] + dp::A32 * f) * h; });
        occm_small_stage<MODEL>(sys, tab_s, plan, 1, ysB, K3, t + dp::C3 * h, qs, bs, ksc,
                                [&](int e, double f) { ysA[e] = y[e] + (dp::A41 * K1[e] + dp::A42 * K2[e] + dp::A43 * f) * h; });
        occm_small_stage<MODEL>(sys, tab_s, plan, 2, ysA, K4, t + dp::C4 * h, qs, bs, ksc,
                                [&](int e, double f) { ysB[e] = y[e] + (dp::A51 * K1[e] + dp::A52 * K2[e] + dp::A53 * K3[e] + dp::A54 * f) * h; });
        occm_small_stage<MODEL>(sys, tab_s, plan, 3, ysB, K5, t + dp::C5 * h, qs, bs, ksc,
                                [&](int e, double f) { ysA[e] = y[e] + (dp::A61 * K1[e] + dp::A62 * K2[e] + dp::A63 * K3[e] + dp::A64 * K4[e] + dp::A65 * f) * h; });
        occm_small_stage<MODEL>(sys, tab_s, plan, 4, ysA, K6, t + h, qs, bs, ksc,
                                [&](int e, double f) { yn[e] = y[e] + h * (dp::B1 * K1[e] + dp::B3 * K3[e] + dp::B4 * K4[e] + dp::B5 * K5[e] + dp::B6 * f); });
        double e2 = 0.0;
        occm_small_stage<MODEL>(sys, tab_s, plan, 5, yn, K7, t + h, qs, bs, ksc, [&](int e, double f) {
            const double part = dp::E1 * K1[e] + dp::E3 * K3[e] + dp::E4 * K4[e] + dp::E5 * K5[e] + dp::E6 * K6[e];
            const double err = (part + dp::E7 * f) * h;
            const double scale = rk.atol + fmax(fabs(y[e]), fabs(yn[e])) * rk.rtol;
            const double q = err / scale;
            ysA[e] = q * q;                                   // (ysA is free here) summed below in element order: the norm — and with it
        });                                                   // every step-size decision — does not depend on the element plan
        for (int e = tid; e < ndof; e += nthr) e2 += ysA[e];
        const double sum = occm_block_sum(e2, scratch);       // valid in thread 0; contains the barriers
        // ---- controller (rk.py:111-167; same statements as occm_rk45_control)
        if (tid == 0) {
            const double error_norm = sqrt(sum / (double)ndof);
            double tt = ctl.t, h_abs = ctl.h_abs;
            ctl.nfev += 6;
            ctl.accepted = 0;
            if (error_norm < 1.0) {
                double factor = error_norm == 0.0 ? dp::MAX_FACTOR : fmin(dp::MAX_FACTOR, dp::SAFETY * pow(error_norm, dp::ERROR_EXPONENT));
                if (ctl.rejected) factor = fmin(1.0, factor);
                h_abs *= factor;
                ctl.t_old = tt;
                tt = ctl.t_new;
                ctl.t = tt;
                ctl.rejected = 0;
                ctl.accepted = 1;
                const long long n_acc = ++ctl.n_acc;
                int status = OCCM_MEMBER_RUNNING;
                if (tt - rk.t_bound >= 0.0) status = OCCM_MEMBER_FINISHED;
                else if (rk.max_steps > 0 && n_acc + ctl.n_rej >= rk.max_steps) status = OCCM_MEMBER_MAX_STEPS;
                int lo = ctl.out_hi, hi = lo;
                while (hi < n_t_eval && t_eval[hi] <= tt) ++hi;      // searchsorted(t_eval, t, side='right')
                ctl.out_lo = lo; ctl.out_hi = hi;
                ctl.status = status;
                if (status == OCCM_MEMBER_RUNNING) {
                    const double min_step = 10.0 * fabs(occm_next_up(tt) - tt);
                    if (h_abs < min_step) h_abs = min_step;
                    double hh = h_abs, t_new = tt + hh;              // occm_rk45_plan_attempt
                    if (t_new - rk.t_bound > 0.0) t_new = rk.t_bound;
                    hh = t_new - tt;
                    ctl.h = hh; ctl.h_abs = fabs(hh); ctl.t_new = t_new;
                } else ctl.h_abs = h_abs;
            } else {
                double shrink = dp::SAFETY * pow(error_norm, dp::ERROR_EXPONENT);
                shrink = shrink > dp::MIN_FACTOR ? shrink : dp::MIN_FACTOR;
                h_abs *= shrink;
                ctl.rejected = 1;
                const long long n_rej = ++ctl.n_rej;
                const double min_step = 10.0 * fabs(occm_next_up(tt) - tt);
                if (h_abs < min_step) { ctl.status = isfinite(error_norm) ? OCCM_MEMBER_STEP_TOO_SMALL : OCCM_MEMBER_NONFINITE; ctl.h_abs = h_abs; }
                else if (rk.max_steps > 0 && n_rej + ctl.n_acc >= rk.max_steps) { ctl.status = OCCM_MEMBER_MAX_STEPS; ctl.h_abs = h_abs; }
                else {
                    double hh = h_abs, t_new = tt + hh;
                    if (t_new - rk.t_bound > 0.0) t_new = rk.t_bound;
                    hh = t_new - tt;
                    ctl.h = hh; ctl.h_abs = fabs(hh); ctl.t_new = t_new;
                }
            }
        }
        __syncthreads();
        if (ctl.accepted) {
            // ---- dense output at the t_eval points in (t_old, t]  (RkDenseOutput, rk.py:575-601; occm_rk45_output)
            const int lo = ctl.out_lo, hi = ctl.out_hi;
            if (hi > lo) {
                const double t_old = ctl.t_old, hh = ctl.t - t_old;
                for (int i = tid; i < n_save; i += nthr) {
                    const int e = save_idx ? save_idx[i] : i;
                    const double k1 = K1[e], k3 = K3[e], k4 = K4[e], k5 = K5[e], k6 = K6[e], k7 = K7[e];
                    const double q0 = k1 * dp::P11;
                    const double q1 = k1 * dp::P12 + k3 * dp::P32 + k4 * dp::P42 + k5 * dp::P52 + k6 * dp::P62 + k7 * dp::P72;
                    const double q2 = k1 * dp::P13 + k3 * dp::P33 + k4 * dp::P43 + k5 * dp::P53 + k6 * dp::P63 + k7 * dp::P73;
                    const double q3 = k1 * dp::P14 + k3 * dp::P34 + k4 * dp::P44 + k5 * dp::P54 + k6 * dp::P64 + k7 * dp::P74;
                    const double y_old = y[e];
                    for (int j = lo; j < hi; ++j) {
                        const double x = (t_eval[j] - t_old) / hh;
                        const double x2 = x * x, x3 = x2 * x, x4 = x3 * x;
                        out[((long long)m * n_t_eval + j) * n_save + i] = hh * (q0 * x + q1 * x2 + q2 * x3 + q3 * x4) + y_old;
                        if (out_t && i == 0) out_t[(long long)m * n_t_eval + j] = t_eval[j];
                    }
                }
                __syncthreads();
            }
            { double* sw = y; y = yn; yn = sw; }            // accepted: y_new is the state (no copy pass)
            const int sw = i1; i1 = i7; i7 = sw;             // first same as last: K7 is the next step's K1
        }
    }
    // ---- back to the workspace: slot 0 of the ping-pong pairs holds the state
    {
        const double* K1 = Kb + i1 * ndof;
        for (int e = tid; e < ndof; e += nthr) { rk.Y[0][base + e] = y[e]; rk.KF[0][base + e] = K1[e]; }
        if (tid == 0) {
            rk.parity[m] = 0;
            rk.t[m] = ctl.t; rk.t_old[m] = ctl.t_old; rk.t_new[m] = ctl.t_new; rk.h[m] = ctl.h; rk.h_abs[m] = ctl.h_abs;
            rk.n_acc[m] = ctl.n_acc; rk.n_rej[m] = ctl.n_rej; rk.nfev[m] = ctl.nfev;
            rk.status[m] = ctl.status; rk.rejected[m] = ctl.rejected; rk.out_lo[m] = ctl.out_lo; rk.out_hi[m] = ctl.out_hi;
            rk.accepted_now[m] = 0;
        }
    }
}
