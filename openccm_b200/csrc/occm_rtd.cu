// Residence-time distribution on the device.
// Mirrors postprocessing/analysis.py:77-100 of the reference:
//   F_s(t) = sum_{outlets} c[s, outlet node, t] * Q_conn / sum_{inlets} Q_conn        (cumulative distribution)
//   E_s(t) = np.gradient(F_s, t, edge_order=2)                                         (2nd order, non-uniform grid)
// plus the moments the acceptance criterion asks for (the reference computes none):
//   m0 = trapz(E, t), mean = trapz(t E, t) / m0, var = trapz((t - mean)^2 E, t) / m0.
// One CTA per (member, species); everything is O(T) and tiny next to the solve.
#include "occm_host.h"

__global__ void __launch_bounds__(256)
occm_rtd_kernel(const double* __restrict__ out, int M, int T, int n_save, const double* __restrict__ t_all, int t_per_member,
                const int* __restrict__ sel_idx, const double* __restrict__ sel_w, const int* __restrict__ sel_species,
                int n_sel, int S, double q_in_total, double* __restrict__ rtd, double* __restrict__ moments) {
    __shared__ double scratch[8];
    __shared__ double mean_s;
    const int m = blockIdx.x / S, s = blockIdx.x % S;
    const double* t = t_all + (t_per_member ? (long long)m * T : 0);
    double* r = rtd + ((long long)m * S + s) * T * 3;
    // F(t)
    for (int i = threadIdx.x; i < T; i += blockDim.x) {
        double acc = 0.0;
        const double* row = out + ((long long)m * T + i) * n_save;
        for (int k = 0; k < n_sel; ++k)
            if (sel_species[k] == s) acc += row[sel_idx[k]] * sel_w[k];
        r[i * 3 + 0] = t[i];
        r[i * 3 + 2] = acc / q_in_total;
    }
    __syncthreads();
    // E(t) = dF/dt, numpy.gradient with coordinate array and edge_order=2
    for (int i = threadIdx.x; i < T; i += blockDim.x) {
        double g;
        if (i == 0) {
            const double dx1 = t[1] - t[0], dx2 = t[2] - t[1];
            const double a = -(2.0 * dx1 + dx2) / (dx1 * (dx1 + dx2)), b = (dx1 + dx2) / (dx1 * dx2), c = -dx1 / (dx2 * (dx1 + dx2));
            g = a * r[2] + b * r[5] + c * r[8];
        } else if (i == T - 1) {
            const double dx1 = t[T - 2] - t[T - 3], dx2 = t[T - 1] - t[T - 2];
            const double a = dx2 / (dx1 * (dx1 + dx2)), b = -(dx2 + dx1) / (dx1 * dx2), c = (2.0 * dx2 + dx1) / (dx2 * (dx1 + dx2));
            g = a * r[(T - 3) * 3 + 2] + b * r[(T - 2) * 3 + 2] + c * r[(T - 1) * 3 + 2];
        } else {
            const double hd = t[i + 1] - t[i], hs = t[i] - t[i - 1];
            const double a = -hd / (hs * (hd + hs)), b = (hd - hs) / (hd * hs), c = hs / (hd * (hd + hs));
            g = a * r[(i - 1) * 3 + 2] + b * r[i * 3 + 2] + c * r[(i + 1) * 3 + 2];
        }
        r[i * 3 + 1] = g;
    }
    __syncthreads();
    // moments by the trapezoid rule
    double a0 = 0.0, a1 = 0.0;
    for (int i = threadIdx.x; i + 1 < T; i += blockDim.x) {
        const double dt = t[i + 1] - t[i];
        a0 += dt * (r[i * 3 + 1] + r[(i + 1) * 3 + 1]) * 0.5;
        a1 += dt * (t[i] * r[i * 3 + 1] + t[i + 1] * r[(i + 1) * 3 + 1]) * 0.5;
    }
    const double m0 = occm_block_sum(a0, scratch);
    const double m1 = occm_block_sum(a1, scratch);
    if (threadIdx.x == 0) mean_s = m1 / m0;
    __syncthreads();
    const double mean = mean_s;
    double a2 = 0.0;
    for (int i = threadIdx.x; i + 1 < T; i += blockDim.x) {
        const double dt = t[i + 1] - t[i];
        const double d0 = t[i] - mean, d1 = t[i + 1] - mean;
        a2 += dt * (d0 * d0 * r[i * 3 + 1] + d1 * d1 * r[(i + 1) * 3 + 1]) * 0.5;
    }
    const double m2 = occm_block_sum(a2, scratch);
    if (threadIdx.x == 0) {
        double* mo = moments + ((long long)m * S + s) * 3;
        mo[0] = m0; mo[1] = mean; mo[2] = m2 / m0;
    }
}

extern "C" int occm_rtd(const double* out, int32_t n_members, int32_t n_times, int32_t n_save, const double* t,
                        int32_t t_per_member, const int32_t* sel_idx, const double* sel_w, const int32_t* sel_species,
                        int32_t n_sel, int32_t n_species, double q_in_total, double* rtd, double* moments, void* stream) {
    if (!out || !t || !sel_idx || !sel_w || !sel_species || !rtd || !moments) { occm_set_error("occm_rtd: null argument"); return OCCM_ERR_INVALID; }
    if (n_times < 3) { occm_set_error("occm_rtd needs at least 3 output times (np.gradient edge_order=2)"); return OCCM_ERR_INVALID; }
    if (!(q_in_total > 0.0)) { occm_set_error("occm_rtd: total inlet flow must be positive"); return OCCM_ERR_INVALID; }
    occm_rtd_kernel<<<n_members * n_species, 256, 0, (cudaStream_t)stream>>>(out, n_members, n_times, n_save, t, t_per_member,
                                                                            sel_idx, sel_w, sel_species, n_sel, n_species,
                                                                            q_in_total, rtd, moments);
    OCCM_LAUNCH_CHECK("occm_rtd_kernel");
    return OCCM_OK;
}

// Pulse-tracer route: the recorded outlet concentrations are the response y = E * u to a narrow inlet (or point-mass) pulse u(t)
// of known moments (area p0, mean p1, variance p2), so E(t) = y(t) / p0 directly — no differentiation of F — and F is its
// running integral.  Moments of a convolution add: m0(E) = m0(y) / p0, mean(E) = mean(y) - p1, var(E) = var(y) - p2.
// Same layout of `rtd` (t, E, F) and `moments` as occm_rtd.
__global__ void __launch_bounds__(256)
occm_rtd_pulse_kernel(const double* __restrict__ out, int M, int T, int n_save, const double* __restrict__ t_all, int t_per_member,
                      const int* __restrict__ sel_idx, const double* __restrict__ sel_w, const int* __restrict__ sel_species,
                      int n_sel, int S, double q_in_total, double p0, double p1, double p2, double* __restrict__ rtd,
                      double* __restrict__ moments) {
    __shared__ double scratch[8];
    __shared__ double mean_s;
    const int m = blockIdx.x / S, s = blockIdx.x % S;
    const double* t = t_all + (t_per_member ? (long long)m * T : 0);
    double* r = rtd + ((long long)m * S + s) * T * 3;
    for (int i = threadIdx.x; i < T; i += blockDim.x) {
        double acc = 0.0;
        const double* row = out + ((long long)m * T + i) * n_save;
        for (int k = 0; k < n_sel; ++k)
            if (sel_species[k] == s) acc += row[sel_idx[k]] * sel_w[k];
        r[i * 3 + 0] = t[i];
        r[i * 3 + 1] = acc / q_in_total / p0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {                          // F = running trapezoid integral of E (sequential: fixed summation order)
        double f = 0.0;
        r[2] = 0.0;
        for (int i = 1; i < T; ++i) { f += (t[i] - t[i - 1]) * (r[(i - 1) * 3 + 1] + r[i * 3 + 1]) * 0.5; r[i * 3 + 2] = f; }
    }
    double a0 = 0.0, a1 = 0.0;
    for (int i = threadIdx.x; i + 1 < T; i += blockDim.x) {
        const double dt = t[i + 1] - t[i];
        a0 += dt * (r[i * 3 + 1] + r[(i + 1) * 3 + 1]) * 0.5;
        a1 += dt * (t[i] * r[i * 3 + 1] + t[i + 1] * r[(i + 1) * 3 + 1]) * 0.5;
    }
    const double m0 = occm_block_sum(a0, scratch);
    const double m1 = occm_block_sum(a1, scratch);
    if (threadIdx.x == 0) mean_s = m1 / m0;
    __syncthreads();
    const double mean = mean_s;
    double a2 = 0.0;
    for (int i = threadIdx.x; i + 1 < T; i += blockDim.x) {
        const double dt = t[i + 1] - t[i];
        const double d0 = t[i] - mean, d1 = t[i + 1] - mean;
        a2 += dt * (d0 * d0 * r[i * 3 + 1] + d1 * d1 * r[(i + 1) * 3 + 1]) * 0.5;
    }
    const double m2 = occm_block_sum(a2, scratch);
    if (threadIdx.x == 0) {
        double* mo = moments + ((long long)m * S + s) * 3;
        mo[0] = m0; mo[1] = mean - p1; mo[2] = m2 / m0 - p2;
    }
}

extern "C" int occm_rtd_pulse(const double* out, int32_t n_members, int32_t n_times, int32_t n_save, const double* t,
                              int32_t t_per_member, const int32_t* sel_idx, const double* sel_w, const int32_t* sel_species,
                              int32_t n_sel, int32_t n_species, double q_in_total, double pulse_area, double pulse_mean,
                              double pulse_variance, double* rtd, double* moments, void* stream) {
    if (!out || !t || !sel_idx || !sel_w || !sel_species || !rtd || !moments) { occm_set_error("occm_rtd_pulse: null argument"); return OCCM_ERR_INVALID; }
    if (n_times < 2) { occm_set_error("occm_rtd_pulse needs at least 2 output times"); return OCCM_ERR_INVALID; }
    if (!(q_in_total > 0.0) || !(pulse_area > 0.0)) { occm_set_error("occm_rtd_pulse: total inlet flow and pulse area must be positive"); return OCCM_ERR_INVALID; }
    occm_rtd_pulse_kernel<<<n_members * n_species, 256, 0, (cudaStream_t)stream>>>(out, n_members, n_times, n_save, t, t_per_member,
                                                                                  sel_idx, sel_w, sel_species, n_sel, n_species, q_in_total,
                                                                                  pulse_area, pulse_mean, pulse_variance, rtd, moments);
    OCCM_LAUNCH_CHECK("occm_rtd_pulse_kernel");
    return OCCM_OK;
}

// ------------------------------------------------------------------------------------------------ results -> mesh elements
// Replaces the time x species x dof Python loops of `export_to_vtu_openfoam` (postprocessing/vtu_output.py:316-322):
//   field[t][s][element] = c[dof] * w + c[dof_other] * (1 - w),   -1 for elements outside every compartment (:314).
// y is the solver's recorded output [n_times][ndof] in device layout (point-major, species innermost).
// Products and sum are rounded separately (no FMA): the same bits as the reference's numpy arithmetic.
__global__ void __launch_bounds__(256)
occm_elements_kernel(const double* __restrict__ y, int n_times, int S, long long ndof, int n_elements, const int* __restrict__ el_dof,
                     const int* __restrict__ el_other, const double* __restrict__ el_w, double* __restrict__ out) {
    const long long total = (long long)n_times * n_elements;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i / n_elements);
        const int el = (int)(i - (long long)t * n_elements);
        const int dof = el_dof[el];
        double* o = out + (long long)t * S * n_elements + el;
        if (dof < 0) {
            for (int s = 0; s < S; ++s) o[(long long)s * n_elements] = -1.0;
            continue;
        }
        const int other = el_other[el];
        const double w = el_w[el], w1 = 1.0 - w;
        const double* a = y + (long long)t * ndof + (long long)dof * S;
        const double* b = y + (long long)t * ndof + (long long)other * S;
        for (int s = 0; s < S; ++s) o[(long long)s * n_elements] = __dadd_rn(__dmul_rn(a[s], w), __dmul_rn(b[s], w1));
    }
}

extern "C" int occm_dofs_to_elements(const double* y, int32_t n_times, int32_t n_species, int64_t ndof, int32_t n_elements,
                                     const int32_t* el_dof, const int32_t* el_other, const double* el_w, double* out, void* stream) {
    if (!y || !el_dof || !el_other || !el_w || !out) { occm_set_error("occm_dofs_to_elements: null argument"); return OCCM_ERR_INVALID; }
    if (n_times < 1 || n_species < 1 || n_elements < 1 || ndof < n_species || ndof % n_species) {
        occm_set_error("occm_dofs_to_elements: bad sizes"); return OCCM_ERR_INVALID;
    }
    const long long total = (long long)n_times * n_elements;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    occm_elements_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(y, n_times, n_species, ndof, n_elements, el_dof, el_other, el_w, out);
    OCCM_LAUNCH_CHECK("occm_elements_kernel");
    return OCCM_OK;
}
