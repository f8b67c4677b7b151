/*
 * occm_b200.h — C ABI of the B200-native solver for OpenCCM compartment networks.
 *
 * This is the drop-in boundary for the reactive-transport hot path of uw-comphys/openccm.  The reference is
 * pure Python and has no FFI of its own; each entry point below names the Python interface it replaces
 * (paths relative to /root/reference/openccm).  INTEGRATION.md shows the ctypes binding and the two-line patch
 * that routes `openccm.system_solvers.solve_system` here.
 *
 * Conventions
 *   - every function returns 0 on success and a negative occm_status on failure; occm_last_error() returns a
 *     thread-local human readable message for the last failure.
 *   - no ownership transfer: every buffer is caller-owned DEVICE memory (e.g. a torch CUDA tensor's data_ptr())
 *     passed as a plain pointer + sizes; the library allocates no device memory (a few pinned host bytes only).
 *   - all kernels are launched on the caller's stream (`stream` = a cudaStream_t cast to void*, NULL = default).
 *   - state layout: y[member][point][species] (species innermost, fp64).  The reference's flattening is
 *     species-major y_ref[s * n_points + p] (system_solvers/cstr_system.py:194, pfr_system.py:290); the host
 *     side permutes at the boundary.
 *   - indices are int32 on the device (the reference uses numpy int64, pfr_system.py:146).
 */
#ifndef OCCM_B200_H
#define OCCM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OCCM_ABI_VERSION 2

typedef enum {
    OCCM_OK = 0,
    OCCM_ERR_INVALID = -1,      /* bad argument / inconsistent description           */
    OCCM_ERR_CUDA = -2,         /* a CUDA runtime call or kernel launch failed       */
    OCCM_ERR_WORKSPACE = -3,    /* caller's workspace too small                      */
    OCCM_ERR_UNSUPPORTED = -4   /* construct not supported on the device (no CPU fallback exists) */
} occm_status;

typedef enum { OCCM_MODEL_CSTR = 0, OCCM_MODEL_PFR = 1 } occm_model;

/* per-member integration status (values of occm_rk45_stats / occm_bdf_stats `status`) */
typedef enum {
    OCCM_MEMBER_RUNNING = 0,
    OCCM_MEMBER_FINISHED = 1,       /* reached t_bound                                   (solve_ivp status 0)  */
    OCCM_MEMBER_PAUSED = 2,         /* output buffer full in "all steps" mode; harvest and run again           */
    OCCM_MEMBER_STEP_TOO_SMALL = -1,/* "Required step size is less than spacing between numbers." (status -1)  */
    OCCM_MEMBER_NONFINITE = -2,     /* non-finite error estimate                                               */
    OCCM_MEMBER_MAX_STEPS = -3      /* step budget exhausted                                                   */
} occm_member_status;

/*
 * Network operator + mechanism, everything device resident.
 *
 * Replaces the `args` tuples the reference hands to solve_ivp:
 *   CSTR  (Q_div_v, c_shape, reactions, bcs)                                   cstr_system.py:149
 *   PFR   (Q_weight, _ddt0, connected_to_another_inlet, all_inlet_ids, c_shape, reactions, bcs)   pfr_system.py:190
 *
 * row_ptr/row_src/row_w: CSR *by destination model*.
 *   CSTR: dc[j,s]/dt = q_scale * sum_k row_w[k] * X_k + r_s(c[j,:]),
 *         X_k = c[row_src[k], s]                       if row_src[k] >= 0   (Q_div_v column j incl. diagonal, cstr_system.py:128-135)
 *             = bc_scale * g_{b,s}(t), b = -1-row_src[k] otherwise          (domain inlet, weight Q/V :125; generated code
 *                                                                            boundary_and_initial_conditions.py:238)
 *   PFR : node (j, i>=1): -q_scale * a[j] * (c[j,i,s] - c[j,i-1,s]) + r_s(c[j,i,:]) + pulse sources     pfr_system.py:296,308
 *         inlet node (j, 0): sum_k row_w[k] * Y_k,
 *         Y_k = d/dt c at the outlet node of PFR row_src[k] (recomputed in place)  if row_src[k] >= 0    pfr_system.py:313-314
 *             = bc_scale * g'_{b,s}(t)                                              otherwise             :191-196
 * pulse_*: CSR by model of smoothed point-mass sources applied to every non-inlet node of the model (CSTR: the node)
 *          (boundary_and_initial_conditions.py:174-180); may be NULL.
 * tables:  packed mechanism blob (stoichiometry terms, rate-expression and boundary-function byte code), layout in
 *          openccm_b200/csrc/occm_tables.h; built on the host by openccm_b200/system_solvers/tables.py from the same
 *          reactions file / CONFIG strings the reference parses (system_solvers/reactions.py:270-392).
 * fields:  optional per-point arrays referenced by rate expressions ("CFD, file" extra terms, reactions.py:216-231),
 *          fields[k * n_points + p]; may be NULL.
 */
typedef struct {
    int32_t model;            /* occm_model */
    int32_t n_species;        /* S */
    int32_t n_models;         /* N CSTRs or Np PFRs */
    int32_t points_per_model; /* 1 for CSTR, P >= 2 for PFR */
    const int32_t* row_ptr;   /* [n_models + 1] */
    const int32_t* row_src;   /* [nnz] */
    const double*  row_w;     /* [nnz] */
    const double*  a;         /* PFR: Q_pfr / dV per PFR [n_models]; NULL for CSTR */
    const int32_t* pulse_ptr; /* [n_models + 1] or NULL */
    const int32_t* pulse_fn;  /* boundary-function group per entry */
    const double*  pulse_w;
    const void*    tables;    /* device pointer to the packed table blob */
    int32_t        tables_bytes;
    const double*  fields;    /* or NULL */
    int32_t        n_fields;
    /* CSTR only, optional (ell_width = 0: every row walks the CSR arrays): an ELL copy of the rows, column-major
     * [ell_width][n_models], so the kernel reaches index and weight without the row_ptr indirection. Rows with fewer
     * entries are padded with (own index, weight 0). Rows that do not fit (more entries than ell_width, a boundary
     * entry, a pulse source) must have point_flags = 1 and are evaluated from the CSR arrays instead. */
    int32_t        ell_width;
    const int32_t* ell_src;
    const double*  ell_w;
    const uint8_t* point_flags; /* [n_points] or NULL: 1 = evaluate this point on the generic path */
    /* optional: the same ELL operator packed as 16-byte entries {int32 src; int32 flag; double w}, column-major
     * [ell_width][n_points], flag (column 0 only) = point_flags; enables the 128-bit fast kernels for even n_species.
     * PFR networks: ell_width = 1, one entry per point {src = upwind point (own index on inlet nodes), flag, w = Q/dV of
     * the PFR}; every inlet node (point % points_per_model == 0) must be flagged (ell_src / ell_w stay NULL). */
    const void*    ell_packed;
    const int32_t* flagged_points; /* [n_flagged] indices of the points with point_flags = 1 (needed with ell_packed) */
    int32_t        n_flagged;
} occm_system_desc;

/* Per-member parameters of an ensemble (all device pointers, any may be NULL = 1.0).
 * The reference has no ensemble API: sweeps are Python loops re-running `run`. */
typedef struct {
    int32_t n_members;
    const double* k_scale;   /* [n_members][n_reactions] multiplies each reaction's rate constant */
    const double* q_scale;   /* [n_members] multiplies every volumetric flow (transport operator and inlet weights) */
    const double* bc_scale;  /* [n_members] multiplies the inlet boundary values g(t) */
} occm_members;

typedef struct occm_system occm_system;   /* opaque host-side handle */

int         occm_abi_version(void);
const char* occm_last_error(void);

int occm_system_create(occm_system** out, const occm_system_desc* desc);
int occm_system_destroy(occm_system* sys);
/* n_points * n_species */
int64_t occm_system_ndof(const occm_system* sys);

/*
 * Mechanism-specialised kernels (optional).  The reference writes the reactions file as the source of one @njit function and lets
 * numba compile it per run (openccm/system_solvers/reactions.py:41-93 generate_reaction_system, 395-467); the device analogue is a
 * kernel image in which the stoichiometric terms are straight-line code: openccm_b200/specialize.py generates it from the parsed
 * mechanism and compiles openccm_b200/csrc/spec/occm_spec.cu to a cubin (nvcc, sm_100a).  `image` is that cubin (host memory,
 * copied by the driver).  The image must match the system (species count, ELL width, hash of the mechanism tables, ABI of
 * csrc/occm_spec.h), else OCCM_ERR_INVALID; CSTR networks with polynomial rate laws only, else OCCM_ERR_UNSUPPORTED.  After a
 * successful load the plain RHS and Dormand-Prince stages 2 and 7 run the image's kernels; without one the table-driven kernels
 * run.  Results are bit-identical either way.
 */
int occm_system_load_specialized(occm_system* sys, const void* image, size_t bytes);
int occm_system_is_specialized(const occm_system* sys);

/*
 * Fused right-hand side: dy[m] = ddt(t[m], y[m]) for every member in one launch.
 * Replaces `cstr_system.ddt` (cstr_system.py:161-203) / `pfr_system.ddt` (pfr_system.py:218-316) including the
 * generated `_reactions` (reactions.py:395-467) and `_boundary_conditions` (boundary_and_initial_conditions.py:211-245).
 * t_dev: device [n_members] or NULL, in which case t_scalar is used for all members.
 */
int occm_rhs(const occm_system* sys, const occm_members* mem, const double* t_dev, double t_scalar,
             const double* y, double* dy, void* stream);

/*
 * Adaptive explicit RK45 (Dormand-Prince) — a step-for-step mirror of scipy.integrate.RK45 as driven by
 * solve_ivp (scipy/integrate/_ivp/rk.py:14-71,74-180,  ivp.py:659-737), batched over members with per-member t and h.
 * Replaces `solve_ivp(ddt, t_span, c0, method='RK45', atol, rtol, first_step, t_eval)` at cstr_system.py:151 /
 * pfr_system.py:191.
 */
typedef struct {
    double t0, t_bound;
    double first_step;        /* > 0 (the reference always passes first_timestep) */
    double rtol, atol;
    const double* y0;         /* device [n_members][ndof]; copied into the workspace */
    const double* t_eval;     /* device [n_t_eval] ascending, or NULL = store every accepted step ("all") */
    int32_t n_t_eval;         /* with t_eval == NULL: capacity (rows) of `out` per member, incl. the initial state */
    const int32_t* save_idx;  /* device [n_save] flat element indices (p * S + s) to record, or NULL = all ndof */
    int32_t n_save;
    double* out;              /* device [n_members][n_t_eval][n_save or ndof] */
    double* out_t;            /* device [n_members][n_t_eval] times of the stored rows ("all" mode), may be NULL otherwise */
    int64_t max_steps;        /* per member, 0 = unlimited */
} occm_rk45_problem;

typedef struct occm_rk45 occm_rk45;   /* opaque host-side handle: carved workspace pointers + a copy of the problem */

size_t occm_rk45_workspace_bytes(const occm_system* sys, int32_t n_members);
/* Copies y0 into the (caller-owned, device) workspace, evaluates f(t0, y0), writes the initial output row in
 * "all" mode.  `mem` may be NULL (one member, no scaling); its arrays must stay valid until occm_rk45_free. */
int occm_rk45_init(occm_rk45** out, const occm_system* sys, const occm_members* mem, const occm_rk45_problem* prob,
                   void* workspace, size_t workspace_bytes, void* stream);
/* Advance until every member is finished / failed / paused (or, if max_launch_steps > 0, for at most that many
 * step launches).  n_running (host, may be NULL) receives the number of members still RUNNING. */
int occm_rk45_run(occm_rk45* rk, int64_t max_launch_steps, int32_t* n_running, void* stream);
/* "all" mode: after harvesting `out`, reset the per-member row counters and un-pause. */
int occm_rk45_resume(occm_rk45* rk, void* stream);
/* Copy per-member statistics to HOST arrays (any may be NULL): status, accepted steps, rejected steps, nfev,
 * current t, number of output rows written.  Synchronises the stream. */
int occm_rk45_stats(occm_rk45* rk, int32_t* status, int64_t* n_accepted, int64_t* n_rejected, int64_t* nfev,
                    double* t, int32_t* n_rows, void* stream);
/* Copy the current state y[m][ndof] out of the workspace (device to device). */
int occm_rk45_get_state(occm_rk45* rk, double* y_out, void* stream);
int occm_rk45_free(occm_rk45* rk);
/*
 * Variable-order BDF/NDF (orders 1..5) for stiff networks — step-size / order control of scipy.integrate.BDF
 * (scipy/integrate/_ivp/bdf.py:36-70,310-455) batched over members; the Newton systems (I - c J) dy = b are solved by
 * right-preconditioned restarted GMRES with a matrix-free Jacobian action (directional difference of the fused RHS) and an
 * on-the-fly S x S block-Jacobi preconditioner, instead of scipy's dense finite-difference Jacobian + LU.
 * Replaces `solve_ivp(..., method='BDF' | 'LSODA' | 'Radau')` at cstr_system.py:151 / pfr_system.py:191 on stiff networks.
 */
typedef struct {
    double t0, t_bound;
    double first_step;
    double rtol, atol;
    const double* y0;         /* device [n_members][ndof] */
    const double* t_eval;     /* device [n_t_eval] ascending or NULL ("all": every accepted step) */
    int32_t n_t_eval;         /* with t_eval == NULL: row capacity of `out` per member incl. the initial state */
    const int32_t* save_idx;  /* device [n_save] or NULL = all ndof */
    int32_t n_save;
    double* out;              /* device [n_members][n_t_eval][n_save or ndof] */
    double* out_t;            /* device [n_members][n_t_eval] */
    int64_t max_steps;        /* accepted + rejected attempts per member, 0 = unlimited */
    const double* tdiag;      /* device [n_models]: transport diagonal (CSTR: Q_div_v[j,j]; PFR: -Q_pfr/dV) for the preconditioner */
    int32_t gmres_restart;    /* Krylov dimension per Newton solve (0 = 30) */
    double lin_tol_factor;    /* GMRES stops at |r|_rms <= factor * newton_tol in error-weight scaling (0 = 0.05) */
    int32_t precond_mode;     /* 0: block inverses kept in the workspace (ndof * n_species doubles per member) and rebuilt only when
                                 c = h/alpha moved by > 20 %, after 20 attempts or after a slow / failed Newton solve;
                                 1: blocks rebuilt and solved on every application, nothing stored */
} occm_bdf_problem;

typedef struct occm_bdf occm_bdf;

size_t occm_bdf_workspace_bytes(const occm_system* sys, int32_t n_members, int32_t gmres_restart, int32_t precond_mode);
int occm_bdf_init(occm_bdf** out, const occm_system* sys, const occm_members* mem, const occm_bdf_problem* prob,
                  void* workspace, size_t workspace_bytes, void* stream);
/* Advance until every member is finished / failed / paused (or for at most max_attempts step attempts if > 0). */
int occm_bdf_run(occm_bdf* bdf, int64_t max_attempts, int32_t* n_running, void* stream);
int occm_bdf_resume(occm_bdf* bdf, void* stream);
/* HOST arrays, any may be NULL. nfev counts Newton and Jacobian-action RHS evaluations. */
int occm_bdf_stats(occm_bdf* bdf, int32_t* status, int64_t* n_accepted, int64_t* n_rejected, int64_t* nfev, int64_t* n_gmres,
                   int64_t* n_newton_fail, double* t, int32_t* n_rows, int32_t* order, void* stream);
/* HOST arrays [n_members]: preconditioner applications and rebuilds of the stored block inverses. */
int occm_bdf_precond_stats(occm_bdf* bdf, int64_t* n_apply, int64_t* n_factor, void* stream);
int occm_bdf_get_state(occm_bdf* bdf, double* y_out, void* stream);
int occm_bdf_free(occm_bdf* bdf);

/*
 * Residence-time distribution from recorded outlet concentrations.
 * Replaces `postprocessing.analysis.network_to_rtd` (postprocessing/analysis.py:34-103):
 *   F_s(t) = sum_k out[m][t][sel_idx[k]] * sel_w[k] (over entries with sel_species[k] == s) / q_in_total,
 *   E_s = np.gradient(F_s, t, edge_order=2);  rtd[m][s][t] = (t, E, F).
 * Additionally moments[m][s] = (m0, mean, variance) of E by the trapezoid rule (not computed by the reference).
 * out: device [n_members][n_times][n_save] as written by occm_rk45 / occm_bdf; t: device [n_times], or
 * [n_members][n_times] when t_per_member != 0.
 */
int occm_rtd(const double* out, int32_t n_members, int32_t n_times, int32_t n_save, const double* t, int32_t t_per_member,
             const int32_t* sel_idx, const double* sel_w, const int32_t* sel_species, int32_t n_sel, int32_t n_species,
             double q_in_total, double* rtd, double* moments, void* stream);

/*
 * The same from a PULSE experiment (north star item 4; the reference's precedent is the point-mass initial condition turned
 * into a short smoothed rectangular source, system_solvers/boundary_and_initial_conditions.py:174-180, used by
 * examples/OpenCMP/closed_box/CONFIG:20-27): the recorded outlet concentrations are the response to an inlet pulse u(t) of
 * known area / mean / variance, so E(t) = sum_k out * w / q_in_total / pulse_area with no differentiation, F = its running
 * trapezoid integral, and the moments are those of the response minus the pulse's own (moments of a convolution add).
 */
int occm_rtd_pulse(const double* out, int32_t n_members, int32_t n_times, int32_t n_save, const double* t, int32_t t_per_member,
                   const int32_t* sel_idx, const double* sel_w, const int32_t* sel_species, int32_t n_sel, int32_t n_species,
                   double q_in_total, double pulse_area, double pulse_mean, double pulse_variance, double* rtd, double* moments,
                   void* stream);

/*
 * Recorded solution -> mesh elements.  Replaces the time x species x dof loops of `export_to_vtu_openfoam`
 * (postprocessing/vtu_output.py:311-322) over `create_dof_to_element_map` (mesh/__init__.py:142-198):
 *   out[t][s][element] = y[t][el_dof*S + s] * el_w + y[t][el_other*S + s] * (1 - el_w);  -1 where el_dof[element] < 0.
 * y: device [n_times][ndof] (device layout, as written by occm_rk45 / occm_bdf with save_idx = NULL); out: device
 * [n_times][n_species][n_elements].  Products and sum are rounded separately (bit-identical to the numpy expression).
 */
int occm_dofs_to_elements(const double* y, int32_t n_times, int32_t n_species, int64_t ndof, int32_t n_elements,
                          const int32_t* el_dof, const int32_t* el_other, const double* el_w, double* out, void* stream);

/* Development/measurement aid: launch Dormand-Prince stage kernel `stage` (2..7) `reps` times back to back on the
 * handle's current buffers and return the mean duration (CUDA events on `stream`). Leaves y, t, h untouched. */
int occm_rk45_time_stage(occm_rk45* rk, int32_t stage, int32_t reps, float* ms_per_launch, void* stream);

/* The same for the kernels of the BDF Newton / GMRES loop, on the handle's current buffers (run a few attempts first):
 * which = 0 fused RHS at the predictor, 1 matrix-free Jacobian action (RHS of y + eps z), 2 preconditioner application. */
int occm_bdf_time_kernel(occm_bdf* bdf, int32_t which, int32_t reps, float* ms_per_launch, void* stream);

/* number of kernels this library has launched since load (bench.py's gpu_launches) */
int64_t occm_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* OCCM_B200_H */
