// liboccm_b200.so — system handle, fused RHS kernel and the batched RK45 integrator (sm_100a).
//
// Kernels (all HBM-bound element-wise / gather kernels, fp64):
//   occm_stage<MODEL, 0>      plain fused RHS   dy = A^T c + b(t) + r(c)            (16 B per element + CSR)
//   occm_stage<MODEL, 2..7>   one Dormand-Prince stage: fused RHS of the stage state + production of the next
//                             stage state (or y_new, or the error-norm partial sums) in the same pass
//   occm_rk45_control         per-member error norm, accept/reject, step-size controller (scipy rk.py:111-167)
//   occm_rk45_output          per-member dense output at the t_eval points inside the accepted step
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include "occm_host.h"

#ifndef OCCM_E
#define OCCM_E 2          // sub-chunks of tc points per loop trip of a CTA in the Dormand-Prince stage kernels
#endif
#define OCCM_KREG 4       // ELL columns the hot path keeps in registers (wider rows take the generic path)
#ifndef OCCM_E0
#define OCCM_E0 2         // the plain RHS kernel has one HBM read per element: more sub-chunks per trip
#endif
#ifndef OCCM_MIN_CTAS
#define OCCM_MIN_CTAS 2   // stage kernels: register budget 65536 / (256 * 2) = 128 per thread (measured best on B200)
#endif
#ifndef OCCM_MIN_CTAS0
#define OCCM_MIN_CTAS0 4  // plain RHS kernel: 64 registers per thread
#endif
#define OCCM_EMAX (OCCM_E > OCCM_E0 ? OCCM_E : OCCM_E0)

#ifndef OCCM_STAGE_TU
// ------------------------------------------------------------------------------------------------ errors
static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void occm_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void occm_count_launch(int n) { g_launches += n; }

#include <mutex>
static std::mutex g_pin_mutex;
static char* g_pin_page = nullptr;          // 64 slots x 64 bytes, pinned, never freed
static unsigned long long g_pin_used = 0;   // bitmap
void* occm_pinned_acquire(void) {
    std::lock_guard<std::mutex> lock(g_pin_mutex);
    if (!g_pin_page && cudaHostAlloc((void**)&g_pin_page, 64 * 64, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); g_pin_page = nullptr; }
    if (g_pin_page)
        for (int i = 0; i < 64; ++i)
            if (!(g_pin_used >> i & 1ull)) { g_pin_used |= 1ull << i; return g_pin_page + 64 * i; }
    return calloc(1, 64);                   // pool exhausted (or no pinned memory): pageable memory works too, copies just stage
}
void occm_pinned_release(void* p) {
    if (!p) return;
    std::lock_guard<std::mutex> lock(g_pin_mutex);
    const char* c = (const char*)p;
    if (g_pin_page && c >= g_pin_page && c < g_pin_page + 64 * 64) g_pin_used &= ~(1ull << ((c - g_pin_page) / 64));
    else free(p);
}

static std::mutex g_launch_mutex;
int occm_kernel_prepare(OccmLaunchCache* c, const void* kernel, int device, int block, size_t smem, int* ctas_per_sm) {
    std::lock_guard<std::mutex> lock(g_launch_mutex);
    const int d = device >= 0 && device < OCCM_MAX_DEVICES ? device : 0;
    if (smem > 48 * 1024 && smem > c->smem_limit[d]) {      // 48 KB is available without opting in
        OCCM_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        c->smem_limit[d] = smem;
    }
    if (ctas_per_sm) {
        if (c->occ[d] == 0 || c->occ_smem[d] != smem || c->occ_block[d] != block) {
            int n = 0;
            OCCM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, block, smem));
            c->occ[d] = n < 1 ? 1 : n; c->occ_smem[d] = smem; c->occ_block[d] = block;
        }
        *ctas_per_sm = c->occ[d];
    }
    return OCCM_OK;
}


// ------------------------------------------------------------------------------------------------ mechanism-specialised images
// A cubin compiled per mechanism (spec/occm_spec.cu) is loaded with the driver API; its entry points are resolved through the
// runtime (cudaGetDriverEntryPoint), so liboccm_b200.so carries no link-time dependency on libcuda.
#include <cuda.h>
#include "occm_spec.h"
#include "occm_rk.cuh"
struct OccmDrv {
    CUresult (*moduleLoadData)(CUmodule*, const void*);
    CUresult (*moduleUnload)(CUmodule);
    CUresult (*moduleGetFunction)(CUfunction*, CUmodule, const char*);
    CUresult (*moduleGetGlobal)(CUdeviceptr*, size_t*, CUmodule, const char*);
    CUresult (*funcSetAttribute)(CUfunction, CUfunction_attribute, int);
    CUresult (*occupancy)(int*, CUfunction, int, size_t);
    CUresult (*launchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**);
    CUresult (*getErrorString)(CUresult, const char**);
    int state;      // 0 = not tried, 1 = resolved, -1 = unavailable
};
static OccmDrv g_drv;
static int occm_drv_load(void) {
    std::lock_guard<std::mutex> lock(g_launch_mutex);
    if (g_drv.state) return g_drv.state > 0 ? OCCM_OK : OCCM_ERR_CUDA;
    struct { const char* name; void** fn; } want[] = {
        {"cuModuleLoadData", (void**)&g_drv.moduleLoadData}, {"cuModuleUnload", (void**)&g_drv.moduleUnload},
        {"cuModuleGetFunction", (void**)&g_drv.moduleGetFunction}, {"cuModuleGetGlobal", (void**)&g_drv.moduleGetGlobal},
        {"cuFuncSetAttribute", (void**)&g_drv.funcSetAttribute},
        {"cuOccupancyMaxActiveBlocksPerMultiprocessor", (void**)&g_drv.occupancy},
        {"cuLaunchKernel", (void**)&g_drv.launchKernel}, {"cuGetErrorString", (void**)&g_drv.getErrorString}};
    for (auto& w : want) {
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint(w.name, w.fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !*w.fn) {
            cudaGetLastError();
            g_drv.state = -1;
            occm_set_error("driver entry point %s is not available", w.name);
            return OCCM_ERR_CUDA;
        }
    }
    g_drv.state = 1;
    return OCCM_OK;
}
static const char* occm_drv_str(CUresult r) { const char* s = nullptr; if (g_drv.getErrorString) g_drv.getErrorString(r, &s); return s ? s : "unknown driver error"; }
#define OCCM_DRV_CHECK(expr)                                                                       \
    do {                                                                                           \
        CUresult _r = (expr);                                                                      \
        if (_r != CUDA_SUCCESS) { occm_set_error("%s failed: %s", #expr, occm_drv_str(_r)); return OCCM_ERR_CUDA; } \
    } while (0)

struct OccmSpec {
    CUmodule mod;
    CUfunction fn[8];                  // by stage (0, 1, 2, 7)
    int S, K, n_term;
    int warps[8], ctas[8];             // by stage
    size_t smem_limit[8], occ_smem[8]; int occ[8];
};

void occm_spec_unload(occm_system* s) {
    if (!s->spec) return;
    if (g_drv.state > 0 && s->spec->mod) g_drv.moduleUnload(s->spec->mod);
    free(s->spec);
    s->spec = nullptr;
    s->spec_smem_kb = getenv("OCCM_B200_SPEC_SMEM_KB") ? atoi(getenv("OCCM_B200_SPEC_SMEM_KB")) : 196;   // planning cap; the carve-out is set to what the planned CTAs use
    if (s->spec_smem_kb < 32) s->spec_smem_kb = 32;
}

extern "C" int occm_system_load_specialized(occm_system* s, const void* image, size_t bytes) {
    if (!s || !image || bytes < 64) { occm_set_error("occm_system_load_specialized: null system or image"); return OCCM_ERR_INVALID; }
    const OccmSysDev& d = s->dev;
    if (s->any_prog || !s->desc.ell_packed || d.ell_k < 1 || (d.model == OCCM_MODEL_PFR && d.ell_k != 1)) {
        occm_set_error("specialised kernels cover networks with polynomial rate laws and a packed ELL operator");
        return OCCM_ERR_UNSUPPORTED;
    }
    int prev_device = -1;
    OCCM_CUDA_CHECK(cudaGetDevice(&prev_device));
    struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{prev_device == s->device ? -1 : prev_device};
    OCCM_CUDA_CHECK(cudaSetDevice(s->device));
    OCCM_CUDA_CHECK(cudaFree(0));                                      // the primary context exists and is current
    if (int rc = occm_drv_load()) return rc;
    OccmSpec* sp = (OccmSpec*)calloc(1, sizeof(OccmSpec));
    if (!sp) { occm_set_error("out of host memory"); return OCCM_ERR_INVALID; }
    CUresult r = g_drv.moduleLoadData(&sp->mod, image);
    if (r != CUDA_SUCCESS) { occm_set_error("cuModuleLoadData failed: %s (image not built for this GPU?)", occm_drv_str(r)); free(sp); return OCCM_ERR_CUDA; }
    auto fail = [&](int code) { g_drv.moduleUnload(sp->mod); free(sp); return code; };
    CUdeviceptr info_p = 0; size_t info_n = 0;
    unsigned long long info[OCCM_SI_WORDS] = {0};
    if (g_drv.moduleGetGlobal(&info_p, &info_n, sp->mod, "occm_spec_info") != CUDA_SUCCESS || info_n != sizeof(info) ||
        cudaMemcpy(info, (const void*)info_p, sizeof(info), cudaMemcpyDeviceToHost) != cudaSuccess) {
        cudaGetLastError();
        occm_set_error("image has no occm_spec_info of %zu bytes", sizeof(info));
        return fail(OCCM_ERR_INVALID);
    }
    // the image must have been generated from THIS mechanism: hash of the packed tables
    unsigned char* blob = (unsigned char*)malloc((size_t)d.tables_bytes);
    if (!blob || cudaMemcpy(blob, d.tables, (size_t)d.tables_bytes, cudaMemcpyDeviceToHost) != cudaSuccess) {
        cudaGetLastError(); free(blob);
        occm_set_error("could not read the mechanism tables back");
        return fail(OCCM_ERR_CUDA);
    }
    const unsigned long long hash = occm_spec_hash(blob, (size_t)d.tables_bytes);
    free(blob);
    if (info[OCCM_SI_ABI] != OCCM_SPEC_ABI || (int)info[OCCM_SI_MODEL] != d.model || (int)info[OCCM_SI_S] != d.S ||
        (int)info[OCCM_SI_K] != d.ell_k || (int)info[OCCM_SI_NTERM] != s->n_term || info[OCCM_SI_HASH] != hash) {
        occm_set_error("image does not match the system: abi %llu (library %d), model %llu (%d), S %llu (%d), K %llu (%d), terms %llu (%d), "
                       "table hash %016llx (%016llx)", info[OCCM_SI_ABI], OCCM_SPEC_ABI, info[OCCM_SI_MODEL], d.model, info[OCCM_SI_S], d.S,
                       info[OCCM_SI_K], d.ell_k, info[OCCM_SI_NTERM], s->n_term, info[OCCM_SI_HASH], hash);
        return fail(OCCM_ERR_INVALID);
    }
    sp->S = d.S; sp->K = d.ell_k; sp->n_term = s->n_term;
    const int stages[4] = {0, 1, 2, 7};
    const int iw[4] = {OCCM_SI_W0, OCCM_SI_W1, OCCM_SI_W2, OCCM_SI_W7}, ic[4] = {OCCM_SI_C0, OCCM_SI_C1, OCCM_SI_C2, OCCM_SI_C7};
    for (int i = 0; i < 4; ++i) {
        const int st = stages[i];
        sp->warps[st] = (int)info[iw[i]]; sp->ctas[st] = (int)info[ic[i]];
        if (sp->warps[st] < 1 || sp->warps[st] > 31 || sp->ctas[st] < 1 || sp->ctas[st] > 16) {
            occm_set_error("image reports %d warps, %d CTAs per SM for stage %d", sp->warps[st], sp->ctas[st], st);
            return fail(OCCM_ERR_INVALID);
        }
        char name[32];
        snprintf(name, sizeof(name), "occm_spec_stage%d", st);
        if (g_drv.moduleGetFunction(&sp->fn[st], sp->mod, name) != CUDA_SUCCESS) { occm_set_error("image has no kernel %s", name); return fail(OCCM_ERR_INVALID); }
    }
    occm_spec_unload(s);
    s->spec = sp;
    return OCCM_OK;
}

extern "C" int occm_system_is_specialized(const occm_system* s) { return s && s->spec ? 1 : 0; }

// One launch of a specialised stage kernel (stages 0, 1, 2, 7).  Returns 1 when the configuration does not fit it.
int occm_spec_launch(const occm_system* sys, int stage, const OccmMemDev& mem, const OccmRkDev& rk, const double* y_in,
                     double* dy_out, cudaStream_t st, int* err_stride_out, int* main_partials_out) {
    OccmSpec* sp = sys->spec;
    const int S = sp->S, K = sp->K, W = sp->warps[stage], pts = 32 * W;
    const int NS = stage == 0 ? 1 : stage == 7 ? 3 : 2, NSTG = stage == 2 ? 2 : 1;      // staged streams; staging rows per warp (one per output vector)
    if (stage <= 1 && ((uintptr_t)y_in & 15 || (uintptr_t)dy_out & 15 || (uintptr_t)mem.z & 15)) return 1;
    OccmRowGeom g;
    g.cpm = (int)((sys->dev.n_points + pts - 1) / pts);
    g.stream_bytes = pts * S * 8;
    g.slot_bytes = (NS * g.stream_bytes + K * pts * 16 + 127) & ~127;
    size_t o = 0;
    g.coef_off = 0;
    o += (size_t)W * sp->n_term * 8;
    o = (o + 15) & ~(size_t)15;
    g.bar_off = (int)o;
    o += 16 * OCCM_ROW_MAX_DEPTH;
    o = (o + 127) & ~(size_t)127;
    g.out_off = (int)o;
    o += (size_t)W * NSTG * 32 * S * 8;
    o = (o + 127) & ~(size_t)127;
    g.ring_off = (int)o;
    const size_t smem_cfg = (size_t)sys->spec_smem_kb * 1024;
    const size_t smem_sm = (size_t)sys->smem_per_sm < smem_cfg ? (size_t)sys->smem_per_sm : smem_cfg;
    const size_t cap = smem_sm / sp->ctas[stage] - 1024;
    if (cap < o + 2 * (size_t)g.slot_bytes || o + 2 * (size_t)g.slot_bytes > (size_t)sys->smem_per_block) return 1;
    int depth = (int)((cap - o) / (size_t)g.slot_bytes);
    const int want = sys->pipe_depth > 0 ? sys->pipe_depth : 6;
    if (depth > want) depth = want;
    if (depth > OCCM_ROW_MAX_DEPTH) depth = OCCM_ROW_MAX_DEPTH;
    if (depth < 2) depth = 2;
    g.depth = depth;
    const size_t smem = o + (size_t)depth * g.slot_bytes;
    const long long chunks = (long long)mem.M * g.cpm;
    if (chunks >= (1ll << 31)) { occm_set_error("too many chunks (%lld) for one launch", chunks); return OCCM_ERR_INVALID; }
    const int n_flagged = sys->desc.n_flagged;
    const int tcf = OCCM_BLOCK / S;
    const int nfix = n_flagged > 0 ? (n_flagged + tcf - 1) / tcf : 0;
    g.err_stride = g.cpm * W + nfix;
    if (stage == 7 && (size_t)g.err_stride > occm_err_capacity(sys)) {
        occm_set_error("error partials per member (%d) exceed the carved capacity (%zu)", g.err_stride, occm_err_capacity(sys));
        return OCCM_ERR_WORKSPACE;
    }
    int ctas_per_sm = 1;
    {
        std::lock_guard<std::mutex> lock(g_launch_mutex);
        if (smem > 48 * 1024 && smem > sp->smem_limit[stage]) {
            OCCM_DRV_CHECK(g_drv.funcSetAttribute(sp->fn[stage], CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem));
            sp->smem_limit[stage] = smem;
        }
        if (sp->occ[stage] == 0 || sp->occ_smem[stage] != smem) {
            // Carve out exactly what the planned CTAs need: the rest of the 256 KB stays L1, which the row gathers live on (a
            // row is five 16-byte loads of one lane; with the 28 KB of L1 the largest carve-out leaves, its line is evicted
            // between them: measured 2x slower).
            const size_t need = (size_t)sp->ctas[stage] * (smem + 1024);
            int pct = (int)((need * 100 + (size_t)sys->smem_per_sm - 1) / (size_t)sys->smem_per_sm);
            if (pct > 100) pct = 100;
            OCCM_DRV_CHECK(g_drv.funcSetAttribute(sp->fn[stage], CU_FUNC_ATTRIBUTE_PREFERRED_SHARED_MEMORY_CARVEOUT, pct));
            int n = 0;
            OCCM_DRV_CHECK(g_drv.occupancy(&n, sp->fn[stage], 32 * (W + 1), smem));
            if (n > sp->ctas[stage]) n = sp->ctas[stage];               // more CTAs would need a larger carve-out than planned
            sp->occ[stage] = n < 1 ? 1 : n; sp->occ_smem[stage] = smem;
        }
        ctas_per_sm = sp->occ[stage];
    }
    const int grid = occm_grid(sys, chunks, ctas_per_sm);
    const void* ell = sys->desc.ell_packed;
    void* args[] = {(void*)&sys->dev, (void*)&mem, (void*)&rk, (void*)&y_in, (void*)&dy_out, (void*)&ell, (void*)&g};
    OCCM_DRV_CHECK(g_drv.launchKernel(sp->fn[stage], (unsigned)grid, 1, 1, (unsigned)(32 * (W + 1)), 1, 1, (unsigned)smem, (CUstream)st, args, nullptr));
    occm_count_launch();
    *err_stride_out = g.err_stride;
    *main_partials_out = g.cpm * W;
    return OCCM_OK;
}

extern "C" const char* occm_last_error(void) { return g_err; }
extern "C" int occm_abi_version(void) { return OCCM_ABI_VERSION; }
extern "C" int64_t occm_launch_count(void) { return g_launches.load(); }

// ------------------------------------------------------------------------------------------------ system
extern "C" int occm_system_create(occm_system** out, const occm_system_desc* d) {
    if (!out || !d) { occm_set_error("occm_system_create: null argument"); return OCCM_ERR_INVALID; }
    if (d->model != OCCM_MODEL_CSTR && d->model != OCCM_MODEL_PFR) { occm_set_error("unknown model %d", d->model); return OCCM_ERR_INVALID; }
    if (d->n_species < 1 || d->n_species > OCCM_MAX_S) { occm_set_error("n_species=%d outside [1, %d]", d->n_species, OCCM_MAX_S); return OCCM_ERR_INVALID; }
    if (d->n_models < 1) { occm_set_error("n_models must be >= 1"); return OCCM_ERR_INVALID; }
    if (d->model == OCCM_MODEL_CSTR && d->points_per_model != 1) { occm_set_error("CSTR networks have one point per model"); return OCCM_ERR_INVALID; }
    if (d->model == OCCM_MODEL_PFR && (d->points_per_model < 2 || !d->a)) { occm_set_error("PFR networks need points_per_model >= 2 and the a = Q/dV array"); return OCCM_ERR_INVALID; }
    if (!d->row_ptr || !d->tables || d->tables_bytes < OCCM_H_WORDS * 4 || (d->tables_bytes & 7)) { occm_set_error("row_ptr/tables missing or malformed"); return OCCM_ERR_INVALID; }
    int hdr[OCCM_H_WORDS];
    OCCM_CUDA_CHECK(cudaMemcpy(hdr, d->tables, sizeof(hdr), cudaMemcpyDeviceToHost));
    if (hdr[OCCM_H_MAGIC] != OCCM_TABLES_MAGIC || hdr[OCCM_H_BYTES] != d->tables_bytes || hdr[OCCM_H_S] != d->n_species) {
        occm_set_error("table blob header mismatch (magic %x, bytes %d vs %d, S %d vs %d)", hdr[OCCM_H_MAGIC], hdr[OCCM_H_BYTES],
                       d->tables_bytes, hdr[OCCM_H_S], d->n_species);
        return OCCM_ERR_INVALID;
    }
    long long n_points = (long long)d->n_models * d->points_per_model;
    long long ndof = n_points * d->n_species;
    if (n_points >= (1ll << 31) - 2) { occm_set_error("too many points for int32 indices"); return OCCM_ERR_INVALID; }
    occm_system* s = (occm_system*)calloc(1, sizeof(occm_system));
    s->desc = *d;
    OccmSysDev& v = s->dev;
    v.model = d->model; v.S = d->n_species; v.n_models = d->n_models; v.P = d->points_per_model;
    v.n_points = n_points; v.ndof = ndof;
    v.row_ptr = d->row_ptr; v.row_src = d->row_src; v.row_w = d->row_w; v.a = d->a;
    v.pulse_ptr = d->pulse_ptr; v.pulse_fn = d->pulse_fn; v.pulse_w = d->pulse_w;
    v.tables = (const int*)d->tables; v.tables_bytes = d->tables_bytes;
    v.fields = d->fields; v.n_fields = d->n_fields;
    v.n_rxn = hdr[OCCM_H_NRXN];
    v.tc = OCCM_BLOCK / d->n_species;
    v.chunks_per_member = (int)((n_points + (long long)OCCM_E * v.tc - 1) / ((long long)OCCM_E * v.tc));   // Dormand-Prince stage kernels (err_partial layout)
    // PFR: the only ELL form is the packed one-entry row per point {upwind point, flag, Q/dV} of the fast kernels
    v.ell_k = d->model == OCCM_MODEL_CSTR ? d->ell_width : (d->ell_packed && d->ell_width == 1 && d->point_flags ? 1 : 0);
    v.ell_src = d->ell_src; v.ell_w = d->ell_w;
    v.point_flags = d->point_flags;
    s->use_fast = getenv("OCCM_B200_NO_FAST") ? 0 : 1;
    s->use_small = getenv("OCCM_B200_NO_SMALL") ? 0 : 1;        // tests / A-B: batched kernels also for networks that fit in shared memory
    s->use_row = getenv("OCCM_B200_NO_SPEC") ? 0 : 1;           // tests / A-B: ignore a loaded mechanism-specialised image
    s->spec = nullptr;
    s->spec_smem_kb = getenv("OCCM_B200_SPEC_SMEM_KB") ? atoi(getenv("OCCM_B200_SPEC_SMEM_KB")) : 196;   // planning cap; the carve-out is set to what the planned CTAs use
    if (s->spec_smem_kb < 32) s->spec_smem_kb = 32;
    s->any_prog = hdr[OCCM_H_ANY_PROG];
    s->n_term = hdr[OCCM_H_NTERM];
    s->use_pipe = getenv("OCCM_B200_NO_PIPE") ? 0 : 1;          // tests / A-B measurements: register-prefetching fast kernels instead
    s->pipe_depth = getenv("OCCM_B200_PIPE_DEPTH") ? atoi(getenv("OCCM_B200_PIPE_DEPTH")) : 0;
    s->force_wide = getenv("OCCM_B200_FORCE_WIDE") ? 1 : 0;      // tests: run the shared-memory term-list variant on any mechanism
    s->max_terms = hdr[OCCM_H_MAX_TERMS];
    v.fast_ok = hdr[OCCM_H_MAX_TERMS] <= OCCM_NT && hdr[OCCM_H_MAX_SLOTS] <= OCCM_NF && !hdr[OCCM_H_ANY_PROG] && hdr[OCCM_H_NRXN] <= 255;
    v.wide_ok = hdr[OCCM_H_MAX_TERMS] <= OCCM_WIDE_MAX_TERMS && hdr[OCCM_H_MAX_SLOTS] <= OCCM_NF && !hdr[OCCM_H_ANY_PROG] &&
                hdr[OCCM_H_NRXN] <= 255 && d->n_species <= 254;
    if (d->model == OCCM_MODEL_CSTR && (v.ell_k == 0 || !d->point_flags)) { v.fast_ok = 0; v.wide_ok = 0; }   // no ELL copy: CSR walk for every row
    if (v.ell_k < 0 || v.ell_k > OCCM_KREG || (d->model == OCCM_MODEL_CSTR && v.ell_k > 0 && (!d->ell_src || !d->ell_w))) { occm_set_error("bad ELL description"); free(s); return OCCM_ERR_INVALID; }
    if (ndof >= (1ll << 31)) { occm_set_error("ndof per member must fit in int32"); free(s); return OCCM_ERR_INVALID; }
    s->smem_tables = (size_t)d->tables_bytes;
    {   // entries of the operator and of the pulse lists (the persistent small-network kernel stages them in shared memory)
        int last = 0;
        OCCM_CUDA_CHECK(cudaMemcpy(&last, d->row_ptr + d->n_models, sizeof(int), cudaMemcpyDeviceToHost));
        s->nnz = last;
        s->pulse_nnz = 0;
        if (d->pulse_ptr) { OCCM_CUDA_CHECK(cudaMemcpy(&last, d->pulse_ptr + d->n_models, sizeof(int), cudaMemcpyDeviceToHost)); s->pulse_nnz = last; }
    }
    int dev = 0;
    OCCM_CUDA_CHECK(cudaGetDevice(&dev));
    s->device = dev;
    OCCM_CUDA_CHECK(cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, dev));
    int max_smem = 0;
    OCCM_CUDA_CHECK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    s->smem_per_block = max_smem;
    OCCM_CUDA_CHECK(cudaDeviceGetAttribute(&s->smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
    if (occm_elem_smem(s, OCCM_EMAX) > (size_t)max_smem) {
        occm_set_error("mechanism tables (%d B) do not fit in shared memory (%d B)", d->tables_bytes, max_smem);
        free(s);
        return OCCM_ERR_UNSUPPORTED;
    }
    *out = s;
    return OCCM_OK;
}

extern "C" int occm_system_destroy(occm_system* s) { if (s) occm_spec_unload(s); free(s); return OCCM_OK; }
extern "C" int64_t occm_system_ndof(const occm_system* s) { return s ? s->dev.ndof : -1; }

#endif  // !OCCM_STAGE_TU

// ------------------------------------------------------------------------------------------------ RK45 data
#include "occm_rk.cuh"

struct occm_rk45 {
    const occm_system* sys;
    occm_members mem; bool has_mem;
    occm_rk45_problem prob;
    OccmRkDev d;
    int M;
    int* h_running;     // pinned host
    long long launched_steps;
    // CUDA graph of OCCM_GRAPH_STEPS consecutive step launches (launch-bound small systems); built lazily
    cudaGraphExec_t graph_exec;
    int graph_state;          // 0 = not tried, 1 = ready, -1 = unavailable (fall back to direct launches)
    int launches_per_step;
};

// ------------------------------------------------------------------------------------------------ stage kernel
// STAGE 0: plain RHS (y_in -> dy_out).  STAGE s in 2..7: Dormand-Prince stage s of rk_step (rk.py:14-71).
//
// Hot path per element (CSTR): one coalesced own load, K (ELL width) index/weight loads + K row gathers that hit L2
// (a member's state is a few MB), the reaction terms from registers + the staged row, the stage epilogue (own loads of
// y, K_i; stores of K_s and the next stage state).  No branches except the per-point `flags` test that routes the rare
// rows with boundary / pulse / overflow entries to the generic path.
template <int STAGE>
struct OccmStageIn {       // the stage state this kernel differentiates, as a function of the flat element index
    const double* a; const double* b; double hb;
    __device__ __forceinline__ double operator()(int e) const {
        if (STAGE == 2) return __ldg(a + e) + (dp::A21 * __ldg(b + e)) * hb;   // y + h * a21 * K1 formed on the fly
        if (STAGE == 1) return __ldg(a + e) + __ldg(b + e) * hb;               // y + eps * z (matrix-free Jacobian action)
        return __ldg(a + e);
    }
};

// Own-element operands of the stage epilogue, loaded at the top of a trip so that all HBM reads of a trip are in
// flight together.
struct OccmOps { double ye, k1, ka, kb, kc, kd; };

template <int STAGE>
__device__ __forceinline__ OccmOps occm_load_ops(const OccmRkDev& rk, const double* y, const double* k1, long long base, int e) {
    OccmOps o;
    o.ye = o.k1 = o.ka = o.kb = o.kc = o.kd = 0.0;
    if (STAGE == 0 || STAGE == 1) return o;
    o.ye = __ldg(y + e);
    if (STAGE != 7) o.k1 = __ldg(k1 + e);
    if (STAGE == 3) { o.ka = __ldg(rk.K2 + base + e); }
    if (STAGE == 4) { o.ka = __ldg(rk.K2 + base + e); o.kb = __ldg(rk.K3 + base + e); }
    if (STAGE == 5) { o.ka = __ldg(rk.K2 + base + e); o.kb = __ldg(rk.K3 + base + e); o.kc = __ldg(rk.K4 + base + e); }
    if (STAGE == 6) { o.ka = __ldg(rk.K3 + base + e); o.kb = __ldg(rk.K4 + base + e); o.kc = __ldg(rk.K5 + base + e); }
    if (STAGE == 7) { o.ka = __ldg(rk.K2 + base + e); }      // K2's buffer holds the partial error combination written by stage 6
    return o;
}

// Stores K_s and the next stage state (rk.py:59-69); stage 7 returns the squared scaled error of the element.
template <int STAGE>
__device__ __forceinline__ double occm_finish(const OccmRkDev& rk, const OccmOps& o, double f, double h, int par, long long base,
                                              int e, double y_new_own, double* __restrict__ dy_out) {
    if (STAGE == 0 || STAGE == 1) { dy_out[base + e] = f; return 0.0; }
    if (STAGE == 2) {
        rk.K2[base + e] = f;
        rk.YS[0][base + e] = o.ye + (dp::A31 * o.k1 + dp::A32 * f) * h;
    } else if (STAGE == 3) {
        rk.K3[base + e] = f;
        rk.YS[1][base + e] = o.ye + (dp::A41 * o.k1 + dp::A42 * o.ka + dp::A43 * f) * h;
    } else if (STAGE == 4) {
        rk.K4[base + e] = f;
        rk.YS[0][base + e] = o.ye + (dp::A51 * o.k1 + dp::A52 * o.ka + dp::A53 * o.kb + dp::A54 * f) * h;
    } else if (STAGE == 5) {
        rk.K5[base + e] = f;
        rk.YS[1][base + e] = o.ye + (dp::A61 * o.k1 + dp::A62 * o.ka + dp::A63 * o.kb + dp::A64 * o.kc + dp::A65 * f) * h;
    } else if (STAGE == 6) {
        rk.K6[base + e] = f;
        // y_new = y + h * dot(K[:-1].T, B)   (rk.py:64)
        rk.Y[par ^ 1][base + e] = o.ye + h * (dp::B1 * o.k1 + dp::B3 * o.ka + dp::B4 * o.kb + dp::B5 * o.kc + dp::B6 * f);
        // first six terms of the error combination dot(K.T, E) (rk.py:157): K2 is dead after stage 5 (b2 = e2 = 0, not used
        // by the dense output), so its buffer carries the partial sum to stage 7 and saves four vector reads there
        rk.K2[base + e] = dp::E1 * o.k1 + dp::E3 * o.ka + dp::E4 * o.kb + dp::E5 * o.kc + dp::E6 * f;
    } else {  // STAGE 7: f_new = K7, error estimate (rk.py:66-69, 157-162)
        rk.KF[par ^ 1][base + e] = f;
        const double err = (o.ka + dp::E7 * f) * h;
        const double scale = rk.atol + fmax(fabs(o.ye), fabs(y_new_own)) * rk.rtol;
        const double q = err / scale;
        return q * q;
    }
    return 0.0;
}

// Cold path of one element: generic evaluation + epilogue (kept out of line so the hot loop carries no call and few
// live values: everything it needs is rebuilt from (member, parity)).
template <int MODEL, int STAGE>
__device__ __noinline__ double occm_cold_element(const OccmSysDev& sys, const OccmMemDev& mem, const OccmRkDev& rk,
                                                 const int* tab_s, const double* y_in, const double* myrow, int p, int s,
                                                 double t_stage, int m, const double* ksc, double* dy_out) {
    const long long base = (long long)m * sys.ndof;
    const int RS = sys.S + 1;
    double h = 0.0; int par = 0;
    const double* y = nullptr; const double* k1 = nullptr;
    if (STAGE >= 2) { h = rk.h[m]; par = rk.parity[m]; y = rk.Y[par] + base; k1 = rk.KF[par] + base; }
    if (STAGE == 1) { h = __ldg(mem.eps + m); k1 = mem.z + base; }
    OccmStageIn<STAGE> in;
    in.a = STAGE <= 1 ? y_in + base : STAGE == 2 ? y
         : STAGE == 3 || STAGE == 5 ? rk.YS[0] + base
         : STAGE == 4 || STAGE == 6 ? rk.YS[1] + base : rk.Y[par ^ 1] + base;
    in.b = k1; in.hb = h;
    const double qs = mem.q_scale ? __ldg(mem.q_scale + m) : 1.0;
    const double bs = mem.bc_scale ? __ldg(mem.bc_scale + m) : 1.0;
    const int e = p * sys.S + s;
    const double f = occm_point_generic<MODEL>(sys, tab_s, in, myrow, myrow - RS, p, s, t_stage, qs, bs, ksc);
    const OccmOps o = occm_load_ops<STAGE>(rk, y, k1, base, e);
    return occm_finish<STAGE>(rk, o, f, h, par, base, e, myrow[s], dy_out);
}

template <int MODEL, int STAGE, int E>
__global__ void __launch_bounds__(OCCM_BLOCK, (STAGE <= 1 ? OCCM_MIN_CTAS0 : OCCM_MIN_CTAS))
occm_stage(const __grid_constant__ OccmSysDev sys, const __grid_constant__ OccmMemDev mem, const __grid_constant__ OccmRkDev rk,
           const double* __restrict__ t_dev, double t_scalar, const double* __restrict__ y_in, double* __restrict__ dy_out) {
    extern __shared__ double smem_d[];
    int* tab_s = reinterpret_cast<int*>(smem_d);
    occm_load_tables(sys.tables, sys.tables_bytes, tab_s);
    const int S = sys.S, TC = sys.tc, RS = S + 1;           // staged rows carry a trailing 1.0
    const int tst = (TC + 1) * RS;                           // [upwind halo row | TC rows] per sub-chunk
    double* tile = smem_d + (sys.tables_bytes >> 3);
    double* ksc_s = tile + E * tst;                          // k_scale row of the current member (or ones)
    double* rxc_s = ksc_s + ((sys.n_rxn + 1) & ~1);          // padded per-species coefficients [S][OCCM_NT]
    double* scratch = rxc_s + S * OCCM_NT;
    const int tid = threadIdx.x;
    const bool lane_on = tid < TC * S;
    const int pl = tid / S;
    const int s = tid - pl * S;
    const int nt_u = tab_s[OCCM_H_MAX_TERMS], nf_u = tab_s[OCCM_H_MAX_SLOTS];
    const bool fast = sys.fast_ok != 0;
    unsigned meta[OCCM_NT];
    {
        const OccmTab T = occm_tab_views(tab_s);
        const OccmRxnRegs R = occm_rxn_regs(T, lane_on ? s : 0, S);
#pragma unroll
        for (int q = 0; q < OCCM_NT; ++q) { meta[q] = R.meta[q]; if (lane_on && pl == 0) rxc_s[s * OCCM_NT + q] = R.coeff[q]; }
    }
    for (int r = tid; r < E * (TC + 1); r += OCCM_BLOCK) tile[r * RS + S] = 1.0;
    for (int r = tid; r < sys.n_rxn; r += OCCM_BLOCK) ksc_s[r] = 1.0;
    const int n_points = (int)sys.n_points;
    const int cpm = sys.chunks_per_member;
    const int K = sys.ell_k, N = sys.n_models;
    double* const myslot = tile + (pl + 1) * RS + s;         // my slot inside sub-chunk 0's tile
    const double* const myrow0 = tile + (pl + 1) * RS;
    __syncthreads();

    // (member, chunk) cursor of this CTA: advanced by gridDim.x chunks per trip without a division in the loop
    int m = (int)(blockIdx.x / (unsigned)cpm);
    int ci = (int)(blockIdx.x - (unsigned)m * (unsigned)cpm);
    const int step_m = (int)(gridDim.x / (unsigned)cpm), step_c = (int)(gridDim.x - (unsigned)step_m * (unsigned)cpm);
    int coeff_m = -1;                                        // member whose k_scale is folded into coeffm
    double coeffm[OCCM_NT];
#pragma unroll
    for (int q = 0; q < OCCM_NT; ++q) coeffm[q] = 0.0;

    for (; m < mem.M; m += step_m, ci += step_c, m += (ci >= cpm), ci -= (ci >= cpm) ? cpm : 0) {
        if (STAGE >= 2 && rk.status[m] != OCCM_MEMBER_RUNNING) continue;   // uniform over the CTA
        if (STAGE <= 1 && mem.active && !mem.active[m]) continue;
        const long long base = (long long)m * sys.ndof;
        double t_stage, h = 0.0;
        int par = 0;
        const double* y = nullptr; const double* k1 = nullptr;
        if (STAGE <= 1) {
            t_stage = t_dev ? __ldg(t_dev + m) : t_scalar;
            if (STAGE == 1) { h = __ldg(mem.eps + m); k1 = mem.z + base; }
        } else {
            const double t0 = rk.t[m];
            h = rk.h[m];
            par = rk.parity[m];
            y = rk.Y[par] + base;
            k1 = rk.KF[par] + base;
            t_stage = STAGE == 2 ? t0 + dp::C2 * h : STAGE == 3 ? t0 + dp::C3 * h : STAGE == 4 ? t0 + dp::C4 * h
                    : STAGE == 5 ? t0 + dp::C5 * h : t0 + h;
        }
        OccmStageIn<STAGE> in;
        in.a = STAGE <= 1 ? y_in + base : STAGE == 2 ? y
             : STAGE == 3 || STAGE == 5 ? rk.YS[0] + base
             : STAGE == 4 || STAGE == 6 ? rk.YS[1] + base : rk.Y[par ^ 1] + base;
        in.b = k1; in.hb = h;

        // ---- phase A: every load whose address is known up front is issued here: own stage values, epilogue
        //      operands, point flags and the ELL index / weight columns
        const int p0 = ci * (E * TC);
        int pt[E]; double own[E]; unsigned slow[E]; OccmOps ops[E];
        int esrc[E][OCCM_KREG]; double ew[E][OCCM_KREG];
#pragma unroll
        for (int i = 0; i < E; ++i) {
            pt[i] = p0 + i * TC + pl;
            if (!(lane_on && pt[i] < n_points)) pt[i] = -1;
            const bool v = pt[i] >= 0;
            own[i] = v ? in(pt[i] * S + s) : 0.0;
            slow[i] = (v && sys.point_flags) ? (unsigned)__ldg(sys.point_flags + pt[i]) : 0u;
            if (v) ops[i] = occm_load_ops<STAGE>(rk, y, k1, base, pt[i] * S + s);
            if (MODEL == 0) {
                const int* cs = sys.ell_src + (v ? pt[i] : 0);
                const double* cw = sys.ell_w + (v ? pt[i] : 0);
#pragma unroll
                for (int k = 0; k < OCCM_KREG; ++k) {
                    const bool on = v && k < K;
                    esrc[i][k] = on ? __ldg(cs + k * N) : 0;
                    ew[i][k] = on ? __ldg(cw + k * N) : 0.0;
                }
            }
        }
        if (mem.k_scale && coeff_m != m)                 // any number of reactions (mechanisms with > 255 always run this kernel)
            for (int r = tid; r < sys.n_rxn; r += OCCM_BLOCK) ksc_s[r] = __ldg(mem.k_scale + (long long)m * sys.n_rxn + r);
        if (MODEL == 1 && tid < S) {           // upwind neighbour point of each sub-chunk's first point
#pragma unroll
            for (int i = 0; i < E; ++i) {
                const int pp = p0 + i * TC - 1;
                tile[i * tst + tid] = (pp >= 0 && pp < n_points) ? in(pp * S + tid) : 0.0;
            }
        }
#pragma unroll
        for (int i = 0; i < E; ++i) if (lane_on) myslot[i * tst] = own[i];
        __syncthreads();

        // ---- phase B: all row gathers of the trip in flight together (they hit L2: a member's state is a few MB)
        double g[E][OCCM_KREG];
        if (MODEL == 0) {
#pragma unroll
            for (int i = 0; i < E; ++i)
#pragma unroll
                for (int k = 0; k < OCCM_KREG; ++k)
                    g[i][k] = (pt[i] >= 0 && k < K && ew[i][k] != 0.0) ? in(esrc[i][k] * S + s) : 0.0;
        }
        if (coeff_m != m) {                     // uniform: fold the member's rate-constant scales into the coefficients
#pragma unroll
            for (int q = 0; q < OCCM_NT; ++q) coeffm[q] = rxc_s[s * OCCM_NT + q] * ksc_s[meta[q] & 0xffu];
            coeff_m = m;
        }
        const double qs = mem.q_scale ? __ldg(mem.q_scale + m) : 1.0;

        double err2 = 0.0;
        unsigned any_slow = 0;
#pragma unroll
        for (int i = 0; i < E; ++i) {
            if (pt[i] < 0) continue;
            const double* myrow = myrow0 + i * tst;
            double f;
            if (MODEL == 0) {
                if (!fast || slow[i]) { slow[i] = 1; any_slow = 1; continue; }
                const double r = occm_reaction_fast(coeffm, meta, myrow, nt_u, nf_u);   // while the gathers fly
                double acc = 0.0;
#pragma unroll
                for (int k = 0; k < OCCM_KREG; ++k) acc += ew[i][k] * g[i][k];
                f = __dadd_rn(__dmul_rn(qs, acc), r);
            } else {
                const int P = sys.P;
                const int k = pt[i] / P;
                const int node = pt[i] - k * P;
                if (!fast || slow[i] || node == 0) { slow[i] = 1; any_slow = 1; continue; }
                const double conv = __dmul_rn(-(qs * __ldg(sys.a + k)), own[i] - myrow[s - RS]);
                f = __dadd_rn(conv, occm_reaction_fast(coeffm, meta, myrow, nt_u, nf_u));
            }
            err2 += occm_finish<STAGE>(rk, ops[i], f, h, par, base, pt[i] * S + s, own[i], dy_out);
        }
        // ---- cold path (rows with boundary / pulse / overflow entries, PFR inlet nodes, table-walk mechanisms)
        if (any_slow) {
#pragma unroll 1
            for (int i = 0; i < E; ++i)
                if (pt[i] >= 0 && slow[i])
                    err2 += occm_cold_element<MODEL, STAGE>(sys, mem, rk, tab_s, y_in, myrow0 + i * tst, pt[i], s, t_stage, m,
                                                            mem.k_scale ? ksc_s : nullptr, dy_out);
        }
        if (STAGE == 7) {
            const double sum = occm_block_sum(err2, scratch);   // contains __syncthreads
            if (tid == 0) rk.err_partial[(long long)m * cpm + ci] = sum;
        } else {
            __syncthreads();   // tile / ksc_s are overwritten by the next trip
        }
    }
}

#include "occm_fast.cuh"
#include "occm_pipe.cuh"

#ifndef OCCM_STAGE_TU
// ------------------------------------------------------------------------------------------------ controller
// np.nextafter(t, +inf) for t >= 0 (the reference asserts t0 >= 0, helper_functions.py:82)
__device__ __forceinline__ double occm_next_up(double t) { return __longlong_as_double(__double_as_longlong(t) + 1ll); }

__device__ __forceinline__ void occm_rk45_plan_attempt(const OccmRkDev& rk, int m, double t, double h_abs) {
    // rk.py:125-135 (direction = +1): h = h_abs; t_new = t + h; clip to t_bound
    double h = h_abs;
    double t_new = t + h;
    if (t_new - rk.t_bound > 0.0) t_new = rk.t_bound;
    h = t_new - t;
    rk.h[m] = h;
    rk.h_abs[m] = fabs(h);
    rk.t_new[m] = t_new;
}

__global__ void __launch_bounds__(128)
occm_rk45_control(const OccmRkDev rk, int cpm, long long ndof, const double* __restrict__ t_eval, int n_t_eval) {
    __shared__ double scratch[8];
    const int m = blockIdx.x;
    if (rk.status[m] != OCCM_MEMBER_RUNNING) {
        if (threadIdx.x == 0) rk.accepted_now[m] = 0;
        return;
    }
    // deterministic reduction of the per-chunk partial sums
    double v = 0.0;
    for (int i = threadIdx.x; i < cpm; i += blockDim.x) v += rk.err_partial[(long long)m * cpm + i];
    const double sum = occm_block_sum(v, scratch);
    if (threadIdx.x != 0) return;

    const double error_norm = sqrt(sum / (double)ndof);       // norm(x) = ||x||_2 / sqrt(n)   (common.py:64-66)
    double t = rk.t[m];
    double h_abs = rk.h_abs[m];
    rk.nfev[m] += 6;
    if (error_norm < 1.0) {                                    // rk.py:150-161 accepted
        double factor = error_norm == 0.0 ? dp::MAX_FACTOR : fmin(dp::MAX_FACTOR, dp::SAFETY * pow(error_norm, dp::ERROR_EXPONENT));
        if (rk.rejected[m]) factor = fmin(1.0, factor);
        h_abs *= factor;
        rk.t_old[m] = t;
        t = rk.t_new[m];
        rk.t[m] = t;
        rk.parity[m] ^= 1;
        rk.rejected[m] = 0;
        rk.accepted_now[m] = 1;
        const long long n_acc = ++rk.n_acc[m];
        int status = OCCM_MEMBER_RUNNING;
        if (t - rk.t_bound >= 0.0) status = OCCM_MEMBER_FINISHED;        // base.py:193-197
        else if (rk.max_steps > 0 && n_acc + rk.n_rej[m] >= rk.max_steps) status = OCCM_MEMBER_MAX_STEPS;
        // output bookkeeping (ivp.py:701-730)
        if (t_eval) {
            int lo = rk.out_hi[m], hi = lo;
            while (hi < n_t_eval && t_eval[hi] <= t) ++hi;      // searchsorted(t_eval, t, side='right')
            rk.out_lo[m] = lo; rk.out_hi[m] = hi;
        } else {
            const int row = rk.n_rows[m];
            rk.out_lo[m] = row; rk.n_rows[m] = row + 1;
            if (status == OCCM_MEMBER_RUNNING && row + 1 >= n_t_eval) status = OCCM_MEMBER_PAUSED;
        }
        rk.status[m] = status;
        if (status == OCCM_MEMBER_RUNNING || status == OCCM_MEMBER_PAUSED) {
            // start of the next _step_impl (rk.py:113-120): h_abs is clipped from below to 10 ulp(t)
            const double min_step = 10.0 * fabs(occm_next_up(t) - t);
            if (h_abs < min_step) h_abs = min_step;
            occm_rk45_plan_attempt(rk, m, t, h_abs);
        }
    } else {                                                   // rk.py:163-166 rejected (NaN norms land here too)
        double shrink = dp::SAFETY * pow(error_norm, dp::ERROR_EXPONENT);
        shrink = shrink > dp::MIN_FACTOR ? shrink : dp::MIN_FACTOR;   // Python max(MIN_FACTOR, x): NaN -> MIN_FACTOR
        h_abs *= shrink;
        rk.rejected[m] = 1;
        rk.accepted_now[m] = 0;
        const long long n_rej = ++rk.n_rej[m];
        const double min_step = 10.0 * fabs(occm_next_up(t) - t);
        if (h_abs < min_step) rk.status[m] = isfinite(error_norm) ? OCCM_MEMBER_STEP_TOO_SMALL : OCCM_MEMBER_NONFINITE;
        else if (rk.max_steps > 0 && n_rej + rk.n_acc[m] >= rk.max_steps) rk.status[m] = OCCM_MEMBER_MAX_STEPS;
        else occm_rk45_plan_attempt(rk, m, t, h_abs);
    }
}

__global__ void occm_rk45_init_members(const OccmRkDev rk, int M, double t0, double first_step, int all_mode,
                                       const double* __restrict__ t_eval, int n_t_eval) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    rk.t[m] = t0; rk.t_old[m] = t0;
    rk.parity[m] = 0; rk.status[m] = OCCM_MEMBER_RUNNING; rk.rejected[m] = 0; rk.accepted_now[m] = 0;
    rk.n_acc[m] = 0; rk.n_rej[m] = 0; rk.nfev[m] = 1;       // f(t0, y0)
    rk.out_lo[m] = 0; rk.out_hi[m] = 0; rk.n_rows[m] = all_mode ? 1 : 0;
    double h_abs = first_step;
    const double min_step = 10.0 * fabs(occm_next_up(t0) - t0);
    if (h_abs < min_step) h_abs = min_step;
    occm_rk45_plan_attempt(rk, m, t0, h_abs);
    if (m == 0) *rk.n_running = M;
}

__global__ void occm_count_running(const int* __restrict__ status, int M, int* out) {
    __shared__ int cnt;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    int local = 0;
    for (int m = threadIdx.x; m < M; m += blockDim.x) local += status[m] == OCCM_MEMBER_RUNNING;
    atomicAdd(&cnt, local);
    __syncthreads();
    if (threadIdx.x == 0) *out = cnt;
}

// ------------------------------------------------------------------------------------------------ output kernel
// t_eval mode: RkDenseOutput (rk.py:575-601) at every t_eval in (t_old, t]; "all" mode: copy y_new.
__global__ void __launch_bounds__(OCCM_BLOCK)
occm_rk45_output(const OccmRkDev rk, int M, long long ndof, int n_save, const int* __restrict__ save_idx,
                 const double* __restrict__ t_eval, int n_t_eval, double* __restrict__ out, double* __restrict__ out_t) {
    const int cps = (n_save + OCCM_BLOCK - 1) / OCCM_BLOCK;
    const long long total = (long long)M * cps;
    for (long long c = blockIdx.x; c < total; c += gridDim.x) {
        const int m = (int)(c / cps);
        if (!rk.accepted_now[m]) continue;
        const int lo = rk.out_lo[m];
        const int hi = t_eval ? rk.out_hi[m] : lo + 1;
        if (hi <= lo) continue;
        const int i = (int)(c - (long long)m * cps) * OCCM_BLOCK + threadIdx.x;
        if (i >= n_save) continue;
        const long long e = save_idx ? save_idx[i] : i;
        const long long base = (long long)m * ndof;
        const int par = rk.parity[m];               // already flipped: Y[par] = y_new, Y[par^1] = y_old
        const double t = rk.t[m];
        if (!t_eval) {
            out[((long long)m * n_t_eval + lo) * n_save + i] = rk.Y[par][base + e];
            if (out_t && i == 0) out_t[(long long)m * n_t_eval + lo] = t;
            continue;
        }
        const double t_old = rk.t_old[m];
        const double h = t - t_old;
        const double k1 = rk.KF[par ^ 1][base + e], k3 = rk.K3[base + e], k4 = rk.K4[base + e], k5 = rk.K5[base + e],
                     k6 = rk.K6[base + e], k7 = rk.KF[par][base + e];
        const double y_old = rk.Y[par ^ 1][base + e];
        // Q = K^T P  (rk.py:170)
        const double q0 = k1 * dp::P11;
        const double q1 = k1 * dp::P12 + k3 * dp::P32 + k4 * dp::P42 + k5 * dp::P52 + k6 * dp::P62 + k7 * dp::P72;
        const double q2 = k1 * dp::P13 + k3 * dp::P33 + k4 * dp::P43 + k5 * dp::P53 + k6 * dp::P63 + k7 * dp::P73;
        const double q3 = k1 * dp::P14 + k3 * dp::P34 + k4 * dp::P44 + k5 * dp::P54 + k6 * dp::P64 + k7 * dp::P74;
        for (int j = lo; j < hi; ++j) {
            const double x = (t_eval[j] - t_old) / h;
            const double x2 = x * x, x3 = x2 * x, x4 = x3 * x;      // np.cumprod([x, x, x, x])
            const double yv = h * (q0 * x + q1 * x2 + q2 * x3 + q3 * x4) + y_old;
            out[((long long)m * n_t_eval + j) * n_save + i] = yv;
            if (out_t && i == 0) out_t[(long long)m * n_t_eval + j] = t_eval[j];
        }
    }
}

__global__ void occm_rk45_resume_kernel(const OccmRkDev rk, int M) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    rk.n_rows[m] = 0;
    if (rk.status[m] == OCCM_MEMBER_PAUSED) rk.status[m] = OCCM_MEMBER_RUNNING;
}

__global__ void occm_rk45_gather_state(const OccmRkDev rk, int M, long long ndof, double* __restrict__ out) {
    const long long total = (long long)M * ndof;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int m = (int)(i / ndof);
        out[i] = rk.Y[rk.parity[m]][i];
    }
}

__global__ void occm_rk45_write_initial(const double* __restrict__ y0, int M, long long ndof, int n_save,
                                        const int* __restrict__ save_idx, int cap, double t0, double* __restrict__ out,
                                        double* __restrict__ out_t) {
    const long long total = (long long)M * n_save;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
        const int m = (int)(k / n_save);
        const int i = (int)(k - (long long)m * n_save);
        const long long e = save_idx ? save_idx[i] : i;
        out[((long long)m * cap) * n_save + i] = y0[(long long)m * ndof + e];
        if (i == 0) out_t[(long long)m * cap] = t0;
    }
}

#include "occm_small.cuh"

#endif  // !OCCM_STAGE_TU

// ------------------------------------------------------------------------------------------------ launch helpers
template <int MODEL, int STAGE>
static int occm_launch_stage_m(const occm_system* sys, const OccmMemDev& mem, const OccmRkDev& rk, const double* t_dev,
                               double t_scalar, const double* y_in, double* dy_out, cudaStream_t st) {
    static OccmLaunchCache cache;
    int ctas_per_sm = 1;
    constexpr int E = STAGE <= 1 ? OCCM_E0 : OCCM_E;
    const size_t smem = occm_elem_smem(sys, E);
    auto kern = occm_stage<MODEL, STAGE, E>;
    if (int rc = occm_kernel_prepare(&cache, (const void*)kern, sys->device, OCCM_BLOCK, smem, &ctas_per_sm)) return rc;
    OccmSysDev dev = sys->dev;
    dev.chunks_per_member = (int)((dev.n_points + (long long)E * dev.tc - 1) / ((long long)E * dev.tc));
    const long long chunks = (long long)mem.M * dev.chunks_per_member;
    if (chunks >= (1ll << 31)) { occm_set_error("too many chunks (%lld) for one launch", chunks); return OCCM_ERR_INVALID; }
    const int grid = occm_grid(sys, chunks, ctas_per_sm);   // persistent CTAs: a multiple of the SM count
    kern<<<grid, OCCM_BLOCK, smem, st>>>(dev, mem, rk, t_dev, t_scalar, y_in, dy_out);
    OCCM_LAUNCH_CHECK("occm_stage");
    return OCCM_OK;
}

typedef void (*occm_fast_fn)(const OccmSysDev, const OccmMemDev, const OccmRkDev, const double*, double, const double*, double*,
                             const OccmEllEntry*, int, int, int);
template <int MODEL, int STAGE, int K, int WIDE, int V>
static occm_fast_fn occm_fast_kernel() {
    if constexpr (MODEL == 0) return occm_fast_cstr<STAGE, K, WIDE, V>;
    else return occm_fast_pfr<STAGE, WIDE, V>;
}

// Flagged points of the fast / pipelined path (occm_fixup): launched right after the hot kernel of the same stage
template <int MODEL, int STAGE>
static int occm_launch_fixup(const occm_system* sys, const OccmMemDev& mem, const OccmRkDev& rk, const double* t_dev, double t_scalar,
                             const double* y_in, double* dy_out, cudaStream_t st, int err_stride, int err_offset) {
    static OccmLaunchCache fix_cache;
    const int S = sys->dev.S, n_flagged = sys->desc.n_flagged;
    const int tcf = OCCM_BLOCK / S;
    const int nfix = n_flagged > 0 ? (n_flagged + tcf - 1) / tcf : 0;
    if (nfix == 0) return OCCM_OK;
    const size_t fsmem = (size_t)sys->dev.tables_bytes + sizeof(double) * ((size_t)(MODEL == 1 ? 2 : 1) * tcf * (S + 1) + 8);
    if (int rc = occm_kernel_prepare(&fix_cache, (const void*)occm_fixup<MODEL, STAGE>, sys->device, OCCM_BLOCK, fsmem, nullptr)) return rc;
    occm_fixup<MODEL, STAGE><<<dim3(nfix, mem.M), OCCM_BLOCK, fsmem, st>>>(sys->dev, mem, rk, t_dev, t_scalar, y_in, dy_out, sys->desc.flagged_points,
                                                                      n_flagged, err_stride, err_offset);
    OCCM_LAUNCH_CHECK("occm_fixup");
    return OCCM_OK;
}

// V = values per thread: 2 (even species counts, 128-bit accesses) or 1 (odd species counts, 64-bit accesses)
template <int MODEL, int STAGE, int K, int WIDE, int V>
static int occm_launch_fast(const occm_system* sys, const OccmMemDev& mem, const OccmRkDev& rk, const double* t_dev,
                            double t_scalar, const double* y_in, double* dy_out, cudaStream_t st, int* cpm_out) {
    static OccmLaunchCache cache;
    int ctas_per_sm = 1;
    const int S = sys->dev.S;
    const int H = S / V;                                                                                // threads per point
    const int tcv = (OccmWarpMap<STAGE>::value && !WIDE) ? (OCCM_FAST_BLOCK / 32) * (32 / H)            // a warp owns 32 / H whole points
                                                          : OCCM_FAST_BLOCK / H;
    const size_t nw = (size_t)S * sys->max_terms;          // WIDE: unscaled + two scaled coefficient copies + meta words
    const size_t smem = (((size_t)sys->dev.tables_bytes + 15) & ~(size_t)15) +
                        sizeof(double) * ((size_t)2 * (OCCM_TILE_BUF_BYTES / 8) + (WIDE ? 3 * nw + (nw + 1) / 2 : (size_t)S * OCCM_NT) + 40) +
                        (OccmOpsAsync<STAGE>::value ? (size_t)2 * OccmNops<STAGE>::value * OCCM_OPS_STRIDE : 0);
    auto kern = occm_fast_kernel<MODEL, STAGE, K, WIDE, V>();
    if (int rc = occm_kernel_prepare(&cache, (const void*)kern, sys->device, OCCM_FAST_BLOCK, smem, &ctas_per_sm)) return rc;
    const int cpm = (int)((sys->dev.n_points + tcv - 1) / tcv);
    const long long chunks = (long long)mem.M * cpm;
    if (chunks >= (1ll << 31)) { occm_set_error("too many chunks (%lld) for one launch", chunks); return OCCM_ERR_INVALID; }
    const int grid = occm_grid(sys, chunks, ctas_per_sm);
    const int n_flagged = sys->desc.n_flagged;
    const int tcf = OCCM_BLOCK / S;
    const int nfix = n_flagged > 0 ? (n_flagged + tcf - 1) / tcf : 0;
    const int nwarp = OCCM_FAST_BLOCK / 32;
    const int err_stride = cpm * nwarp + nfix;            // partial sums per member: one per warp and chunk, then fix-up blocks
    if (STAGE == 7 && (size_t)err_stride > occm_err_capacity(sys)) {
        occm_set_error("error partials per member (%d) exceed the carved capacity (%zu)", err_stride, occm_err_capacity(sys));
        return OCCM_ERR_WORKSPACE;
    }
    kern<<<grid, OCCM_FAST_BLOCK, smem, st>>>(sys->dev, mem, rk, t_dev, t_scalar, y_in, dy_out, (const OccmEllEntry*)sys->desc.ell_packed, cpm, tcv, err_stride);
    OCCM_LAUNCH_CHECK("occm_fast_cstr");
    if (int rc = occm_launch_fixup<MODEL, STAGE>(sys, mem, rk, t_dev, t_scalar, y_in, dy_out, st, err_stride, cpm * nwarp)) return rc;
    if (cpm_out) *cpm_out = err_stride;
    return OCCM_OK;
}

static inline bool occm_aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

// TMA-pipelined kernels (occm_pipe.cuh).  Returns 1 when the configuration does not fit them (the caller then takes the
// register-prefetching fast kernels), 0 after a launch, < 0 on error.
template <int MODEL, int STAGE, int K, int WIDE, int V>
static int occm_launch_pipe(const occm_system* sys, const OccmMemDev& mem, const OccmRkDev& rk, const double* t_dev,
                            double t_scalar, const double* y_in, double* dy_out, cudaStream_t st, int* cpm_out) {
    static OccmLaunchCache cache;
    constexpr int NS = OccmPipeStreams<STAGE>::value;
    constexpr int CTAS = OccmPipeLight<STAGE>::value ? OCCM_PIPE_CTAS_LIGHT : OCCM_PIPE_CTAS;
    const int S = sys->dev.S, H = S / V, ppw = 32 / H;
    const long long ndof = sys->dev.ndof;
    // bulk copies move multiples of 16 bytes from 16-byte aligned addresses: every member base and every chunk must start on one
    if ((ndof & 1) || ((sys->dev.n_points * S) & 1)) return 1;
    if (STAGE <= 1 && (!occm_aligned16(y_in) || !occm_aligned16(dy_out) || (STAGE == 1 && !occm_aligned16(mem.z)))) return 1;
    constexpr int E = OccmPipeE<STAGE>::value;
    OccmPipeGeom g;
    g.psub = OCCM_PIPE_WARPS * ppw;
    g.pc = E * g.psub;
    g.cpm = (int)((sys->dev.n_points + g.pc - 1) / g.pc);
    g.ell_stride = g.pc * 16;
    g.slot_bytes = (int)((NS * OccmPipeStreamBytes<STAGE>::value + (unsigned)(K * g.ell_stride) + 127u) & ~127u);
    const size_t nw = (size_t)S * sys->max_terms;
    size_t o = ((size_t)sys->dev.tables_bytes + 15) & ~(size_t)15;
    g.coef_off = (int)o;
    g.wide_stride = (int)(8 * nw);
    o += WIDE ? 8 * nw + 4 * (nw + (nw & 1)) + (size_t)OCCM_PIPE_WARPS * g.wide_stride : sizeof(double) * (size_t)S * OCCM_NT;
    o = (o + 15) & ~(size_t)15;
    g.tile_off = (int)o;
    o += sizeof(double) * (size_t)E * OCCM_PIPE_WARPS * ppw * (S + 2);
    o = (o + 15) & ~(size_t)15;
    g.bar_off = (int)o;
    o += 16 * OCCM_PIPE_MAX_DEPTH;
    o = (o + 127) & ~(size_t)127;
    g.ring_off = (int)o;
    // 1 KB per resident CTA is reserved by the system.  The resident CTAs together stay inside the 196 KB shared-memory
    // configuration: the next one (228 KB) leaves 28 KB of L1 for the row gathers and measured 18 % slower (stage 5, depth 4 vs 3)
    const size_t smem_sm = (size_t)sys->smem_per_sm < 196 * 1024 ? (size_t)sys->smem_per_sm : 196 * 1024;
    const size_t budget = smem_sm / CTAS - 1024;
    const size_t cap = (size_t)sys->smem_per_block < budget ? (size_t)sys->smem_per_block : budget;
    if (cap < o + 2 * (size_t)g.slot_bytes) return 1;
    int depth = (int)((cap - o) / (size_t)g.slot_bytes);
    // Ring depth, measured per stage on C4 at 512 members (profiles/r02_notes.md): the operand-heavy stages want the ring as deep
    // as two resident CTAs allow, stage 2 and the J*v kernel (two summands per gathered row) are best at 4, stage 7 at 5.
    int want = sys->pipe_depth > 0 ? sys->pipe_depth                         // OCCM_B200_PIPE_DEPTH overrides
             : STAGE == 0 ? 8 : (STAGE == 1 || STAGE == 2) ? 4 : STAGE == 7 ? 5 : 6;
    if (want < 2) want = 2;
    if (want > OCCM_PIPE_MAX_DEPTH) want = OCCM_PIPE_MAX_DEPTH;
    if (depth > want) depth = want;
    auto kern = occm_pipe<MODEL, STAGE, K, WIDE, V>;
    int ctas_per_sm = 1;
    {   // never trade a resident CTA for a deeper ring (measured: stage 5 at one CTA per SM 6.5 ms, at two 5.5 ms)
        const unsigned long long key = ((unsigned long long)o << 32) ^ ((unsigned long long)g.slot_bytes << 8) ^ (unsigned long long)depth;
        const int d = sys->device >= 0 && sys->device < OCCM_MAX_DEVICES ? sys->device : 0;
        if (cache.geom_key[d] == key && cache.geom_depth[d] > 0) depth = cache.geom_depth[d];
        else {
            int dd = depth;
            for (; dd >= 2; --dd) {
                if (int rc = occm_kernel_prepare(&cache, (const void*)kern, sys->device, OCCM_PIPE_THREADS, o + (size_t)dd * g.slot_bytes, &ctas_per_sm)) return rc;
                if (ctas_per_sm >= CTAS) break;
            }
            if (dd < 2) dd = 2;
            cache.geom_key[d] = key; cache.geom_depth[d] = dd;
            depth = dd;
        }
    }
    g.depth = depth;
    const size_t smem = o + (size_t)depth * g.slot_bytes;
    const long long chunks = (long long)mem.M * g.cpm;
    if (chunks >= (1ll << 31)) { occm_set_error("too many chunks (%lld) for one launch", chunks); return OCCM_ERR_INVALID; }
    const int n_flagged = sys->desc.n_flagged;
    const int tcf = OCCM_BLOCK / S;
    const int nfix = n_flagged > 0 ? (n_flagged + tcf - 1) / tcf : 0;
    g.err_stride = g.cpm * OCCM_PIPE_WARPS + nfix;
    if (STAGE == 7 && (size_t)g.err_stride > occm_err_capacity(sys)) {
        occm_set_error("error partials per member (%d) exceed the carved capacity (%zu)", g.err_stride, occm_err_capacity(sys));
        return OCCM_ERR_WORKSPACE;
    }
    if (int rc = occm_kernel_prepare(&cache, (const void*)kern, sys->device, OCCM_PIPE_THREADS, smem, &ctas_per_sm)) return rc;
    const int grid = occm_grid(sys, chunks, ctas_per_sm);
    kern<<<grid, OCCM_PIPE_THREADS, smem, st>>>(sys->dev, mem, rk, y_in, dy_out, (const OccmEllEntry*)sys->desc.ell_packed, g);
    OCCM_LAUNCH_CHECK("occm_pipe");
    if (int rc = occm_launch_fixup<MODEL, STAGE>(sys, mem, rk, t_dev, t_scalar, y_in, dy_out, st, g.err_stride, g.cpm * OCCM_PIPE_WARPS)) return rc;
    if (cpm_out) *cpm_out = g.err_stride;
    return OCCM_OK;
}

// Mechanism-specialised kernels of the light stages (occm_spec.h; image loaded with occm_system_load_specialized), then the
// generic fixup of the flagged rows.  Returns 1 when there is no image or it does not apply.
template <int STAGE>
static int occm_launch_spec(const occm_system* sys, const OccmMemDev& mem, const OccmRkDev& rk, const double* t_dev,
                            double t_scalar, const double* y_in, double* dy_out, cudaStream_t st, int* cpm_out) {
    if constexpr (STAGE == 0 || STAGE == 1 || STAGE == 2 || STAGE == 7) {
        if (!sys->spec || !sys->use_row) return 1;
        int err_stride = 0, main_partials = 0;
        const int rc = occm_spec_launch(sys, STAGE, mem, rk, y_in, dy_out, st, &err_stride, &main_partials);
        if (rc != OCCM_OK) return rc;
        const int rf = sys->dev.model == OCCM_MODEL_PFR
            ? occm_launch_fixup<1, STAGE>(sys, mem, rk, t_dev, t_scalar, y_in, dy_out, st, err_stride, main_partials)
            : occm_launch_fixup<0, STAGE>(sys, mem, rk, t_dev, t_scalar, y_in, dy_out, st, err_stride, main_partials);
        if (rf) return rf;
        if (cpm_out) *cpm_out = err_stride;
        return OCCM_OK;
    }
    return 1;
}

// the pipelined kernel when it applies, else the register-prefetching one
template <int MODEL, int STAGE, int K, int WIDE, int V>
static int occm_launch_both(const occm_system* sys, const OccmMemDev& mem, const OccmRkDev& rk, const double* t_dev,
                            double t_scalar, const double* y_in, double* dy_out, cudaStream_t st, int* cpm_out) {
    if (sys->use_pipe) {
        const int rc = occm_launch_pipe<MODEL, STAGE, K, WIDE, V>(sys, mem, rk, t_dev, t_scalar, y_in, dy_out, st, cpm_out);
        if (rc <= 0) return rc;
    }
    return occm_launch_fast<MODEL, STAGE, K, WIDE, V>(sys, mem, rk, t_dev, t_scalar, y_in, dy_out, st, cpm_out);
}

// Every STAGE pulls in ~12 kernel instantiations; the build compiles this file once per stage with -DOCCM_STAGE_TU=<stage>
// (explicit instantiation below) in parallel, and once without it for everything else (extern template).
template <int STAGE>
int occm_launch_stage(const occm_system* sys, const OccmMemDev& mem, const OccmRkDev& rk, const double* t_dev,
                      double t_scalar, const double* y_in, double* dy_out, cudaStream_t st, int* cpm_out = nullptr) {
    const OccmSysDev& d = sys->dev;
    const bool wide = !d.fast_ok || sys->force_wide;      // term lists in shared memory instead of registers
    const bool even = d.S % 2 == 0;                        // two species per thread (128-bit) or one (64-bit)
    const bool fast = sys->use_fast && (d.fast_ok || d.wide_ok) && d.S / (even ? 2 : 1) <= 32 && sys->desc.ell_packed &&
                      (sys->desc.n_flagged == 0 || sys->desc.flagged_points) && mem.M <= 65535 && d.ell_k >= 1 &&
                      d.ell_k <= 4 && d.n_points * d.S < (1ll << 31) &&
                      (!even || (occm_aligned16(y_in) && occm_aligned16(dy_out) && occm_aligned16(mem.z)));
#define OCCM_FAST_CALL(MODEL, K, WIDE, V) occm_launch_both<MODEL, STAGE, K, WIDE, V>(sys, mem, rk, t_dev, t_scalar, y_in, dy_out, st, cpm_out)
#define OCCM_FAST_CALL_V(MODEL, K, WIDE) (even ? OCCM_FAST_CALL(MODEL, K, WIDE, 2) : OCCM_FAST_CALL(MODEL, K, WIDE, 1))
    if (fast && d.model == OCCM_MODEL_PFR && d.ell_k == 1 && sys->desc.n_flagged >= d.n_models) {
        const int rc = occm_launch_spec<STAGE>(sys, mem, rk, t_dev, t_scalar, y_in, dy_out, st, cpm_out);
        if (rc <= 0) return rc;
        return wide ? OCCM_FAST_CALL_V(1, 1, 1) : OCCM_FAST_CALL_V(1, 1, 0);
    }
    if (fast && d.model == OCCM_MODEL_CSTR) {
        {
            const int rc = occm_launch_spec<STAGE>(sys, mem, rk, t_dev, t_scalar, y_in, dy_out, st, cpm_out);
            if (rc <= 0) return rc;
        }
        if (wide) switch (d.ell_k) {
            case 1: return OCCM_FAST_CALL_V(0, 1, 1);
            case 2: return OCCM_FAST_CALL_V(0, 2, 1);
            case 3: return OCCM_FAST_CALL_V(0, 3, 1);
            default: return OCCM_FAST_CALL_V(0, 4, 1);
        }
        switch (d.ell_k) {
            case 1: return OCCM_FAST_CALL_V(0, 1, 0);
            case 2: return OCCM_FAST_CALL_V(0, 2, 0);
            case 3: return OCCM_FAST_CALL_V(0, 3, 0);
            default: return OCCM_FAST_CALL_V(0, 4, 0);
        }
    }
#undef OCCM_FAST_CALL_V
#undef OCCM_FAST_CALL
    if (cpm_out) *cpm_out = d.chunks_per_member;
    return d.model == OCCM_MODEL_CSTR ? occm_launch_stage_m<0, STAGE>(sys, mem, rk, t_dev, t_scalar, y_in, dy_out, st)
                                      : occm_launch_stage_m<1, STAGE>(sys, mem, rk, t_dev, t_scalar, y_in, dy_out, st);
}

#define OCCM_STAGE_SIG(N) int occm_launch_stage<N>(const occm_system*, const OccmMemDev&, const OccmRkDev&, const double*, double, \
                                                   const double*, double*, cudaStream_t, int*)
#ifdef OCCM_STAGE_TU
template OCCM_STAGE_SIG(OCCM_STAGE_TU);
#else
extern template OCCM_STAGE_SIG(0); extern template OCCM_STAGE_SIG(1); extern template OCCM_STAGE_SIG(2); extern template OCCM_STAGE_SIG(3);
extern template OCCM_STAGE_SIG(4); extern template OCCM_STAGE_SIG(5); extern template OCCM_STAGE_SIG(6); extern template OCCM_STAGE_SIG(7);

// used by occm_bdf.cu
int occm_launch_rhs(const occm_system* sys, const OccmMemDev& mem, const double* t_dev, double t_scalar, const double* y,
                    double* dy, cudaStream_t st) {
    OccmRkDev none;
    memset(&none, 0, sizeof(none));
    return mem.z ? occm_launch_stage<1>(sys, mem, none, t_dev, t_scalar, y, dy, st)
                 : occm_launch_stage<0>(sys, mem, none, t_dev, t_scalar, y, dy, st);
}

extern "C" int occm_rhs(const occm_system* sys, const occm_members* mem, const double* t_dev, double t_scalar,
                        const double* y, double* dy, void* stream) {
    if (!sys || !y || !dy) { occm_set_error("occm_rhs: null argument"); return OCCM_ERR_INVALID; }
    OccmRkDev none;
    memset(&none, 0, sizeof(none));
    return occm_launch_stage<0>(sys, occm_mem_dev(mem, 1), none, t_dev, t_scalar, y, dy, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------ RK45 host side
static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct RkLayout {
    size_t vec, off_vec[11], off_t, off_h, off_habs, off_tnew, off_told, off_par, off_status, off_rej, off_acc_now,
        off_lo, off_hi, off_rows, off_nacc, off_nrej, off_nfev, off_err, off_running, total;
};

static RkLayout rk_layout(const occm_system* sys, int M) {
    RkLayout L;
    size_t o = 0;
    L.vec = align256((size_t)M * sys->dev.ndof * sizeof(double));
    for (int i = 0; i < 11; ++i) { L.off_vec[i] = o; o += L.vec; }
    auto take = [&](size_t bytes) { size_t r = o; o += align256(bytes); return r; };
    L.off_t = take(M * 8); L.off_h = take(M * 8); L.off_habs = take(M * 8); L.off_tnew = take(M * 8); L.off_told = take(M * 8);
    L.off_par = take(M * 4); L.off_status = take(M * 4); L.off_rej = take(M * 4); L.off_acc_now = take(M * 4);
    L.off_lo = take(M * 4); L.off_hi = take(M * 4); L.off_rows = take(M * 4);
    L.off_nacc = take(M * 8); L.off_nrej = take(M * 8); L.off_nfev = take(M * 8);
    L.off_err = take((size_t)M * occm_err_capacity(sys) * 8);        // any chunking of any stage-7 variant
    L.off_running = take(256);
    L.total = o;
    return L;
}

extern "C" size_t occm_rk45_workspace_bytes(const occm_system* sys, int32_t M) {
    if (!sys || M < 1) return 0;
    return rk_layout(sys, M).total;
}

// device side of occm_rk45_init: initial state, member scalars, K1 = f(t0, y0), first output row
static int occm_rk45_init_device(occm_rk45* rk, cudaStream_t st) {
    const occm_system* sys = rk->sys;
    const occm_rk45_problem* p = &rk->prob;
    const OccmRkDev& d = rk->d;
    const int M = rk->M;
    const size_t state_bytes = (size_t)M * sys->dev.ndof * sizeof(double);
    OCCM_CUDA_CHECK(cudaMemcpyAsync(d.Y[0], p->y0, state_bytes, cudaMemcpyDeviceToDevice, st));
    occm_rk45_init_members<<<(M + 127) / 128, 128, 0, st>>>(d, M, p->t0, p->first_step, p->t_eval ? 0 : 1, p->t_eval, p->n_t_eval);
    OCCM_LAUNCH_CHECK("occm_rk45_init_members");
    // K1 = f(t0, y0)   (rk.py:94)
    const OccmMemDev md = occm_mem_dev(rk->has_mem ? &rk->mem : nullptr, 1);
    const int rc = occm_launch_stage<0>(sys, md, d, nullptr, p->t0, d.Y[0], d.KF[0], st);
    if (rc) return rc;
    if (!p->t_eval) {
        // ivp.py:653-655: ts = [t0], ys = [y0]
        const int n_save = p->save_idx ? p->n_save : (int)sys->dev.ndof;
        const long long total = (long long)M * n_save;
        const int grid = (int)((total + OCCM_BLOCK - 1) / OCCM_BLOCK < 65535 ? (total + OCCM_BLOCK - 1) / OCCM_BLOCK : 65535);
        occm_rk45_write_initial<<<grid, OCCM_BLOCK, 0, st>>>(d.Y[0], M, sys->dev.ndof, n_save, p->save_idx, p->n_t_eval, p->t0,
                                                             p->out, p->out_t);
        OCCM_LAUNCH_CHECK("occm_rk45_write_initial");
    }
    return OCCM_OK;
}

extern "C" int occm_rk45_init(occm_rk45** out, const occm_system* sys, const occm_members* mem, const occm_rk45_problem* p,
                              void* workspace, size_t workspace_bytes, void* stream) {
    if (!out || !sys || !p || !workspace || !p->y0 || !p->out) { occm_set_error("occm_rk45_init: null argument"); return OCCM_ERR_INVALID; }
    const int M = mem ? mem->n_members : 1;
    if (M < 1) { occm_set_error("n_members must be >= 1"); return OCCM_ERR_INVALID; }
    if (!(p->t_bound > p->t0)) { occm_set_error("t_bound must be greater than t0 (forward integration only, helper_functions.py:82)"); return OCCM_ERR_INVALID; }
    if (!(p->first_step > 0.0)) { occm_set_error("`first_step` must be positive."); return OCCM_ERR_INVALID; }
    if (p->first_step > fabs(p->t_bound - p->t0)) { occm_set_error("`first_step` exceeds bounds."); return OCCM_ERR_INVALID; }
    if (!(p->rtol > 0.0) || !(p->atol >= 0.0)) { occm_set_error("rtol must be > 0 and atol >= 0"); return OCCM_ERR_INVALID; }
    if (p->n_t_eval < 1) { occm_set_error("n_t_eval (or the row capacity in all-steps mode) must be >= 1"); return OCCM_ERR_INVALID; }
    if (!p->t_eval && p->n_t_eval < 2) { occm_set_error("all-steps mode needs a row capacity >= 2"); return OCCM_ERR_INVALID; }
    if (!p->t_eval && !p->out_t) { occm_set_error("all-steps mode needs out_t"); return OCCM_ERR_INVALID; }
    const RkLayout L = rk_layout(sys, M);
    if (workspace_bytes < L.total) { occm_set_error("workspace too small: %zu < %zu", workspace_bytes, L.total); return OCCM_ERR_WORKSPACE; }
    if ((uintptr_t)workspace & 255) { occm_set_error("workspace must be 256-byte aligned"); return OCCM_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;

    occm_rk45* rk = (occm_rk45*)calloc(1, sizeof(occm_rk45));
    rk->sys = sys; rk->M = M; rk->prob = *p;
    rk->has_mem = mem != nullptr;
    if (mem) rk->mem = *mem; else { memset(&rk->mem, 0, sizeof(rk->mem)); rk->mem.n_members = 1; }
    char* w = (char*)workspace;
    OccmRkDev& d = rk->d;
    double* vecs[11];
    for (int i = 0; i < 11; ++i) vecs[i] = (double*)(w + L.off_vec[i]);
    d.Y[0] = vecs[0]; d.Y[1] = vecs[1]; d.KF[0] = vecs[2]; d.KF[1] = vecs[3];
    d.K2 = vecs[4]; d.K3 = vecs[5]; d.K4 = vecs[6]; d.K5 = vecs[7]; d.K6 = vecs[8]; d.YS[0] = vecs[9]; d.YS[1] = vecs[10];
    d.t = (double*)(w + L.off_t); d.h = (double*)(w + L.off_h); d.h_abs = (double*)(w + L.off_habs);
    d.t_new = (double*)(w + L.off_tnew); d.t_old = (double*)(w + L.off_told);
    d.parity = (int*)(w + L.off_par); d.status = (int*)(w + L.off_status); d.rejected = (int*)(w + L.off_rej);
    d.accepted_now = (int*)(w + L.off_acc_now); d.out_lo = (int*)(w + L.off_lo); d.out_hi = (int*)(w + L.off_hi);
    d.n_rows = (int*)(w + L.off_rows);
    d.n_acc = (long long*)(w + L.off_nacc); d.n_rej = (long long*)(w + L.off_nrej); d.nfev = (long long*)(w + L.off_nfev);
    d.err_partial = (double*)(w + L.off_err);
    d.n_running = (int*)(w + L.off_running);
    d.rtol = p->rtol; d.atol = p->atol; d.t_bound = p->t_bound; d.max_steps = p->max_steps;

    rk->h_running = (int*)occm_pinned_acquire();
    if (!rk->h_running) {
        occm_set_error("host allocation failed"); free(rk); return OCCM_ERR_CUDA;
    }
    const int rc = occm_rk45_init_device(rk, st);
    if (rc) { occm_pinned_release(rk->h_running); free(rk); return rc; }      // no handle, no pinned slot left behind
    rk->launched_steps = 0;
    *out = rk;
    return OCCM_OK;
}

static int occm_rk45_launch_step(occm_rk45* rk, cudaStream_t st) {
    const occm_system* sys = rk->sys;
    const OccmMemDev md = occm_mem_dev(rk->has_mem ? &rk->mem : nullptr, 1);
    const OccmRkDev& d = rk->d;
    int rc;
    if ((rc = occm_launch_stage<2>(sys, md, d, nullptr, 0.0, nullptr, nullptr, st))) return rc;
    if ((rc = occm_launch_stage<3>(sys, md, d, nullptr, 0.0, nullptr, nullptr, st))) return rc;
    if ((rc = occm_launch_stage<4>(sys, md, d, nullptr, 0.0, nullptr, nullptr, st))) return rc;
    if ((rc = occm_launch_stage<5>(sys, md, d, nullptr, 0.0, nullptr, nullptr, st))) return rc;
    if ((rc = occm_launch_stage<6>(sys, md, d, nullptr, 0.0, nullptr, nullptr, st))) return rc;
    int err_cpm = sys->dev.chunks_per_member;
    if ((rc = occm_launch_stage<7>(sys, md, d, nullptr, 0.0, nullptr, nullptr, st, &err_cpm))) return rc;
    occm_rk45_control<<<rk->M, 128, 0, st>>>(d, err_cpm, sys->dev.ndof, rk->prob.t_eval, rk->prob.n_t_eval);
    OCCM_LAUNCH_CHECK("occm_rk45_control");
    const int n_save = rk->prob.save_idx ? rk->prob.n_save : (int)sys->dev.ndof;
    const long long chunks = (long long)rk->M * ((n_save + OCCM_BLOCK - 1) / OCCM_BLOCK);
    occm_rk45_output<<<occm_grid(sys, chunks, 8), OCCM_BLOCK, 0, st>>>(d, rk->M, sys->dev.ndof, n_save, rk->prob.save_idx,
                                                                     rk->prob.t_eval, rk->prob.n_t_eval, rk->prob.out, rk->prob.out_t);
    OCCM_LAUNCH_CHECK("occm_rk45_output");
    return OCCM_OK;
}

#define OCCM_GRAPH_STEPS 8

// Capture OCCM_GRAPH_STEPS step launches into a graph (after one direct step has initialised the per-kernel launch state).
static void occm_rk45_build_graph(occm_rk45* rk, cudaStream_t st) {
    rk->graph_state = -1;
    if (getenv("OCCM_B200_NO_GRAPH")) return;
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return; }
    const long long before = occm_launch_count();
    int rc = OCCM_OK;
    for (int i = 0; i < OCCM_GRAPH_STEPS && rc == OCCM_OK; ++i) rc = occm_rk45_launch_step(rk, st);
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(st, &graph);
    const long long captured = occm_launch_count() - before;
    occm_count_launch(-(int)captured);                       // nothing ran yet: launches are counted when the graph is launched
    if (rc != OCCM_OK || e != cudaSuccess || !graph) { cudaGetLastError(); if (graph) cudaGraphDestroy(graph); return; }
    if (cudaGraphInstantiate(&rk->graph_exec, graph, 0) != cudaSuccess) { cudaGetLastError(); cudaGraphDestroy(graph); return; }
    cudaGraphDestroy(graph);
    rk->launches_per_step = (int)(captured / OCCM_GRAPH_STEPS);
    rk->graph_state = 1;
}

// OCCM_B200_TIMING=1: report host calls of the run loop that take longer than 50 ms (diagnosis of stalls on shared hosts)
#include <chrono>
struct OccmSlowCall {
    const char* what; bool on; std::chrono::steady_clock::time_point t0;
    explicit OccmSlowCall(const char* w) : what(w), on(getenv("OCCM_B200_TIMING") != nullptr), t0(std::chrono::steady_clock::now()) {}
    ~OccmSlowCall() {
        if (!on) return;
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (ms > 50.0) fprintf(stderr, "[occm timing] %s took %.1f ms on the host\n", what, ms);
    }
};

// Networks that fit in shared memory: one persistent CTA per member (occm_small.cuh).  Returns 1 when the problem does not
// qualify (all-steps mode, too many DOF, tables too large), 0 after running, < 0 on error.
static int occm_rk45_run_small(occm_rk45* rk, int64_t max_launch_steps, int32_t* n_running, cudaStream_t st) {
    static OccmLaunchCache cache[2];
    const occm_system* sys = rk->sys;
    const long long ndof = sys->dev.ndof;
    if (!sys->use_small || !rk->prob.t_eval || ndof > OCCM_SMALL_MAX_NDOF || rk->M > 65535) return 1;
    const int nm = sys->dev.n_models, nnz = sys->nnz, pnnz = sys->pulse_nnz;
    const size_t smem = (size_t)sys->dev.tables_bytes + sizeof(double) * (11 * (size_t)ndof + nnz + nm + pnnz) +
                        sizeof(int) * (2 * (size_t)(nm + 1) + nnz + pnnz) + 4 * (size_t)ndof + 16;      // + element plan (2 x u16 per unknown)
    if (smem + 8192 > (size_t)sys->smem_per_block) return 1;        // + the kernel's static shared memory (controller, boundary-value table)
    int block = (int)((ndof + 31) / 32 * 32);
    block = block < 64 ? 64 : block > 512 ? 512 : block;          // (1024 threads = 64 registers: the generic RHS spilled)
    const bool pfr = sys->dev.model == OCCM_MODEL_PFR;
    const void* kern = pfr ? (const void*)occm_rk45_small<1> : (const void*)occm_rk45_small<0>;
    if (int rc = occm_kernel_prepare(&cache[pfr ? 1 : 0], kern, sys->device, block, smem, nullptr)) return rc;
    const OccmMemDev md = occm_mem_dev(rk->has_mem ? &rk->mem : nullptr, 1);
    const int n_save = rk->prob.save_idx ? rk->prob.n_save : (int)ndof;
    long long done = 0;
    int running = rk->M;
    while (true) {
        long long todo = 1ll << 18;                          // attempts per launch (a launch of a few seconds at most)
        if (max_launch_steps > 0 && done + todo > max_launch_steps) todo = max_launch_steps - done;
        if (pfr) occm_rk45_small<1><<<rk->M, block, smem, st>>>(sys->dev, md, rk->d, rk->prob.t_eval, rk->prob.n_t_eval, rk->prob.save_idx, n_save,
                                                                rk->prob.out, rk->prob.out_t, todo, nnz, pnnz);
        else     occm_rk45_small<0><<<rk->M, block, smem, st>>>(sys->dev, md, rk->d, rk->prob.t_eval, rk->prob.n_t_eval, rk->prob.save_idx, n_save,
                                                                rk->prob.out, rk->prob.out_t, todo, nnz, pnnz);
        OCCM_LAUNCH_CHECK("occm_rk45_small");
        done += todo;
        occm_count_running<<<1, 256, 0, st>>>(rk->d.status, rk->M, rk->d.n_running);
        OCCM_LAUNCH_CHECK("occm_count_running");
        OCCM_CUDA_CHECK(cudaMemcpyAsync(rk->h_running, rk->d.n_running, sizeof(int), cudaMemcpyDeviceToHost, st));
        OCCM_CUDA_CHECK(cudaStreamSynchronize(st));
        running = *rk->h_running;
        if (running == 0 || (max_launch_steps > 0 && done >= max_launch_steps)) break;
    }
    rk->launched_steps += done;
    if (n_running) *n_running = running;
    return OCCM_OK;
}

extern "C" int occm_rk45_run(occm_rk45* rk, int64_t max_launch_steps, int32_t* n_running, void* stream) {
    if (!rk) { occm_set_error("occm_rk45_run: null handle"); return OCCM_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    {
        const int rc = occm_rk45_run_small(rk, max_launch_steps, n_running, st);
        if (rc <= 0) return rc;
    }
    long long launched = 0;
    int batch = 4;
    int running = rk->M;
    while (true) {
        int todo = batch;
        if (max_launch_steps > 0 && launched + todo > max_launch_steps) todo = (int)(max_launch_steps - launched);
        int i = 0;
        if (rk->graph_state == 0 && rk->launched_steps + launched >= 1) occm_rk45_build_graph(rk, st);
        while (i < todo) {
            if (rk->graph_state == 1 && todo - i >= OCCM_GRAPH_STEPS) {
                OccmSlowCall tm("cudaGraphLaunch");
                OCCM_CUDA_CHECK(cudaGraphLaunch(rk->graph_exec, st));
                occm_count_launch(rk->launches_per_step * OCCM_GRAPH_STEPS);
                i += OCCM_GRAPH_STEPS;
            } else {
                OccmSlowCall tm("occm_rk45_launch_step");
                int rc = occm_rk45_launch_step(rk, st);
                if (rc) return rc;
                ++i;
            }
        }
        launched += todo;
        // capture and instantiate the step graph while the GPU is busy with the first batch (it used to idle meanwhile)
        if (rk->graph_state == 0 && rk->launched_steps + launched >= 1) { OccmSlowCall tm("graph capture + instantiate"); occm_rk45_build_graph(rk, st); }
        occm_count_running<<<1, 256, 0, st>>>(rk->d.status, rk->M, rk->d.n_running);
        OCCM_LAUNCH_CHECK("occm_count_running");
        OCCM_CUDA_CHECK(cudaMemcpyAsync(rk->h_running, rk->d.n_running, sizeof(int), cudaMemcpyDeviceToHost, st));
        OCCM_CUDA_CHECK(cudaStreamSynchronize(st));
        running = *rk->h_running;
        if (running == 0) break;
        if (max_launch_steps > 0 && launched >= max_launch_steps) break;
        if (batch < 64) batch *= 2;
    }
    rk->launched_steps += launched;
    if (n_running) *n_running = running;
    return OCCM_OK;
}

extern "C" int occm_rk45_resume(occm_rk45* rk, void* stream) {
    if (!rk) { occm_set_error("occm_rk45_resume: null handle"); return OCCM_ERR_INVALID; }
    occm_rk45_resume_kernel<<<(rk->M + 127) / 128, 128, 0, (cudaStream_t)stream>>>(rk->d, rk->M);
    OCCM_LAUNCH_CHECK("occm_rk45_resume_kernel");
    return OCCM_OK;
}

extern "C" int occm_rk45_stats(occm_rk45* rk, int32_t* status, int64_t* n_accepted, int64_t* n_rejected, int64_t* nfev,
                               double* t, int32_t* n_rows, void* stream) {
    if (!rk) { occm_set_error("occm_rk45_stats: null handle"); return OCCM_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    const int M = rk->M;
    if (status)     OCCM_CUDA_CHECK(cudaMemcpyAsync(status, rk->d.status, M * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (n_accepted) OCCM_CUDA_CHECK(cudaMemcpyAsync(n_accepted, rk->d.n_acc, M * sizeof(long long), cudaMemcpyDeviceToHost, st));
    if (n_rejected) OCCM_CUDA_CHECK(cudaMemcpyAsync(n_rejected, rk->d.n_rej, M * sizeof(long long), cudaMemcpyDeviceToHost, st));
    if (nfev)       OCCM_CUDA_CHECK(cudaMemcpyAsync(nfev, rk->d.nfev, M * sizeof(long long), cudaMemcpyDeviceToHost, st));
    if (t)          OCCM_CUDA_CHECK(cudaMemcpyAsync(t, rk->d.t, M * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (n_rows) {
        // t_eval mode: rows written so far = out_hi; all-steps mode: n_rows
        OCCM_CUDA_CHECK(cudaMemcpyAsync(n_rows, rk->prob.t_eval ? rk->d.out_hi : rk->d.n_rows, M * sizeof(int), cudaMemcpyDeviceToHost, st));
    }
    OCCM_CUDA_CHECK(cudaStreamSynchronize(st));
    return OCCM_OK;
}

extern "C" int occm_rk45_get_state(occm_rk45* rk, double* y_out, void* stream) {
    if (!rk || !y_out) { occm_set_error("occm_rk45_get_state: null argument"); return OCCM_ERR_INVALID; }
    const long long total = (long long)rk->M * rk->sys->dev.ndof;
    const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    occm_rk45_gather_state<<<grid, 256, 0, (cudaStream_t)stream>>>(rk->d, rk->M, rk->sys->dev.ndof, y_out);
    OCCM_LAUNCH_CHECK("occm_rk45_gather_state");
    return OCCM_OK;
}

extern "C" int occm_rk45_time_stage(occm_rk45* rk, int32_t stage, int32_t reps, float* ms_per_launch, void* stream) {
    if (!rk || !ms_per_launch || reps < 1 || stage < 2 || stage > 7) { occm_set_error("occm_rk45_time_stage: bad argument"); return OCCM_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    const occm_system* sys = rk->sys;
    const OccmMemDev md = occm_mem_dev(rk->has_mem ? &rk->mem : nullptr, 1);
    cudaEvent_t e0, e1;
    OCCM_CUDA_CHECK(cudaEventCreate(&e0));
    OCCM_CUDA_CHECK(cudaEventCreate(&e1));
    OCCM_CUDA_CHECK(cudaEventRecord(e0, st));
    int rc = OCCM_OK;
    for (int i = 0; i < reps && rc == OCCM_OK; ++i) {
        switch (stage) {
            case 2: rc = occm_launch_stage<2>(sys, md, rk->d, nullptr, 0.0, nullptr, nullptr, st); break;
            case 3: rc = occm_launch_stage<3>(sys, md, rk->d, nullptr, 0.0, nullptr, nullptr, st); break;
            case 4: rc = occm_launch_stage<4>(sys, md, rk->d, nullptr, 0.0, nullptr, nullptr, st); break;
            case 5: rc = occm_launch_stage<5>(sys, md, rk->d, nullptr, 0.0, nullptr, nullptr, st); break;
            case 6: rc = occm_launch_stage<6>(sys, md, rk->d, nullptr, 0.0, nullptr, nullptr, st); break;
            default: rc = occm_launch_stage<7>(sys, md, rk->d, nullptr, 0.0, nullptr, nullptr, st); break;
        }
    }
    OCCM_CUDA_CHECK(cudaEventRecord(e1, st));
    OCCM_CUDA_CHECK(cudaEventSynchronize(e1));
    float ms = 0.f;
    OCCM_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *ms_per_launch = ms / reps;
    return rc;
}

extern "C" int occm_rk45_free(occm_rk45* rk) {
    if (!rk) return OCCM_OK;
    occm_pinned_release(rk->h_running);
    if (rk->graph_exec) cudaGraphExecDestroy(rk->graph_exec);
    free(rk);
    return OCCM_OK;
}

#endif  // !OCCM_STAGE_TU
