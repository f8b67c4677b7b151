// Host-side internals shared by the translation units of liboccm_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include "../../include/occm_b200.h"
#include "occm_device.cuh"

struct occm_system {
    occm_system_desc desc;
    OccmSysDev dev;
    int sm_count;
    int force_wide;       // OCCM_B200_FORCE_WIDE=1: fast kernels keep the term lists in shared memory even when they fit registers
    int max_terms;        // most terms of any species (table header)
    int use_fast;         // 0: always run the generic kernels (OCCM_B200_NO_FAST=1, used by the tests to cover both)
    size_t smem_tables;   // bytes of the staged table blob (8-byte aligned)
    int device;           // CUDA device the handle was created on (launch state is kept per device)
    int use_pipe;         // 0: no TMA-pipelined kernels (OCCM_B200_NO_PIPE=1)
    int use_row;          // 0: a loaded mechanism-specialised image is ignored (OCCM_B200_NO_SPEC=1)
    struct OccmSpec* spec; // mechanism-specialised kernel image (occm_system_load_specialized), or NULL
    int spec_smem_kb;     // shared memory per SM the specialised kernels plan their rings for (OCCM_B200_SPEC_SMEM_KB, default 196)
    int any_prog;         // the mechanism has non-polynomial terms (byte-code programs)
    int n_term;           // stoichiometric terms of the mechanism (table header)
    int use_small;        // 0: no persistent one-CTA-per-member RK45 for networks that fit in shared memory (OCCM_B200_NO_SMALL=1)
    int pipe_depth;       // cap on the ring depth of the pipelined kernels (OCCM_B200_PIPE_DEPTH, 0 = as deep as shared memory allows)
    int smem_per_block, smem_per_sm;
    int nnz, pulse_nnz;   // entries of row_src / row_w and of pulse_fn / pulse_w
};

// Launch state of ONE kernel, per device: the dynamic shared-memory opt-in that has been granted and the occupancy at the
// last shared-memory size asked for.  cudaFuncSetAttribute applies per device and systems differ in their table sizes, so
// neither may be a process-wide scalar; guarded by one mutex (host threads driving different GPUs).
#define OCCM_MAX_DEVICES 32
struct OccmLaunchCache {
    size_t smem_limit[OCCM_MAX_DEVICES]; size_t occ_smem[OCCM_MAX_DEVICES]; int occ[OCCM_MAX_DEVICES]; int occ_block[OCCM_MAX_DEVICES];
    unsigned long long geom_key[OCCM_MAX_DEVICES]; int geom_depth[OCCM_MAX_DEVICES];      // pipelined kernels: ring depth chosen for a shared-memory layout
};
// opt in to `smem` bytes of dynamic shared memory if needed; *ctas_per_sm (may be NULL) receives the occupancy
int occm_kernel_prepare(OccmLaunchCache* cache, const void* kernel, int device, int block, size_t smem, int* ctas_per_sm);

void occm_set_error(const char* fmt, ...);
void occm_count_launch(int n = 1);
// 64-byte slots of one pinned page that lives as long as the process: cudaHostAlloc / cudaFreeHost per integrator handle
// cost up to hundreds of milliseconds on a busy host (they serialise on driver locks)
void* occm_pinned_acquire(void);
void occm_pinned_release(void* p);
struct occm_system;
struct OccmMemDev;
int occm_launch_rhs(const occm_system* sys, const OccmMemDev& mem, const double* t_dev, double t_scalar, const double* y,
                    double* dy, cudaStream_t st);
struct OccmRkDev;
// specialised image (occm_spec.h): launch of stage 0 / 2 / 7 (returns 1 when it does not apply); unload at system destruction
int occm_spec_launch(const occm_system* sys, int stage, const OccmMemDev& mem, const OccmRkDev& rk, const double* y_in,
                     double* dy_out, cudaStream_t st, int* err_stride_out, int* main_partials_out);
void occm_spec_unload(occm_system* s);

#define OCCM_CUDA_CHECK(expr)                                                                      \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            occm_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return OCCM_ERR_CUDA;                                                                  \
        }                                                                                          \
    } while (0)

#define OCCM_LAUNCH_CHECK(name)                                                                    \
    do {                                                                                           \
        cudaError_t _e = cudaGetLastError();                                                       \
        if (_e != cudaSuccess) {                                                                   \
            occm_set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));               \
            return OCCM_ERR_CUDA;                                                                  \
        }                                                                                          \
        occm_count_launch();                                                                       \
    } while (0)

static inline OccmMemDev occm_mem_dev(const occm_members* mem, int fallback_M) {
    OccmMemDev m;
    m.M = mem ? mem->n_members : fallback_M;
    m.k_scale = mem ? mem->k_scale : nullptr;
    m.q_scale = mem ? mem->q_scale : nullptr;
    m.bc_scale = mem ? mem->bc_scale : nullptr;
    m.active = nullptr; m.z = nullptr; m.eps = nullptr;
    return m;
}

// dynamic shared memory of the element kernels: tables + tile + k_scale row + reduction scratch
static inline size_t occm_elem_smem(const occm_system* sys, int E) {
    const size_t tile = (size_t)E * (sys->dev.tc + 1) * (sys->dev.S + 1);     // [halo row | tc rows] x (S + 1) per sub-chunk
    return sys->smem_tables + sizeof(double) * (tile + sys->dev.n_rxn + (size_t)sys->dev.S * 4 + 44);
}

// Error partial sums per member that the stage-7 kernels may write (the workspace is carved for this many; every launcher
// checks its own count against it).  One partial per warp and chunk; the smallest chunk any variant uses is the warp mapping
// with one species per thread: (OCCM_FAST_BLOCK / 32) * (32 / S) points (odd S; the dense mapping holds OCCM_BLOCK / S >= that).
static inline size_t occm_err_capacity(const occm_system* sys) {
    const int S = sys->dev.S, nwarp = OCCM_FAST_BLOCK / 32;
    long long min_pts = OCCM_FAST_BLOCK / S < OCCM_BLOCK / S ? OCCM_FAST_BLOCK / S : OCCM_BLOCK / S;
    if (S <= 32 && (long long)nwarp * (32 / S) < min_pts) min_pts = (long long)nwarp * (32 / S);
    if (min_pts < 1) min_pts = 1;
    const long long chunks = (sys->dev.n_points + min_pts - 1) / min_pts + 2;
    return (size_t)(chunks * nwarp + sys->desc.n_flagged + 8);
}

// persistent grid: a multiple of the SM count, capped by the amount of work
static inline int occm_grid(const occm_system* sys, long long n_chunks, int ctas_per_sm) {
    long long g = (long long)sys->sm_count * ctas_per_sm;
    if (g > n_chunks) g = n_chunks;
    if (g < 1) g = 1;
    return (int)g;
}
