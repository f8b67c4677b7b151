// One mechanism-specialised kernel image (see ../occm_spec.h): compiled to a cubin by openccm_b200/specialize.py with
//   nvcc -cubin -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -DOCCM_SPEC_MODEL=<0 CSTR | 1 PFR> -DOCCM_SPEC_S=<species> -DOCCM_SPEC_K=<ELL width>
//        -DOCCM_SPEC_HASH=<table hash>ull -include <generated mechanism header> spec/occm_spec.cu
// The generated header defines OCCM_SPEC_NTERM, occm_spec_coef[], occm_spec_rxn[] and OCCM_SPEC_TERMS(c, sc, r).
// Not part of liboccm_b200.so: the library loads the image through occm_system_load_specialized().
#ifndef OCCM_SPEC_MODEL
#define OCCM_SPEC_MODEL 0
#endif
#include "../../../include/occm_b200.h"
#include "../occm_device.cuh"
#include "../occm_rk.cuh"
#include "../occm_fast.cuh"
#include "../occm_pipe.cuh"
#include "../occm_row.cuh"

#if !defined(OCCM_SPEC_S) || !defined(OCCM_SPEC_K) || !defined(OCCM_SPEC_NTERM) || !defined(OCCM_SPEC_HASH)
#error "occm_spec.cu is compiled per mechanism: OCCM_SPEC_S, OCCM_SPEC_K, OCCM_SPEC_HASH and the generated mechanism header are required"
#endif

extern "C" __device__ unsigned long long occm_spec_info[OCCM_SI_WORDS] = {
    OCCM_SPEC_ABI, OCCM_SPEC_S, OCCM_SPEC_K, OCCM_SPEC_NTERM, OCCM_SPEC_HASH, OCCM_SPEC_MODEL,
    OCCM_ROW_W0, OCCM_ROW_C0, OCCM_ROW_W1, OCCM_ROW_C1, OCCM_ROW_W2, OCCM_ROW_C2, OCCM_ROW_W7, OCCM_ROW_C7};

#define OCCM_SPEC_KERNEL(STAGE, W, C)                                                                                              \
    extern "C" __global__ void __launch_bounds__(32 * (W + 1), C)                                                                  \
    occm_spec_stage##STAGE(const __grid_constant__ OccmSysDev sys, const __grid_constant__ OccmMemDev mem,                         \
                           const __grid_constant__ OccmRkDev rk, const double* __restrict__ y_in, double* __restrict__ dy_out,      \
                           const OccmEllEntry* __restrict__ ell, const __grid_constant__ OccmRowGeom g) {                          \
        occm_row_body<STAGE, W>(sys, mem, rk, y_in, dy_out, ell, g);                                                                \
    }
OCCM_SPEC_KERNEL(0, OCCM_ROW_W0, OCCM_ROW_C0)
OCCM_SPEC_KERNEL(1, OCCM_ROW_W1, OCCM_ROW_C1)
OCCM_SPEC_KERNEL(2, OCCM_ROW_W2, OCCM_ROW_C2)
OCCM_SPEC_KERNEL(7, OCCM_ROW_W7, OCCM_ROW_C7)
