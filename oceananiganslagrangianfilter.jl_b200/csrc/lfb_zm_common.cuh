// lfb_zm_common.cuh -- device helpers of the z-marching stage kernel (lfb_ztma.cu):
// 32-bit shared-window loads/stores, the block barrier and the branch-free division.
#pragma once

#include <cuda_runtime.h>

#include "lfb_internal.h"
#include "lfb_math.cuh"

// ---- shared memory through 32-bit shared-window addresses -----------------------------------------
__device__ __forceinline__ double lds(unsigned a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }

__device__ __forceinline__ void block_barrier() { asm volatile("bar.sync 0;" ::: "memory"); }

// num / den for den > 0 in the normal range: reciprocal seed + two Newton steps (a couple of ulp, far inside
// the 1e-12 parity budget; no special-case branch, unlike the IEEE division sequence)
__device__ __forceinline__ double div_pos(double num, double den)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));
    double e = fma(-den, r, 1.0);
    r = fma(r, e, r);
    e = fma(-den, r, 1.0);
    r = fma(r, e, r);
    return num * r;
}
