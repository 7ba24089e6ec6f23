// device_common.cuh -- integer-cast helpers shared by every kernel header of libfftlines_cuda.
#pragma once
#include <stdint.h>
#include <limits.h>
#include "fft_device.cuh"
#include "../../include/fftlines.h"

namespace fftl {

// The reference's (int) casts compile to cvttss2si / cvttsd2si, which return
// INT32_MIN ("integer indefinite") for NaN and out-of-range inputs.
__device__ __forceinline__ int trunc_i32(float x, int strict)
{
    if (strict && (x != x || x >= 2147483648.0f)) return INT_MIN;
    return __float2int_rz(x); // saturating, NaN -> 0
}
__device__ __forceinline__ int trunc_i32(double x, int strict)
{
    if (strict && (x != x || x >= 2147483648.0)) return INT_MIN;
    return __double2int_rz(x);
}
__device__ __forceinline__ int wrap_add(int a, int b) { return (int)((unsigned)a + (unsigned)b); }
__device__ __forceinline__ int wrap_sub(int a, int b) { return (int)((unsigned)a - (unsigned)b); }

} // namespace fftl
