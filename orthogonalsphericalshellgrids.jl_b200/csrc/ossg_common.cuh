// csrc/ossg_common.cuh -- error plumbing shared by the C-ABI translation units.
#pragma once

#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstdint>

#include "../../include/ossg_b200.h"

namespace ossg {

char* last_error_buffer();                       // thread-local, defined in ossg_api.cu
int set_error(int code, const char* fmt, ...);   // formats into the buffer, returns code

inline int cuda_fail(cudaError_t e, const char* what)
{
    return set_error((int)e, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}

#define OSSG_CUDA_TRY(expr)                                              \
    do {                                                                 \
        cudaError_t _e = (expr);                                         \
        if (_e != cudaSuccess) return ::ossg::cuda_fail(_e, #expr);      \
    } while (0)

// after a kernel launch: surfaces launch-configuration errors without synchronising
#define OSSG_LAUNCH_CHECK(name)                                          \
    do {                                                                 \
        cudaError_t _e = cudaGetLastError();                             \
        if (_e != cudaSuccess) return ::ossg::cuda_fail(_e, name);       \
    } while (0)

#if defined(__CUDACC__)
// `sign * c[i', j', k]` of the fold (src/zipper_boundary_condition.jl:81,98,115,131): bc.condition is
// an integer +-1 promoted to the field's float type, so the reference executes an IEEE
// multiplication by +-1.0.  That is exact for every number, and for a NaN operand the CPU the
// reference runs on (x86 mulsd/mulss, AArch64 fmul) returns that NaN QUIETED with its own sign and
// payload -- multiplication by -1.0 does not flip a NaN's sign, a negation would.  The GPU's own
// multiply returns a canonical NaN instead, so the NaN case is spelled out on the bits; the fold is
// then bit-exact against the reference's CPU path for every input, NaN payloads included.
// Branch-free on purpose (selects only): the fill kernels issue all their loads first and must stay
// straight-line code for that; a first version with branches cost the zipper kernel 18 %.
//   fold_unit : bc.condition = +-1 (every Zipper the reference creates); flip_hi = 0x80000000 for -1.
//               High word only: one FP64 compare (that pipe is idle in the fill kernels), two logic ops, one select.
//   apply_sign: any bc.condition (selects between the unit form and the oracle's (T)(s * c)).
__device__ __forceinline__ double fold_unit(double v, unsigned flip_hi)
{
    const int hi = __double2hiint(v);
    return __hiloint2double((v != v) ? (hi | 0x00080000) : (int)((unsigned)hi ^ flip_hi), __double2loint(v));
}
__device__ __forceinline__ float fold_unit(float v, unsigned flip_hi)
{
    const int u = __float_as_int(v);
    return __int_as_float((v != v) ? (u | 0x00400000) : (int)((unsigned)u ^ flip_hi));
}
__device__ __forceinline__ bool sign_is_unit(double s) { return s == 1.0 || s == -1.0; }
__device__ __forceinline__ unsigned sign_flip(double s) { return s == -1.0 ? 0x80000000u : 0u; }
template <typename T>
__device__ __forceinline__ T apply_sign(T v, double s)
{
    const T unit = fold_unit(v, sign_flip(s));
    const T general = (T)(s * (double)v);
    return sign_is_unit(s) ? unit : general;
}
// compile-time choice for the single-GPU fill kernels (issue-sensitive): UNIT is decided once per thread
template <bool UNIT, typename T>
__device__ __forceinline__ T fold_value(T v, double s, unsigned flip_hi)
{
    return UNIT ? fold_unit(v, flip_hi) : (T)(s * (double)v);
}

// Copy of one halo-width piece (Hx elements of a row).  The common case -- 4 doubles, both ends
// 32-byte aligned -- is ONE 256-bit load and ONE 256-bit store (LDG.256 / STG.256, sm_100): one
// DRAM sector, and over NVLink one 32-byte packet instead of four 8-byte ones.
template <typename T>
__device__ __forceinline__ void copy_halo_piece(T* dst, const T* src, int Hx)
{
    const uintptr_t both = reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src);
    if (sizeof(T) == 8 && Hx == 4 && (both & 31) == 0) {
        double a, b, c, d;
        asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(src));
        asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(dst), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
        return;
    }
    if (sizeof(T) == 4 && Hx == 4 && (both & 15) == 0) {
        *reinterpret_cast<float4*>(dst) = *reinterpret_cast<const float4*>(src);
        return;
    }
    int q = 0;
    for (; q + 4 <= Hx; q += 4) {
        const T v0 = src[q], v1 = src[q + 1], v2 = src[q + 2], v3 = src[q + 3];
        dst[q] = v0; dst[q + 1] = v1; dst[q + 2] = v2; dst[q + 3] = v3;
    }
    for (; q < Hx; ++q) dst[q] = src[q];
}
#endif

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// integer tuning knob from the environment; callers cache the value in a
// function-local static, so the environment is read once per process and never on a launch path
int env_int(const char* name, int fallback);

}  // namespace ossg
