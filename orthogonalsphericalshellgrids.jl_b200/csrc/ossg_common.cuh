// csrc/ossg_common.cuh -- error plumbing shared by the C-ABI translation units.
#pragma once

#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>

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

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace ossg
