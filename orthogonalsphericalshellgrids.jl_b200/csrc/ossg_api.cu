// csrc/ossg_api.cu -- housekeeping entry points of the C ABI (version, errors, device check).
#include "ossg_common.cuh"

#include <cstdlib>
#include <cstring>

namespace ossg {

char* last_error_buffer()
{
    static thread_local char buf[512] = "";
    return buf;
}

int set_error(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error_buffer(), 512, fmt, ap);
    va_end(ap);
    return code;
}

int env_int(const char* name, int fallback)
{
    const char* e = getenv(name);
    if (!e || !*e) return fallback;
    char* end = nullptr;
    const long v = strtol(e, &end, 10);
    return (end && *end == 0) ? (int)v : fallback;
}

}  // namespace ossg

#ifdef OSSG_TESTUTIL_BUILD
extern "C" const char* ossg_testutil_last_error(void) { return ossg::last_error_buffer(); }
#else
extern "C" int ossg_version(void) { return OSSG_VERSION_MAJOR * 100 + OSSG_VERSION_MINOR; }

extern "C" const char* ossg_last_error(void) { return ossg::last_error_buffer(); }

extern "C" int ossg_device_ok(void)
{
    int dev = -1, major = 0, minor = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return ossg::cuda_fail(e, "cudaGetDevice");
    // (two attribute queries, microseconds; cudaGetDeviceProperties takes milliseconds and sat in every constructor call)
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    if (e != cudaSuccess) return ossg::cuda_fail(e, "cudaDeviceGetAttribute");
    if (major != 10)
        return ossg::set_error(-1, "device %d has compute capability %d.%d; this library is built for sm_100a only", dev, major, minor);
    return 0;
}
#endif  // OSSG_TESTUTIL_BUILD
