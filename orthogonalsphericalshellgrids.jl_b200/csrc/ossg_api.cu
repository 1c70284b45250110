// csrc/ossg_api.cu -- housekeeping entry points of the C ABI (version, errors, device check).
#include "ossg_common.cuh"

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

}  // namespace ossg

extern "C" int ossg_version(void) { return OSSG_VERSION_MAJOR * 100 + OSSG_VERSION_MINOR; }

extern "C" const char* ossg_last_error(void) { return ossg::last_error_buffer(); }

extern "C" int ossg_device_ok(void)
{
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return ossg::cuda_fail(e, "cudaGetDevice");
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) return ossg::cuda_fail(e, "cudaGetDeviceProperties");
    if (prop.major != 10)
        return ossg::set_error(-1, "device %d (%s) has compute capability %d.%d; this library is built for sm_100a only",
                               dev, prop.name, prop.major, prop.minor);
    return 0;
}
