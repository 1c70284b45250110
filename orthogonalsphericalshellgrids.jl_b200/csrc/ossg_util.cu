// csrc/ossg_util.cu -- libossg_b200_testutil.so (include/ossg_b200_testutil.h): measurement / test
// utilities kept OUT of the product library:
// deterministic synthetic fields and an FP64 FMA rate probe that gives
// the roofline denominator for the coordinate/metric kernels (MEASURED_PEAKS.json has no FP64
// figure).  Nothing here is on the reference's path.
#include "ossg_common.cuh"
#include "ossg_math.cuh"
#include "../../include/ossg_b200_testutil.h"

namespace ossg {

__device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// value in [-1, 1) with 24 random bits, exactly representable in FP32 and FP64
__device__ __forceinline__ double hash_value(uint64_t seed, int gi, int gj, int k)
{
    const uint64_t key = ((uint64_t)(uint32_t)gi) | ((uint64_t)(uint32_t)gj << 21) | ((uint64_t)(uint32_t)k << 42);
    const uint64_t h = splitmix64(seed ^ splitmix64(key));
    return (double)(int64_t)(h >> 40) * (1.0 / 8388608.0) - 1.0;
}

template <typename T> __device__ __forceinline__ T poison();
template <> __device__ __forceinline__ double poison<double>() { return __longlong_as_double(0x7ff8dead0000beefll); }
template <> __device__ __forceinline__ float poison<float>() { return __int_as_float(0x7fc0dead); }

template <typename T>
__global__ void __launch_bounds__(256) synth_kernel(ossg_field_desc f, uint64_t seed, int gi0, int gj0)
{
    const int sx = f.Nx + 2 * f.Hx, sy = f.Ny + 2 * f.Hy;
    const int ix = blockIdx.x * blockDim.x + threadIdx.x;
    const int row = blockIdx.y;                       // row = kz * sy + jy
    if (ix >= sx) return;
    const int kz = row / sy, jy = row - kz * sy;
    const int i = ix - f.Hx + 1, j = jy - f.Hy + 1, k = kz - f.Hz + 1;
    const bool inside = i >= 1 && i <= f.Nx && j >= 1 && j <= f.Ny && k >= 1 && k <= f.Nz;
    T* c = static_cast<T*>(f.data);
    c[((size_t)kz * sy + jy) * sx + ix] = inside ? (T)hash_value(seed, gi0 + i - 1, gj0 + j - 1, k) : poison<T>();
}

__global__ void __launch_bounds__(256) fp64_probe_kernel(int iters, double* sink)
{
    // 8 independent FMA chains per thread: enough ILP to cover the FP64 pipe latency
    double a0 = 1.0 + threadIdx.x * 1e-9, a1 = a0 + 0.125, a2 = a0 + 0.25, a3 = a0 + 0.375;
    double a4 = a0 + 0.5, a5 = a0 + 0.625, a6 = a0 + 0.75, a7 = a0 + 0.875;
    const double m = 0.999999999, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    const double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 12345.678) sink[0] = r;                  // never true; keeps the chains alive
}

}  // namespace ossg

using namespace ossg;

extern "C" int ossg_synth_field(const ossg_field_desc* f, uint64_t seed, int32_t gi0, int32_t gj0, void* stream)
{
    if (!f || !f->data) return set_error(OSSG_EINVAL, "field or its data is NULL");
    if (f->elem_type != OSSG_F64 && f->elem_type != OSSG_F32) return set_error(OSSG_EINVAL, "elem_type must be OSSG_F64 or OSSG_F32");
    const long long sx = f->Nx + 2 * f->Hx, rows = (long long)(f->Ny + 2 * f->Hy) * (f->Nz + 2 * f->Hz);
    if (sx <= 0 || rows <= 0) return set_error(OSSG_EINVAL, "empty field");
    // rows can exceed 65535: loop over chunks of rows by offsetting the data pointer per level
    const int sy = f->Ny + 2 * f->Hy, sz = f->Nz + 2 * f->Hz;
    const int levels_per_launch = 65535 / sy > 0 ? 65535 / sy : 0;
    if (levels_per_launch == 0) return set_error(OSSG_EINVAL, "Ny + 2Hy = %d too large for the synthetic-field kernel", sy);
    for (int kz0 = 0; kz0 < sz; kz0 += levels_per_launch) {
        const int nl = sz - kz0 < levels_per_launch ? sz - kz0 : levels_per_launch;
        ossg_field_desc g = *f;
        // shift the level origin: treat levels [kz0, kz0+nl) as a sub-field whose first level has k = kz0 - Hz + 1
        const size_t esz = f->elem_type == OSSG_F64 ? 8 : 4;
        g.data = static_cast<char*>(f->data) + (size_t)kz0 * sy * sx * esz;
        g.Hz = f->Hz - kz0;            // k = kz - g.Hz + 1 = (kz + kz0) - Hz + 1
        dim3 grid(ceil_div(sx, 256), nl * sy);
        if (f->elem_type == OSSG_F64) synth_kernel<double><<<grid, 256, 0, as_stream(stream)>>>(g, seed, gi0, gj0);
        else synth_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(g, seed, gi0, gj0);
        OSSG_LAUNCH_CHECK("synth_kernel");
    }
    return 0;
}

extern "C" int ossg_fp64_peak_probe(int iters, float* ms, double* flops, void* stream)
{
    if (iters < 1 || !ms || !flops) return set_error(OSSG_EINVAL, "bad arguments");
    int dev = 0, sms = 0;
    OSSG_CUDA_TRY(cudaGetDevice(&dev));
    OSSG_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = sms * 8, threads = 256;
    double* sink = nullptr;
    OSSG_CUDA_TRY(cudaMalloc(&sink, sizeof(double)));
    cudaEvent_t e0, e1;
    OSSG_CUDA_TRY(cudaEventCreate(&e0));
    OSSG_CUDA_TRY(cudaEventCreate(&e1));
    cudaStream_t s = as_stream(stream);
    fp64_probe_kernel<<<blocks, threads, 0, s>>>(iters / 8 + 1, sink);      // warm-up
    OSSG_CUDA_TRY(cudaEventRecord(e0, s));
    fp64_probe_kernel<<<blocks, threads, 0, s>>>(iters, sink);
    OSSG_CUDA_TRY(cudaEventRecord(e1, s));
    OSSG_CUDA_TRY(cudaEventSynchronize(e1));
    OSSG_CUDA_TRY(cudaEventElapsedTime(ms, e0, e1));
    *flops = 2.0 * 8.0 * (double)iters * (double)blocks * (double)threads;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
    OSSG_LAUNCH_CHECK("fp64_probe_kernel");
    return 0;
}

// ------------------------------------------------------------------ sqrt_nb / div_nb self-test
namespace ossg {
__device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// double with a uniformly random mantissa and an exponent uniform in [elo, ehi]
__device__ __forceinline__ double rand_double(uint64_t h, int elo, int ehi)
{
    const uint64_t mant = h & 0x000FFFFFFFFFFFFFull;
    const uint64_t e = (uint64_t)(elo + (int)((h >> 52) % (uint64_t)(ehi - elo + 1)) + 1023);
    return __longlong_as_double((long long)((e << 52) | mant));
}
__global__ void __launch_bounds__(256) selftest_sqrt_div_kernel(uint64_t n, uint64_t seed, unsigned long long* bad)
{
    unsigned long long bs = 0, bd = 0;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t h0 = mix64(seed + 3 * t), h1 = mix64(seed + 3 * t + 1), h2 = mix64(seed + 3 * t + 2);
        // sqrt: alternate the whole admitted range [2^-970, 2^1000] and the kernel's own [2^-60, 2^-11]
        const double a = (t & 1) ? rand_double(h0, -970, 1000) : rand_double(h0, -60, -11);
        if (__double_as_longlong(m::sqrt_nb(a)) != __double_as_longlong(sqrt(a))) ++bs;
        // division: numerators [2^-969, 2^0] (kernel: triple products), denominators [2^-3, 2^2] or [2^-100, 2^100]
        const double nu = (t & 2) ? rand_double(h1, -969, 0) : rand_double(h1, -40, 0);
        double de = (t & 4) ? rand_double(h2, -100, 100) : rand_double(h2, -3, 2);
        if (h2 >> 63) de = -de;
        if (__double_as_longlong(m::div_nb(nu, de)) != __double_as_longlong(nu / de)) ++bd;
    }
    if (bs) atomicAdd(&bad[0], bs);
    if (bd) atomicAdd(&bad[1], bd);
}
}  // namespace ossg

extern "C" int ossg_selftest_sqrt_div(uint64_t n, uint64_t seed, uint64_t mismatches[2], void* stream)
{
    if (!mismatches || n == 0) return set_error(OSSG_EINVAL, "bad arguments");
    unsigned long long* d = nullptr;
    OSSG_CUDA_TRY(cudaMalloc(&d, 2 * sizeof(unsigned long long)));
    cudaStream_t s = as_stream(stream);
    OSSG_CUDA_TRY(cudaMemsetAsync(d, 0, 2 * sizeof(unsigned long long), s));
    ossg::selftest_sqrt_div_kernel<<<148 * 8, 256, 0, s>>>(n, seed, d);
    unsigned long long h[2] = { 0, 0 };
    cudaError_t e = cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(d);
    if (e != cudaSuccess) return cuda_fail(e, "ossg_selftest_sqrt_div");
    mismatches[0] = h[0];
    mismatches[1] = h[1];
    return 0;
}

// ------------------------------------------------------------------ NVLink store-path probe
// Kernel-issued copy of `rows` row segments (seg_bytes each, contiguous source) to a destination
// that is contiguous or laid out like fold rows: row r lands at (r / group) * group_stride +
// (r % group) * pitch.  Persistent blocks stride over (row, chunk) items.  Modes:
//   0  one 8-byte load + store per thread (what the push form of the fold does)
//   1  one 16-byte load + store per thread
//   2  chunk loaded by the threads into shared memory, ONE bulk store (cp.async.bulk S2G) per chunk
//   3  bulk load (G2S on an mbarrier) + bulk store: no SM load/store instruction touches the data
namespace ossg {
struct ProbeArgs {
    char* dst; const char* src;
    unsigned long long rows, group, group_stride, pitch;
    unsigned seg_bytes, chunk_bytes, chunks_per_row;
    int mode;
};
constexpr int PROBE_MAX_CHUNK = 16384;

__global__ void __launch_bounds__(256, 4) peer_copy_probe_kernel(const ProbeArgs a)
{
    __shared__ __align__(128) unsigned char sbuf[PROBE_MAX_CHUNK];
    __shared__ __align__(8) unsigned long long bar;
    const int t = threadIdx.x;
    const unsigned long long nitems = a.rows * a.chunks_per_row;
    unsigned phase = 0;
    if (a.mode == 3) {
        if (t == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"((unsigned)__cvta_generic_to_shared(&bar)) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    for (unsigned long long item = blockIdx.x; item < nitems; item += gridDim.x) {
        const unsigned long long r = item / a.chunks_per_row;
        const unsigned c = (unsigned)(item - r * a.chunks_per_row);
        const unsigned off = c * a.chunk_bytes;
        const unsigned n = min(a.chunk_bytes, a.seg_bytes - off);
        const char* src = a.src + r * a.seg_bytes + off;
        char* dst = a.dst + (r / a.group) * a.group_stride + (r % a.group) * a.pitch + off;
        if (a.mode == 0) {
            for (unsigned b = t * 8; b < n; b += 256 * 8) *reinterpret_cast<double*>(dst + b) = *reinterpret_cast<const double*>(src + b);
        } else if (a.mode == 1) {
            for (unsigned b = t * 16; b < n; b += 256 * 16) *reinterpret_cast<double2*>(dst + b) = *reinterpret_cast<const double2*>(src + b);
        } else if (a.mode == 2) {
            if (t == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncthreads();
            for (unsigned b = t * 16; b < n; b += 256 * 16) *reinterpret_cast<double2*>(sbuf + b) = *reinterpret_cast<const double2*>(src + b);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (t == 0) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                             :: "l"(dst), "r"((unsigned)__cvta_generic_to_shared(sbuf)), "r"(n) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else {
            if (t == 0) {
                const unsigned ba = (unsigned)__cvta_generic_to_shared(&bar);
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(ba), "r"(n) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"((unsigned)__cvta_generic_to_shared(sbuf)), "l"(src), "r"(n), "r"(ba) : "memory");
                unsigned done = 0;
                const long long t0 = clock64();
                while (!done) {
                    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                                 : "=r"(done) : "r"(ba), "r"(phase) : "memory");
                    if (!done && clock64() - t0 > 2000000000ll) __trap();
                }
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                             :: "l"(dst), "r"((unsigned)__cvta_generic_to_shared(sbuf)), "r"(n) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            phase ^= 1;
        }
    }
    if (a.mode >= 2 && t == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
}  // namespace ossg

extern "C" int ossg_peer_copy_probe(void* dst, const void* src, uint64_t rows, uint32_t seg_bytes, uint64_t group,
                                    uint64_t group_stride, uint64_t pitch, uint32_t chunk_bytes, int mode, int blocks, void* stream)
{
    if (!dst || !src || rows == 0 || group == 0) return set_error(OSSG_EINVAL, "bad arguments");
    if (mode < 0 || mode > 3) return set_error(OSSG_EINVAL, "mode must be 0..3");
    if (seg_bytes == 0 || seg_bytes % 16 || pitch % 16 || group_stride % 16 || chunk_bytes == 0 || chunk_bytes % 16 || chunk_bytes > PROBE_MAX_CHUNK)
        return set_error(OSSG_EALIGN, "segment, pitch and chunk sizes must be multiples of 16 bytes (chunk <= %d)", PROBE_MAX_CHUNK);
    if ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) return set_error(OSSG_EALIGN, "pointers must be 16-byte aligned");
    ProbeArgs a;
    a.dst = static_cast<char*>(dst); a.src = static_cast<const char*>(src);
    a.rows = rows; a.group = group; a.group_stride = group_stride; a.pitch = pitch;
    a.seg_bytes = seg_bytes; a.chunk_bytes = chunk_bytes; a.chunks_per_row = (seg_bytes + chunk_bytes - 1) / chunk_bytes;
    a.mode = mode;
    if (blocks < 1) blocks = 148 * 4;
    peer_copy_probe_kernel<<<blocks, 256, 0, as_stream(stream)>>>(a);
    OSSG_LAUNCH_CHECK("peer_copy_probe_kernel");
    return 0;
}
