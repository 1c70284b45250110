// csrc/ossg_zipper.cu -- zipper (north-fold) halo fill on sm_100a.
//
// Reference: src/zipper_boundary_condition.jl:70-155.  There, Oceananigans launches one thread
// per (i, k) which loops j = 1..Hy.  Here one thread moves one element: the destination row
// segment is contiguous in i and the source segment is the same span read in reverse, so a warp
// reads one contiguous 32-element span and writes one (a warp-cooperative gather across the
// fold); rows x levels x fields are spread over blockIdx.y/z so a batch of fields is ONE launch.
// The operation is a copy with a sign: bit-exact by construction.
#include "ossg_common.cuh"

namespace ossg {

constexpr int MAX_BATCH = 24;          // ossg_field_desc is 64 B; 24 of them fit kernel parameter space comfortably
constexpr int ZB = 256;                // threads per block along i

struct FieldBatch {
    ossg_field_desc f[MAX_BATCH];
    int n;
};

// rows handled per level: Hy halo rows (+1 for y-Center: the redundant half of row Ny)
__device__ __host__ inline int fold_rows(const ossg_field_desc& f) { return f.Hy + (f.locy == OSSG_CENTER ? 1 : 0); }

// One thread serves one (i, level): it issues the loads of all Hy source rows (in chunks of 4)
// before any store, so that every warp keeps 4 x 256 B in flight -- the fill is a pure
// latency/bandwidth problem (zero flops).
// Source rows are read once and halo rows written once per fill: streaming accesses (ld/st.global.cs, evict-first in the
// L2), so that the fill replaces its own lines rather than lines a neighbouring kernel still wants -- or, in the flushed
// benchmark configuration, rather than forcing the write-back of another kernel's dirty lines.  -DOSSG_ZIPPER_PLAIN_ACCESS
// keeps ordinary loads and stores.
template <typename T> __device__ __forceinline__ T ld_stream(const T* p)
{
#if defined(OSSG_ZIPPER_PLAIN_ACCESS)
    return *p;
#else
    return __ldcs(p);
#endif
}
template <typename T> __device__ __forceinline__ void st_stream(T* p, T v)
{
#if defined(OSSG_ZIPPER_PLAIN_ACCESS)
    *p = v;
#else
    __stcs(p, v);
#endif
}

template <typename T, bool UNIT>
__device__ __forceinline__ void zipper_column_impl(const ossg_field_desc& f, int i /*1..Nx*/, int kk, double s, int ip)
{
    const int Nx = f.Nx, Ny = f.Ny, Hy = f.Hy;
    const size_t sx = (size_t)(Nx + 2 * f.Hx), sxy = sx * (size_t)(Ny + 2 * Hy);
    T* c = static_cast<T*>(f.data) + (size_t)(f.k0 + kk) * sxy;
    const unsigned flip = sign_flip(s);
    // halo row j = r + 1 (r = 0..Hy-1):  dst row Ny + j, src row Ny - j (y-Center) or Ny - j + 1 (y-Face)
    const T* src = c + (size_t)((f.locy ? Ny : Ny - 1) + Hy - 1) * sx + (ip + f.Hx - 1);   // r = 0, then -sx per row
    T* dst = c + (size_t)(Ny + Hy) * sx + (i + f.Hx - 1);                                   // r = 0, then +sx per row
    // redundant part of row Ny (y-Center fields, i > Nx/2), :102, :135 -- read before the halo stores
    const bool fix_row = (f.locy == OSSG_CENTER) && (i > Nx / 2);
    T vrow = T(0);
    if (fix_row) vrow = c[(size_t)(Ny + Hy - 1) * sx + (ip + f.Hx - 1)];
    int r = 0;
    for (; r + 4 <= Hy; r += 4) {
        const T v0 = ld_stream(src - (ptrdiff_t)((r + 0) * sx)), v1 = ld_stream(src - (ptrdiff_t)((r + 1) * sx));
        const T v2 = ld_stream(src - (ptrdiff_t)((r + 2) * sx)), v3 = ld_stream(src - (ptrdiff_t)((r + 3) * sx));
        st_stream(dst + (size_t)(r + 0) * sx, fold_value<UNIT>(v0, s, flip));
        st_stream(dst + (size_t)(r + 1) * sx, fold_value<UNIT>(v1, s, flip));
        st_stream(dst + (size_t)(r + 2) * sx, fold_value<UNIT>(v2, s, flip));
        st_stream(dst + (size_t)(r + 3) * sx, fold_value<UNIT>(v3, s, flip));
    }
    for (; r < Hy; ++r) st_stream(dst + (size_t)r * sx, fold_value<UNIT>(ld_stream(src - (ptrdiff_t)(r * sx)), s, flip));
    if (fix_row) c[(size_t)(Ny + Hy - 1) * sx + (i + f.Hx - 1)] = fold_value<UNIT>(vrow, s, flip);
}

template <typename T>
__device__ __forceinline__ void zipper_column(const ossg_field_desc& f, int i /*1..Nx*/, int kk)
{
    double s = f.sign;
    int ip = f.locx ? f.Nx - i + 2 : f.Nx - i + 1;             // zipper_boundary_condition.jl:73,90,110,125
    if (ip > f.Nx) { s = fabs(s); ip -= f.Nx; }                 // :74-75, :91-92
    if (sign_is_unit(s)) zipper_column_impl<T, true>(f, i, kk, s, ip);      // (uniform over the field up to the wrapped column)
    else zipper_column_impl<T, false>(f, i, kk, s, ip);
}

__global__ void __launch_bounds__(ZB) zipper_kernel(const __grid_constant__ FieldBatch b)
{
    const ossg_field_desc& f = b.f[blockIdx.z];
    const int i = blockIdx.x * ZB + threadIdx.x + 1;
    const int kk = blockIdx.y;
    if (i > f.Nx || kk >= f.nk) return;
    if (f.elem_type == OSSG_F64) zipper_column<double>(f, i, kk);
    else zipper_column<float>(f, i, kk);
}

// periodic west/east fill over the whole (y, z) parent extent (external Oceananigans periodic
// halo kernel); run after the zipper so the NW/NE corners carry the folded values.
template <typename T>
__device__ __forceinline__ void periodic_element(const ossg_field_desc& f, int h /*0..2Hx-1*/, size_t row)
{
    const size_t sx = (size_t)(f.Nx + 2 * f.Hx);
    T* c = static_cast<T*>(f.data) + row * sx;
    if (h < f.Hx) c[h] = c[f.Nx + h];                            // west halo <- east interior
    else c[f.Nx + h] = c[h];                                     // east halo <- west interior  (h in Hx..2Hx-1)
}

__global__ void __launch_bounds__(ZB) periodic_x_kernel(const __grid_constant__ FieldBatch b)
{
    const ossg_field_desc& f = b.f[blockIdx.z];
    const int w = 2 * f.Hx;
    const size_t rows = (size_t)(f.Ny + 2 * f.Hy) * (size_t)(f.Nz + 2 * f.Hz);
    const size_t t = (size_t)blockIdx.x * ZB + threadIdx.x;
    if (t >= rows * w) return;
    const size_t row = t / w;
    const int h = (int)(t - row * w);
    if (f.elem_type == OSSG_F64) periodic_element<double>(f, h, row);
    else periodic_element<float>(f, h, row);
}

// ------------------------------------------------------------------ fused fill_halo_regions!
// Zipper north fill AND periodic west/east fill in ONE launch (Oceananigans runs them as two
// tasks, the periodic one second so that the NW/NE corners carry folded values).  The ordering is
// removed instead of enforced: the thread that produces a folded value for an edge column
// (i <= Hx or i > Nx-Hx) also stores its periodic image, so no thread ever reads what another
// thread of the same launch writes:
//   part A  one thread per (i, zipper level): the zipper column of zipper_column() plus the periodic
//           images of the north-halo rows and -- y-Center fields -- of row Ny;
//   part B  one thread per (row, side) for every other row of the parent (interior and south-halo
//           rows, all rows of the z-halo levels): a plain periodic copy of Hx elements.
// Requires Hx < Nx/2 (an edge column then has exactly one image and never is the self-mirrored
// column Nx/2+1 of the (Face, Center) sign quirk); smaller grids use the two-launch form.
template <typename T, bool UNIT>
__device__ __forceinline__ void fused_column_impl(const ossg_field_desc& f, int i /*1..Nx*/, int kk, double s, int ip)
{
    const int Nx = f.Nx, Ny = f.Ny, Hy = f.Hy, Hx = f.Hx;
    const size_t sx = (size_t)(Nx + 2 * Hx), sxy = sx * (size_t)(Ny + 2 * Hy);
    T* c = static_cast<T*>(f.data) + (size_t)(f.k0 + kk) * sxy;
    const unsigned flip = sign_flip(s);
    const T* src = c + (size_t)((f.locy ? Ny : Ny - 1) + Hy - 1) * sx + (ip + Hx - 1);
    T* dst = c + (size_t)(Ny + Hy) * sx + (i + Hx - 1);
    T* rowNy = c + (size_t)(Ny + Hy - 1) * sx + (i + Hx - 1);
    const bool img_e = i <= Hx, img_w = i > Nx - Hx;                 // periodic image at i + Nx, resp. i - Nx
    const ptrdiff_t img = img_e ? (ptrdiff_t)Nx : -(ptrdiff_t)Nx;
    const bool has_img = img_e || img_w;
    const bool fix_row = (f.locy == OSSG_CENTER) && (i > Nx / 2);
    T vrow = T(0);
    if (f.locy == OSSG_CENTER && (fix_row || has_img)) vrow = fix_row ? c[(size_t)(Ny + Hy - 1) * sx + (ip + Hx - 1)] : *rowNy;
    int r = 0;
    for (; r + 4 <= Hy; r += 4) {
        T v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = ld_stream(src - (ptrdiff_t)((r + u) * sx));
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const T o = fold_value<UNIT>(v[u], s, flip);
            st_stream(dst + (size_t)(r + u) * sx, o);
            if (has_img) dst[(ptrdiff_t)((r + u) * sx) + img] = o;
        }
    }
    for (; r < Hy; ++r) {
        const T o = fold_value<UNIT>(ld_stream(src - (ptrdiff_t)(r * sx)), s, flip);
        st_stream(dst + (size_t)r * sx, o);
        if (has_img) dst[(ptrdiff_t)(r * sx) + img] = o;
    }
    if (f.locy == OSSG_CENTER) {
        const T o = fix_row ? fold_value<UNIT>(vrow, s, flip) : vrow;      // (the periodic image of an untouched row Ny is a plain copy)
        if (fix_row) *rowNy = o;
        if (has_img) rowNy[img] = o;
    }
}

template <typename T>
__device__ __forceinline__ void fused_column(const ossg_field_desc& f, int i /*1..Nx*/, int kk)
{
    double s = f.sign;
    int ip = f.locx ? f.Nx - i + 2 : f.Nx - i + 1;
    if (ip > f.Nx) { s = fabs(s); ip -= f.Nx; }
    if (sign_is_unit(s)) fused_column_impl<T, true>(f, i, kk, s, ip);
    else fused_column_impl<T, false>(f, i, kk, s, ip);
}

// ---- part B.  A halo piece is Hx elements of a row; the common piece (4 doubles, 32-byte aligned) is
// one 256-bit load and one 256-bit store.  Memory picture of two consecutive parent rows r, r+1
// (row pitch (Nx+2Hx) elements): the last 2Hx elements of row r and the first 2Hx of row r+1 are
// contiguous -- the "seam"
//        [ E-interior(r) | E-halo(r) | W-halo(r+1) | W-interior(r+1) ]        (4 x 32 B = 128 B)
// with  W-halo(r) <- E-interior(r)  and  E-halo(r) <- W-interior(r).
// Measured in round 2 on the 1/12 degree fields (profiles/r02_fill_dram.md): thread per (row, side),
// thread per row, thread per two rows, thread per seam with a lane exchange, two launches instead of
// one, L1::no_allocate, L2::128B, and cudaLimitMaxL2FetchGranularity = 32 / 128 all run within 2 % of
// each other (91 ... 94 us for 4 fields, 99 us in two launches).  The L2 fetches 128 bytes around every
// 32-byte piece by default (3 lines per 2 rows: the row pitch is an odd multiple of 64 bytes), 238 MB
// of DRAM reads for 119 MB requested; the L2::64B prefetch-size qualifier brings that to 178 MB and
// the time from 93 to 91 us.  What bounds this part is the count of small accesses, not their bytes and not the
// row pitch (it costs 51 ... 61 us for 1.87 M + 1.87 M pieces at every pitch from 576 bytes to 34 KB,
// tools/fill_pitch_probe.py): 64-byte DRAM accesses move at ~3.3 TB/s here.  Full-granule stores and
// interleaving this part with the fold part were measured too (no gain, profiles/r02_fill_dram.md).
// Kept: thread per row with the L2::64B loads (fewest bytes, fastest by 2 %).
struct Piece32 { double a, b, c, d; };
// every halo piece of the field is a 32-byte aligned run of 4 doubles (base, row pitch and Nx all multiples of 32 bytes)
__host__ __device__ inline bool piece32_ok(const ossg_field_desc& f)
{
    return f.elem_type == OSSG_F64 && f.Hx == 4 && f.Nx % 4 == 0 && (reinterpret_cast<uintptr_t>(f.data) & 31) == 0;
}
__device__ __forceinline__ Piece32 ld_piece32(const void* p)
{
    Piece32 v;
    asm volatile("ld.global.L2::64B.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.a), "=d"(v.b), "=d"(v.c), "=d"(v.d) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_piece32(void* p, const Piece32& v)
{
#if defined(OSSG_ZIPPER_PLAIN_ACCESS)
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p), "d"(v.a), "d"(v.b), "d"(v.c), "d"(v.d) : "memory");
#else
    asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p), "d"(v.a), "d"(v.b), "d"(v.c), "d"(v.d) : "memory");
#endif
}

// rows whose periodic images part A writes: north-halo rows (and row Ny of y-Center fields) of the zipper levels
__device__ __forceinline__ bool part_a_row(const ossg_field_desc& f, long long row)
{
    const int sy = f.Ny + 2 * f.Hy;
    const int lev = (int)(row / sy), jrow = (int)(row - (long long)lev * sy);
    return lev >= f.k0 && lev < f.k0 + f.nk && jrow >= f.Ny + f.Hy - (f.locy == OSSG_CENTER ? 1 : 0);
}

// general form: one thread per (row, side), load before store
template <typename T>
__device__ __forceinline__ void fused_periodic_row(const ossg_field_desc& f, size_t row, int side)
{
    const int Nx = f.Nx, Hx = f.Hx;
    T* c = static_cast<T*>(f.data) + row * (size_t)(Nx + 2 * Hx);
    const T* src = side ? c + Hx : c + Nx;                           // east halo <- west interior ; west halo <- east interior
    T* dst = side ? c + Nx + Hx : c;
    copy_halo_piece(dst, src, Hx);
}

struct FusedBatch {
    FieldBatch b;
    int blocksA[MAX_BATCH];          // part A blocks of field k: ceil(Nx / ZB) * nk
    int chunksA[MAX_BATCH];          // ceil(Nx / ZB)
};

__global__ void __launch_bounds__(ZB) fill_fused_kernel(const __grid_constant__ FusedBatch fb)
{
    const int k = blockIdx.y;
    const ossg_field_desc& f = fb.b.f[k];
    const int nA = fb.blocksA[k];
    if ((int)blockIdx.x < nA) {
        const int chunk = blockIdx.x % fb.chunksA[k], kk = blockIdx.x / fb.chunksA[k];
        const int i = chunk * ZB + threadIdx.x + 1;
        if (i > f.Nx) return;
        if (f.elem_type == OSSG_F64) fused_column<double>(f, i, kk);
        else fused_column<float>(f, i, kk);
        return;
    }
    const long long rows = (long long)(f.Ny + 2 * f.Hy) * (f.Nz + 2 * f.Hz);
    const long long t = (long long)(blockIdx.x - nA) * ZB + threadIdx.x;
    if (piece32_ok(f)) {                                              // thread per row: both loads, then both stores
        if (t >= rows || part_a_row(f, t)) return;
        char* c = static_cast<char*>(f.data) + (size_t)t * (size_t)(f.Nx + 2 * f.Hx) * 8;
        const Piece32 e = ld_piece32(c + (size_t)f.Nx * 8), w = ld_piece32(c + 32);
        st_piece32(c, e);
        st_piece32(c + (size_t)f.Nx * 8 + 32, w);
        return;
    }
    if (t >= 2 * rows) return;
    const long long row = t >> 1;
    if (part_a_row(f, row)) return;
    if (f.elem_type == OSSG_F64) fused_periodic_row<double>(f, (size_t)row, (int)(t & 1));
    else fused_periodic_row<float>(f, (size_t)row, (int)(t & 1));
}

// ------------------------------------------------------------------ (x,y)-partition fold: pack / unpack
struct FoldArgs {
    ossg_field_desc f;
    ossg_fold_window w;
    void* buffer;
};

template <typename T>
__device__ __forceinline__ void pack_element(const FoldArgs& a, int c, int r, int kk)
{
    const ossg_field_desc& f = a.f;
    const size_t sx = (size_t)(f.Nx + 2 * f.Hx), sxy = sx * (size_t)(f.Ny + 2 * f.Hy);
    const int rows = fold_rows(f);
    const int li = a.w.src_i0 + c - a.w.i0 + 1;                 // local 1-based column of global src_i0 + c
    const int js = (r < f.Hy) ? (f.locy ? f.Ny - r : f.Ny - r - 1) : f.Ny;
    const T* src = static_cast<const T*>(f.data) + (size_t)(f.k0 + kk) * sxy;
    static_cast<T*>(a.buffer)[((size_t)kk * rows + r) * a.w.src_n + c] = src[(size_t)(js + f.Hy - 1) * sx + (li + f.Hx - 1)];
}

__global__ void __launch_bounds__(ZB) fold_pack_kernel(const __grid_constant__ FoldArgs a)
{
    const int c = blockIdx.x * ZB + threadIdx.x;
    if (c >= a.w.src_n) return;
    const int rows = fold_rows(a.f);
    const int row = blockIdx.y;
    const int kk = row / rows, r = row - kk * rows;
    if (a.f.elem_type == OSSG_F64) pack_element<double>(a, c, r, kk);
    else pack_element<float>(a, c, r, kk);
}

template <typename T>
__device__ __forceinline__ void unpack_element(const FoldArgs& a, int li /*1..nx local*/, int r, int kk)
{
    const ossg_field_desc& f = a.f;
    const int gNx = a.w.gNx;
    const int i = a.w.i0 + li - 1;                               // global destination column
    double s = f.sign;
    int ip = f.locx ? gNx - i + 2 : gNx - i + 1;
    if (ip > gNx) { s = fabs(s); ip -= gNx; }
    const int c = ip - a.w.src_i0;
    if (c < 0 || c >= a.w.src_n) return;                         // that column comes from another message
    const size_t sx = (size_t)(f.Nx + 2 * f.Hx), sxy = sx * (size_t)(f.Ny + 2 * f.Hy);
    const int rows = fold_rows(f);
    int jd;
    if (r < f.Hy) jd = f.Ny + r + 1;
    else { if (i <= gNx / 2) return; jd = f.Ny; }
    T* dst = static_cast<T*>(f.data) + (size_t)(f.k0 + kk) * sxy;
    const T v = static_cast<const T*>(a.buffer)[((size_t)kk * rows + r) * a.w.src_n + c];
    dst[(size_t)(jd + f.Hy - 1) * sx + (li + f.Hx - 1)] = apply_sign(v, s);
}

__global__ void __launch_bounds__(ZB) fold_unpack_kernel(const __grid_constant__ FoldArgs a)
{
    const int li = blockIdx.x * ZB + threadIdx.x + 1;
    if (li > a.f.Nx) return;
    const int rows = fold_rows(a.f);
    const int row = blockIdx.y;
    const int kk = row / rows, r = row - kk * rows;
    if (a.f.elem_type == OSSG_F64) unpack_element<double>(a, li, r, kk);
    else unpack_element<float>(a, li, r, kk);
}

// ------------------------------------------------------------------ host side
static int check_field(const ossg_field_desc& in, ossg_field_desc& f, int idx)
{
    f = in;
    if (!f.data) return set_error(OSSG_EINVAL, "field %d: data is NULL", idx);
    if (f.elem_type != OSSG_F64 && f.elem_type != OSSG_F32) return set_error(OSSG_EINVAL, "field %d: elem_type must be OSSG_F64 or OSSG_F32", idx);
    const size_t esz = f.elem_type == OSSG_F64 ? 8 : 4;
    if (reinterpret_cast<uintptr_t>(f.data) % esz) return set_error(OSSG_EALIGN, "field %d: data is not %zu-byte aligned", idx, esz);
    if (f.Nx < 2 || f.Ny < 2 || f.Nz < 1) return set_error(OSSG_EINVAL, "field %d: size (%d, %d, %d) too small", idx, f.Nx, f.Ny, f.Nz);
    if (f.Hx < 0 || f.Hy < 0 || f.Hz < 0) return set_error(OSSG_EINVAL, "field %d: negative halo", idx);
    if (f.Hy > f.Ny - 1 || f.Hx > f.Nx) return set_error(OSSG_EHALO, "field %d: halo (%d, %d) too large for size (%d, %d)", idx, f.Hx, f.Hy, f.Nx, f.Ny);
    if ((f.locx != OSSG_CENTER && f.locx != OSSG_FACE) || (f.locy != OSSG_CENTER && f.locy != OSSG_FACE))
        return set_error(OSSG_ELOCATION, "field %d: location codes must be OSSG_CENTER or OSSG_FACE", idx);
    if (f.nk == 0) { f.k0 = f.Hz; f.nk = f.Nz; }
    if (f.k0 < 0 || f.nk < 0 || f.k0 + f.nk > f.Nz + 2 * f.Hz) return set_error(OSSG_EINVAL, "field %d: k-range [%d, %d) outside the parent array", idx, f.k0, f.k0 + f.nk);
    return 0;
}

static int launch_zipper_batch(const ossg_field_desc* fields, int n, bool with_periodic, cudaStream_t s)
{
    if (!fields || n < 0) return set_error(OSSG_EINVAL, "fields is NULL or nfields < 0");
    for (int base = 0; base < n; base += MAX_BATCH) {
        FieldBatch b;
        b.n = n - base < MAX_BATCH ? n - base : MAX_BATCH;
        int gx = 0, gy = 0;
        size_t pmax = 0;
        bool fusable = with_periodic;
        for (int k = 0; k < b.n; ++k) {
            const ossg_field_desc& in = fields[base + k];
            if (in.Nx % 2 != 0) return set_error(OSSG_EODD_NX, "field %d: Nx must be even for the fold", base + k);
            int rc = check_field(in, b.f[k], base + k);
            if (rc) return rc;
            if (b.f[k].Hy < 1 || b.f[k].Hx < 1 || 2 * b.f[k].Hx >= b.f[k].Nx || b.f[k].nk < 1) fusable = false;
        }
        if (fusable) {                       // zipper + periodic in one launch
            FusedBatch fb;
            fb.b = b;
            long long gxf = 0;
            for (int k = 0; k < b.n; ++k) {
                const ossg_field_desc& f = b.f[k];
                fb.chunksA[k] = ceil_div(f.Nx, ZB);
                fb.blocksA[k] = fb.chunksA[k] * f.nk;
                const long long rows = (long long)(f.Ny + 2 * f.Hy) * (f.Nz + 2 * f.Hz);
                const long long threadsB = piece32_ok(f) ? rows : 2 * rows;
                const long long blocks = fb.blocksA[k] + (threadsB + ZB - 1) / ZB;
                gxf = gxf > blocks ? gxf : blocks;
            }
            if (gxf > 0x7fffffffLL) return set_error(OSSG_EINVAL, "field too large for one launch");
            fill_fused_kernel<<<dim3((unsigned)gxf, b.n), ZB, 0, s>>>(fb);
            OSSG_LAUNCH_CHECK("fill_fused_kernel");
            continue;
        }
        for (int k = 0; k < b.n; ++k) {      // (Hy = 0: zipper_column still performs the row-Ny overwrite of y-Center fields, :102, :135)
            gx = max(gx, ceil_div(b.f[k].Nx, ZB));
            gy = max(gy, b.f[k].nk);
            const size_t prow = (size_t)(b.f[k].Ny + 2 * b.f[k].Hy) * (size_t)(b.f[k].Nz + 2 * b.f[k].Hz) * (size_t)(2 * b.f[k].Hx);
            pmax = pmax > prow ? pmax : prow;
        }
        if (gx > 0 && gy > 0) {
            if (gy > 65535) return set_error(OSSG_EINVAL, "too many levels for one launch: %d", gy);
            zipper_kernel<<<dim3(gx, gy, b.n), ZB, 0, s>>>(b);
            OSSG_LAUNCH_CHECK("zipper_kernel");
        }
        if (with_periodic && pmax > 0) {
            periodic_x_kernel<<<dim3(ceil_div((long long)pmax, ZB), 1, b.n), ZB, 0, s>>>(b);
            OSSG_LAUNCH_CHECK("periodic_x_kernel");
        }
    }
    return 0;
}

}  // namespace ossg

using namespace ossg;

extern "C" int ossg_zipper_fill(const ossg_field_desc* f, void* stream)
{
    return launch_zipper_batch(f, 1, false, as_stream(stream));
}

extern "C" int ossg_zipper_fill_batch(const ossg_field_desc* f, int nfields, void* stream)
{
    return launch_zipper_batch(f, nfields, false, as_stream(stream));
}

extern "C" int ossg_fill_halo_regions_batch(const ossg_field_desc* f, int nfields, void* stream)
{
    return launch_zipper_batch(f, nfields, true, as_stream(stream));
}

static int check_fold(const ossg_field_desc* fin, const ossg_fold_window* w, ossg_field_desc& f)
{
    if (!fin || !w) return set_error(OSSG_EINVAL, "field or window is NULL");
    int rc = check_field(*fin, f, 0);
    if (rc) return rc;
    if (w->gNx < 2 || w->gNx % 2) return set_error(OSSG_EODD_NX, "global Nx must be even, got %d", w->gNx);
    if (w->nx != f.Nx) return set_error(OSSG_EWINDOW, "window nx (%d) differs from the field's local Nx (%d)", w->nx, f.Nx);
    if (w->i0 < 1 || w->i0 + w->nx - 1 > w->gNx) return set_error(OSSG_EWINDOW, "local columns [%d, %d] outside 1..%d", w->i0, w->i0 + w->nx - 1, w->gNx);
    if (w->src_n < 0 || w->src_i0 < 1 || w->src_i0 + w->src_n - 1 > w->gNx) return set_error(OSSG_EWINDOW, "source columns [%d, %d) outside 1..%d", w->src_i0, w->src_i0 + w->src_n, w->gNx);
    return 0;
}

extern "C" size_t ossg_fold_buffer_elems(const ossg_field_desc* fin, const ossg_fold_window* w)
{
    ossg_field_desc f;
    if (check_fold(fin, w, f)) return 0;
    return (size_t)f.nk * fold_rows(f) * (size_t)w->src_n;
}

extern "C" int ossg_fold_pack(const ossg_field_desc* fin, const ossg_fold_window* w, void* buffer, void* stream)
{
    FoldArgs a;
    int rc = check_fold(fin, w, a.f);
    if (rc) return rc;
    if (!buffer) return set_error(OSSG_EINVAL, "buffer is NULL");
    // the packed columns must exist in the local array (interior or x-halo)
    const int lo = w->src_i0 - w->i0 + 1, hi = lo + w->src_n - 1;
    if (w->src_n > 0 && (lo < 1 - a.f.Hx || hi > a.f.Nx + a.f.Hx))
        return set_error(OSSG_EWINDOW, "source columns [%d, %d] are not held by this rank (local %d..%d)", w->src_i0, w->src_i0 + w->src_n - 1, lo, hi);
    a.w = *w; a.buffer = buffer;
    const int gy = fold_rows(a.f) * a.f.nk;
    if (w->src_n == 0 || gy == 0) return 0;
    fold_pack_kernel<<<dim3(ceil_div(w->src_n, ZB), gy), ZB, 0, as_stream(stream)>>>(a);
    OSSG_LAUNCH_CHECK("fold_pack_kernel");
    return 0;
}

extern "C" int ossg_fold_unpack(const ossg_field_desc* fin, const ossg_fold_window* w, const void* buffer, void* stream)
{
    FoldArgs a;
    int rc = check_fold(fin, w, a.f);
    if (rc) return rc;
    if (!buffer) return set_error(OSSG_EINVAL, "buffer is NULL");
    a.w = *w; a.buffer = const_cast<void*>(buffer);
    const int gy = fold_rows(a.f) * a.f.nk;
    if (w->src_n == 0 || gy == 0) return 0;
    fold_unpack_kernel<<<dim3(ceil_div(a.f.Nx, ZB), gy), ZB, 0, as_stream(stream)>>>(a);
    OSSG_LAUNCH_CHECK("fold_unpack_kernel");
    return 0;
}
