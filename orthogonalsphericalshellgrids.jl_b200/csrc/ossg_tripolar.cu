// csrc/ossg_tripolar.cu -- TripolarGrid coordinate + metric generation on sm_100a.
//
// Kernels (all FP64 arithmetic, outputs FP64 or rounded once to FP32):
//   K0  tables_kernel         1-D tables: a*sind/cosd(lam1D), cosh/sinh(psi(phi1D)), lat-lon rows
//   K1  coords_pair_kernel    8 padded coordinate arrays, one thread per mirror pair of columns (the map is
//                             exactly mirror symmetric): whole-x windows and the halves of an Rx = 2 partition;
//                             branch-free straight-line point with a table-driven atan, one copy of the hot loop
//                             per x location; circshift + zipper/periodic halos are index maps (every element
//                             evaluated at its source)
//   K1' coords_kernel         the same per padded element (windows narrower than half the x extent)
//   K2  metrics_tile_kernel   12 metric arrays over the interior: four coordinate patches staged in
//                             shared memory (cp.async; TMA box loads opt-in), per-point cos(phi) and
//                             unit vectors once per patch point, then 8 haversines + 2 spherical
//                             quadrilaterals per cell as one speculative straight-line block whose
//                             preconditions are integer max/min over high words.  The halo frame (north
//                             fold image, x halos, fold row Ny, south continuation rows; one thread per
//                             (element, metric) at its source) rides in extra block columns of this grid
//   K2f metrics_frame_kernel  the halo frame on its own grid (profiled build, opt-in tile variants, windows
//                             without a fast rectangle)
// One stream, three launches per build, chained with programmatic dependent launch (K1 behind K0, K2 behind K1).
//   cartesian_nodes_kernel    lat_lon_to_cartesian over a coordinate pair
// Compile with -fmad=false: see ossg_math.cuh for the determinism contract.
#include <cuda.h>
#include <cooperative_groups.h>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <utility>

#include "ossg_common.cuh"
#include "ossg_tripolar_core.cuh"

namespace ossg {

struct Window {
    int I0, J0;      // global padded index of window element (0, 0)
    int wnx, wny;    // window extents (halos included); pitch == wnx
};

struct BuildArgs {
    GridDims g;
    int flags;                    // ossg_tripolar_params.flags
    Window w;
    double south, north_poles_latitude, first_pole_longitude, radius;
    unsigned long long* stamp;    // workspace header: TABLE_HDR_WORDS 8-byte words (validity stamps of the tables)
    double* tables;               // workspace + header: tables_doubles() doubles
    double* coords64[8];          // FP64 coordinates (outputs when FT = double, scratch otherwise)
    void* coords_out[8];          // FT coordinates (== coords64 when FT = double)
    void* metrics_out[12];        // FT metrics
    // Row-band dependency between the coordinate and the tile kernel (nullptr: off).  One counter per window row:
    // every coordinate block adds its weight when its stores are done, a row is complete at `rows_target`; the tile
    // blocks wait for the ten rows of their patches instead of the whole coordinate grid (metrics_tile_kernel).
    unsigned int* rows_done;
    unsigned int rows_target;
};

__device__ __forceinline__ void row_block_done(unsigned int* rows_done, int jw, unsigned weight)
{
    if (!rows_done) return;                              // (uniform: off by default)
    __syncthreads();                                     // the block's coordinate stores are issued
    if (threadIdx.x == 0) {
        __threadfence();                                 // ... and visible device-wide before the count moves
        atomicAdd(rows_done + jw, weight);
    }
}

__host__ __device__ inline Tables make_tables(const double* base, int Nx, int Ny)
{
    Tables t;
    t.asF = base;            t.acF = base + Nx;       t.asC = base + 2 * (size_t)Nx; t.acC = base + 3 * (size_t)Nx;
    const double* y = base + 4 * (size_t)Nx;
    t.chF = y;               t.shF = y + Ny;          t.chC = y + 2 * (size_t)Ny;    t.shC = y + 3 * (size_t)Ny;
    t.ll = y + 4 * (size_t)Ny;
    return t;
}

// ------------------------------------------------------------------ K0
// The tables depend on a handful of scalars only, and every constructor call of a model run
// (with_halo, reconstruct_global_grid, src/with_halo.jl:5-44, src/distributed_tripolar_grid.jl:201-226)
// repeats them.  The workspace therefore carries two validity stamps in a small header:
//   words 0..4   x/y tables:   magic, Nx, Ny, bits(south), bits(north_poles_latitude)
//   words 5..10  lat-lon rows: magic, Nx, Ny, Hy, bits(south), bits(radius)
//   word  11     arrival counter of the kernel that fills them
// A launch whose keys match returns after reading the header (about 2 us instead of the 15 us
// double-double chain); otherwise the last block to finish writes the keys.  The caller zeroes the
// header once per allocation (ossg_tripolar_workspace_init), so stale memory never passes for a stamp.
constexpr int TABLE_HDR_WORDS = 16;
constexpr unsigned long long STAMP_XY = 0x4f5353475f58595full, STAMP_LL = 0x4f5353475f4c4c5full;   // "OSSG_XY_", "OSSG_LL_"

struct TableKeys { unsigned long long xy[5], ll[6]; };
__host__ __device__ inline TableKeys table_keys(const GridDims& g, int flags, double south, double npl, double radius)
{
    auto bits = [](double x) { unsigned long long u; memcpy(&u, &x, 8); return u; };
    TableKeys k;
    k.xy[0] = STAMP_XY; k.xy[1] = (unsigned)g.Nx | ((unsigned long long)(unsigned)flags << 32); k.xy[2] = (unsigned)g.Ny; k.xy[3] = bits(south); k.xy[4] = bits(npl);
    k.ll[0] = STAMP_LL; k.ll[1] = (unsigned)g.Nx; k.ll[2] = (unsigned)g.Ny; k.ll[3] = (unsigned)g.Hy; k.ll[4] = bits(south); k.ll[5] = bits(radius);
    return k;
}

__global__ void __launch_bounds__(128) tables_kernel(GridDims g, int flags, double south, double north_poles_latitude,
                                                     double radius, double* __restrict__ base, unsigned long long* stamp,
                                                     unsigned int* rows_done, int wny)
{
    // This grid is itself launched with programmatic stream serialisation: behind a tile kernel of an earlier build on the
    // same stream (which lets its dependents go early) its blocks -- and behind them the coordinate kernel's -- are
    // resident before that grid has finished, and wait here.  Nothing of this build touches memory before the wait, and
    // the coordinate kernel waits for this grid in turn, so a build never overtakes the one before it; what is saved is
    // the launch latency of two kernels.  Behind any other operation the launch is an ordinary one.
    cudaTriggerProgrammaticLaunchCompletion();
    cudaGridDependencySynchronize();
    // (row counters of the coordinate -> tile dependency: zero for this build; the coordinate kernel waits for this grid)
    if (rows_done) { const int r = blockIdx.x * blockDim.x + threadIdx.x; if (r < wny) rows_done[r] = 0u; }
    // (the coordinate kernel behind this one was let go above; it waits -- cudaGridDependencySynchronize -- for this grid
    // to finish before it reads a table entry)
    const TableKeys key = table_keys(g, flags, south, north_poles_latitude, radius);
    // all eleven header words in flight at once (six 16-byte loads: one L2 round trip, not eleven in a chain)
    unsigned long long hw[12];
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        const ulonglong2 v = __ldcg(reinterpret_cast<const ulonglong2*>(stamp) + q);
        hw[2 * q] = v.x; hw[2 * q + 1] = v.y;
    }
    unsigned long long dxy = 0, dll = 0;
#pragma unroll
    for (int q = 0; q < 5; ++q) dxy |= hw[q] ^ key.xy[q];
#pragma unroll
    for (int q = 0; q < 6; ++q) dll |= hw[5 + q] ^ key.ll[q];
    const bool xy_ok = dxy == 0, ll_ok = dll == 0;
    if (xy_ok && ll_ok) return;                             // (uniform over the grid: nobody writes the header before every block has read it)
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int Nx = g.Nx, Ny = g.Ny;
    double* xs = base;
    double* ys = base + 4 * (size_t)Nx;
    double* ll = ys + 4 * (size_t)Ny;
    if (t < 2 * Nx) {
        if (!xy_ok) {
            // x tables, stored in post-circshift order; one thread per (face|center, i)
            const int face = t < Nx;
            const int i = face ? t + 1 : t - Nx + 1;
            const int ip = preshift(i, Nx);
            const double a = focal_distance(north_poles_latitude);
            double lam = face ? lam_face(ip, Nx) : lam_center(ip, Nx);
            if (flags & OSSG_LAMBDA1D_F32) lam = (double)(float)lam;    // generate_coordinate(Float32, ...) reading, tripolar_grid.jl:89
            const int o = face ? 0 : 2;
            xs[(size_t)(o + 0) * Nx + (i - 1)] = a * m::sind(lam);      // generate_tripolar_coordinates.jl:67
            xs[(size_t)(o + 1) * Nx + (i - 1)] = a * m::cosd(lam);      // :68
        }
    } else if (t < 2 * Nx + 2 * Ny) {
        if (!xy_ok) {
            const int u = t - 2 * Nx;
            const int face = u < Ny;
            const int j = face ? u + 1 : u - Ny + 1;
            const double a = focal_distance(north_poles_latitude);
            const double phi = face ? phi_face(j, Ny, south) : phi_center(j, Ny, south);
            const int o = face ? 0 : 2;
            double psi, sh, ch;
            m::asinh_sinh_cosh_cr(m::tand((90.0 - phi) / 2) / a, psi, sh, ch);   // :66-68
            ys[(size_t)(o + 0) * Ny + (j - 1)] = ch;
            ys[(size_t)(o + 1) * Ny + (j - 1)] = sh;
        }
    } else if (t < 2 * Nx + 2 * Ny + g.Hy + 1) {
        if (!ll_ok) {
            const int r = t - 2 * Nx - 2 * Ny;                  // row r <-> j = r + 1 - Hy
            latlon_row(r + 1 - g.Hy, Nx, Ny, g.Hy, south, radius, ll + (size_t)r * LL_NCOL);
        }
    }
    // the last block to arrive stamps the header (every block has read it by then)
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned int* counter = reinterpret_cast<unsigned int*>(stamp + 11);
        last = atomicAdd(counter, 1u) == gridDim.x - 1;
        if (last) {
            __threadfence();
            for (int q = 0; q < 5; ++q) stamp[q] = key.xy[q];
            for (int q = 0; q < 6; ++q) stamp[5 + q] = key.ll[q];
            *counter = 0u;
        }
    }
}

// ------------------------------------------------------------------ K1
template <typename FT>
__device__ __forceinline__ void coords_body(const BuildArgs& a)
{
    const int iw = blockIdx.x * blockDim.x + threadIdx.x;
    const int jw = blockIdx.y;
    if (iw >= a.w.wnx) return;
    const int I = a.w.I0 + iw, J = a.w.J0 + jw;
    const GridDims g = a.g;
    const Tables tb = make_tables(a.tables, g.Nx, g.Ny);
    const size_t o = (size_t)jw * a.w.wnx + iw;
    const bool bulk = I >= 1 && I <= g.Nx && J >= 1 && J < g.Ny;   // all four staggerings sit at their own (I, J)
    double asF = 0, acF = 0, asC = 0, acC = 0, chF = 0, shF = 0, chC = 0, shC = 0;
    bool one = false, half = false;
    if (bulk) {
        const int ip = preshift(I, g.Nx);
        one = ip == 1; half = ip <= g.Nx / 2;
        asF = __ldg(tb.asF + I - 1); acF = __ldg(tb.acF + I - 1); asC = __ldg(tb.asC + I - 1); acC = __ldg(tb.acC + I - 1);
        chF = __ldg(tb.chF + J - 1); shF = __ldg(tb.shF + J - 1); chC = __ldg(tb.chC + J - 1); shC = __ldg(tb.shC + J - 1);
    }
    // the four staggerings as a rolled loop: one copy of the point code in the instruction cache
#pragma unroll 1
    for (int loc = 0; loc < 4; ++loc) {
        double lam, phi;
        if (bulk) {
            const bool fx = loc & 1, fy = loc & 2;
            tripolar_point(fx ? asF : asC, fx ? acF : acC, fy ? chF : chC, fy ? shF : shC, one, half,
                           a.first_pole_longitude, lam, phi);
        } else {
            coord_at(tb, g, a.first_pole_longitude, loc, I, J, lam, phi);
        }
        a.coords64[loc][o] = lam;
        a.coords64[4 + loc][o] = phi;
        if (sizeof(FT) == 4) {
            static_cast<FT*>(a.coords_out[loc])[o] = (FT)lam;
            static_cast<FT*>(a.coords_out[4 + loc])[o] = (FT)phi;
        }
    }
}
template <typename FT>
__global__ void __launch_bounds__(256) coords_kernel(const __grid_constant__ BuildArgs a)
{
    cudaGridDependencySynchronize();                          // the table kernel in front of this launch (see launch_coords)
    cudaTriggerProgrammaticLaunchCompletion();                // the tile kernel behind may take the slots this grid's tail frees
    coords_body<FT>(a);
    row_block_done(a.rows_done, blockIdx.y, 1u);              // (a row is complete at gridDim.x)
}

// ------------------------------------------------------------------ K1 (mirror pairs)
// In rows 1..Ny-1 a thread evaluates the map once per staggering for pre-shift column ip = 1+t and
// writes both ip and its mirror column (see tripolar_core) wherever they fall inside the window: half the
// transcendental work of the per-element kernel for every window that holds its own mirror columns --
// the whole x extent (serial grids, Y-slab ranks) and the halves of an Rx = 2 partition (post-shift
// columns mirror as i <-> Nx/2+1-i resp. Nx/2+2-i, and the mirror of a half's east halo is its west halo).
// Narrower windows run the same code with one stored image per evaluation.  Rows outside 1..Ny-1 (row Ny,
// whose y-Center fields carry the fold overwrite, the north and the south halo) and, for full-x windows,
// the x halos (second images of interior columns) go through the per-element source map.
template <typename FT>
__device__ __forceinline__ void store_coord(const BuildArgs& a, int loc, size_t o, double lam, double phi)
{
    a.coords64[loc][o] = lam;
    a.coords64[4 + loc][o] = phi;
    if (sizeof(FT) == 4) {
        static_cast<FT*>(a.coords_out[loc])[o] = (FT)lam;
        static_cast<FT*>(a.coords_out[4 + loc])[o] = (FT)phi;
    }
}

#ifndef OSSG_COORD_ILP
#define OSSG_COORD_ILP 1
#endif
#ifndef OSSG_COORD_SPLIT_X
#define OSSG_COORD_SPLIT_X 1        /* one copy of the hot loop per x location: coordinate kernel 0.2055 -> 0.1914 ms (profiles/r02o_variants.txt) */
#endif
#ifndef OSSG_COORD_BLOCK
#define OSSG_COORD_BLOCK 128      /* threads per block of coords_pair_kernel, 8 blocks per SM: 0.2086 -> 0.2022 ms against 256 x 4 (finer tail), profiles/r02k_tiles.txt */
#endif
#ifndef OSSG_COORD_MINBLOCKS
#define OSSG_COORD_MINBLOCKS (1024 / OSSG_COORD_BLOCK)
#endif
// 1-D grid of coords_pair_kernel: the pair rows (J = 1..Ny-1) need Nx/2 + 1 + 2Hx threads, the other
// rows one thread per padded element; a 2-D grid sized for the latter would launch an empty
// second half for every pair row.
struct PairGrid {
    int jw_lo, n_pair_rows, chunks_pair, chunks_full;
    int t_lo, n_t;                 // pair threads of a row: t = t_lo .. t_lo + n_t - 1 (subset of 0 .. Nx/2)
    int n_halo;                    // per-element x-halo threads of a row behind them (full-x windows: 2 Hx, else 0)
#ifdef OSSG_COORD_MAGIC_DIV
    // ceil(2^40 / chunks) -- b / chunks == (b * magic) >> 40 for b * chunks < 2^40, three instructions instead of
    // the ~25 of an integer division in every block's prologue (default since round 2, see the Makefile)
    unsigned long long magic_pair, magic_full;
#endif
};
#ifdef OSSG_COORD_MAGIC_DIV
__device__ __forceinline__ int magic_div(int b, int d, unsigned long long magic) { (void)d; return (int)(((unsigned long long)(unsigned)b * magic) >> 40); }
#else
__device__ __forceinline__ int magic_div(int b, int d, unsigned long long) { return b / d; }
#endif

template <typename FT, bool SPLIT>
__device__ __forceinline__ void coords_pair_body(const BuildArgs& a, const PairGrid& pg, int jw, int chunk, const double* s_atan);

template <typename FT, bool SPLIT /* window narrower than the x extent: images may fall outside */>
__global__ void __launch_bounds__(OSSG_COORD_BLOCK, OSSG_COORD_MINBLOCKS) coords_pair_kernel(const __grid_constant__ BuildArgs a, const PairGrid pg)
{
    int jw, chunk;
    bool pair_row;
    {
        const int b = blockIdx.x, nb_pair = pg.n_pair_rows * pg.chunks_pair;
#ifdef OSSG_COORD_MAGIC_DIV
        const unsigned long long mp = pg.magic_pair, mf = pg.magic_full;
#else
        const unsigned long long mp = 0, mf = 0;
#endif
        // the rows outside 1..Ny-1 (halo rows of the fold and of the south: one general evaluation per element, the
        // longest threads of the grid) take the FIRST blocks, so that they run next to the pair rows instead of after them
        const int nb_full = (int)gridDim.x - nb_pair;
        pair_row = b >= nb_full;
        if (pair_row) {
            const int r = magic_div(b - nb_full, pg.chunks_pair, mp);
            jw = pg.jw_lo + r;
            chunk = (b - nb_full) - r * pg.chunks_pair;
        } else {
            const int r = magic_div(b, pg.chunks_full, mf);
            chunk = b - r * pg.chunks_full;
            jw = r < pg.jw_lo ? r : r + pg.n_pair_rows;
        }
    }
#ifdef OSSG_ATAN_RANGE_TABLE
    __shared__ double s_atan[30 + 11];                       // atan reduction table + 88 bytes of range table (ossg_math.cuh)
    for (int e = threadIdx.x; e < 30 + 88; e += OSSG_COORD_BLOCK) {
        if (e < 30) s_atan[e] = m::kATAN_LUT[e];
        else reinterpret_cast<unsigned char*>(s_atan + 30)[e - 30] = m::kATAN_RANGE[e - 30];
    }
#else
    __shared__ double s_atan[30];                            // atan reduction table (ossg_math.cuh)
    if (threadIdx.x < 30) s_atan[threadIdx.x] = m::kATAN_LUT[threadIdx.x];
#endif
    __syncthreads();
    cudaGridDependencySynchronize();                          // the table kernel in front of this launch (see launch_coords)
    cudaTriggerProgrammaticLaunchCompletion();                // the tile kernel behind may take the slots this grid's tail frees
    coords_pair_body<FT, SPLIT>(a, pg, jw, chunk, s_atan);
    // every row reaches chunks_pair * chunks_full: a pair row has chunks_pair blocks, any other row chunks_full
    row_block_done(a.rows_done, jw, pair_row ? (unsigned)pg.chunks_full : (unsigned)pg.chunks_pair);
}

template <typename FT, bool SPLIT>
__device__ __forceinline__ void coords_pair_body(const BuildArgs& a, const PairGrid& pg, int jw, int chunk, const double* s_atan)
{
    const int tt = chunk * blockDim.x + threadIdx.x;
    const GridDims g = a.g;
    const int Nx = g.Nx, Hx = g.Hx;
    const int J = a.w.J0 + jw;
    const Tables tb = make_tables(a.tables, Nx, g.Ny);
    const size_t row = (size_t)jw * a.w.wnx;
    const double fpl = a.first_pole_longitude;
    if (J >= 1 && J < g.Ny) {
        const int half = Nx / 2;
        if (tt < pg.n_t) {
            const int t = pg.t_lo + tt;
            const int ip = 1 + t;                               // pre-shift column of this pair
            const int shift = Nx / 4;
            int i = ip + shift; if (i > Nx) i -= Nx;            // its post-shift column
            // x-Face: ip <-> Nx+2-ip (ip = 1 and ip = Nx/2+1 are their own mirrors); x-Center: ip <-> Nx+1-ip
            const int ipmF = Nx + 2 - ip, ipmC = Nx + 1 - ip;
            int imF = ipmF + shift; if (imF > Nx) imF -= Nx;
            int imC = ipmC + shift; if (imC > Nx) imC -= Nx;
            // window column of global column c: the image c, c - Nx or c + Nx that lies inside the window, if any
            // (a window no wider than Nx holds at most one; a full-x window holds the interior image, its x halos
            // are served by the per-element threads below)
            const int I0 = a.w.I0, wnx = a.w.wnx;
            auto wcol = [&](int c, bool& inside) {
                int wc = c - I0;
                if (!SPLIT) { inside = true; return wc; }        // full-x window: c + Hx - 1, always inside
                if (wc >= wnx) wc -= Nx;
                if (wc < 0) wc += Nx;
                inside = (unsigned)wc < (unsigned)wnx;
                return wc;
            };
            bool in_i, in_mF, in_mC;
            const size_t o = row + (size_t)wcol(i, in_i);
            const size_t omF = row + (size_t)wcol(imF, in_mF), omC = row + (size_t)wcol(imC, in_mC);
            const bool mirrorF = (t != 0) && (t != half) && in_mF, pairC = t < half;
            const bool directC = pairC && in_i, mirrorC = pairC && in_mC;
            if (SPLIT && !(in_i || mirrorF || mirrorC)) return; // (bounding thread range of a narrow window)
            const double chF = __ldg(tb.chF + J - 1), shF = __ldg(tb.shF + J - 1);
            const double chC = __ldg(tb.chC + J - 1), shC = __ldg(tb.shC + J - 1);
            const double asF = __ldg(tb.asF + i - 1), acF = __ldg(tb.acF + i - 1);
            const double asC = __ldg(tb.asC + i - 1), acC = __ldg(tb.acC + i - 1);
            const double fpl90 = fpl + 90.0;
            // Speculative straight-line evaluation, OSSG_COORD_ILP staggerings per basic block (table-
            // driven branch-free atan, div_nb, sqrt_nb): ptxas interleaves the independent chains.  A
            // point that leaves the preconditions sets `bad`; the thread then redoes its four points
            // through the general functions and overwrites the speculative stores.
            bool bad = false;
#if OSSG_COORD_SPLIT_X
            // two copies of the loop body, one per x location (compile-time: the column tables, the mirror column and its
            // conditions are no longer selected at run time), each rolled over the y location
#pragma unroll
            for (int kx = 0; kx < 2; ++kx) {
                const bool fx = kx == 0;
                const double as_i = fx ? asF : asC, ac_i = fx ? acF : acC;
                const bool st_i = fx ? in_i : directC, st_m = fx ? mirrorF : mirrorC;
                const size_t om = fx ? omF : omC;
                const bool m_le_half = (fx ? ipmF : ipmC) <= half;
#if defined(OSSG_COORD_UNROLL_Y)
#pragma unroll
#else
#pragma unroll 1
#endif
                for (int ky = 0; ky < 2; ++ky) {                          // FC, FF | CC, CF
                    const bool fy = ky != 0;
                    const PointCore c = tripolar_core_sl(as_i, ac_i, fy ? chF : chC, fy ? shF : shC, s_atan, bad);
                    const int loc = fx ? (fy ? FF : FC) : (fy ? CF : CC);
                    if (st_i) store_coord<FT>(a, loc, o, tripolar_finish_sl(c.lam_a, ip <= half, fpl90, bad), c.phi);
                    if (st_m) store_coord<FT>(a, loc, om, tripolar_finish_sl(-c.lam_a, m_le_half, fpl90, bad), c.phi);
                }
            }
#else
#if defined(OSSG_COORD_UNROLL)
#pragma unroll
#else
#pragma unroll 1
#endif
            for (int k0 = 0; k0 < 4; k0 += OSSG_COORD_ILP) {
#pragma unroll
                for (int kk = 0; kk < OSSG_COORD_ILP; ++kk) {           // k = 0: FC, 1: FF, 2: CC, 3: CF
                    const int k = k0 + kk;
                    const bool fx = k < 2, fy = k & 1;
                    const PointCore c = tripolar_core_sl(fx ? asF : asC, fx ? acF : acC, fy ? chF : chC, fy ? shF : shC, s_atan, bad);
                    const int loc = fx ? (fy ? FF : FC) : (fy ? CF : CC);
                    if (fx ? in_i : directC) store_coord<FT>(a, loc, o, tripolar_finish_sl(c.lam_a, ip <= half, fpl90, bad), c.phi);
                    if (fx ? mirrorF : mirrorC)
                        store_coord<FT>(a, loc, fx ? omF : omC, tripolar_finish_sl(-c.lam_a, (fx ? ipmF : ipmC) <= half, fpl90, bad), c.phi);
                }
            }
#endif
            if (bad) {
#pragma unroll 1
                for (int k = 0; k < 4; ++k) {
                    const bool fx = k < 2, fy = k & 1;
                    if (!fx && !pairC) break;
                    const PointCore c = tripolar_core_lut(fx ? asF : asC, fx ? acF : acC, fy ? chF : chC, fy ? shF : shC, s_atan);
                    const int loc = fx ? (fy ? FF : FC) : (fy ? CF : CC);
                    if (fx ? in_i : directC) store_coord<FT>(a, loc, o, tripolar_finish(c.lam_a, c.pole, ip == 1, ip <= half, fpl), c.phi);
                    if (fx ? mirrorF : mirrorC)
                        store_coord<FT>(a, loc, fx ? omF : omC,
                                        tripolar_finish(-c.lam_a, c.pole, false, (fx ? ipmF : ipmC) <= half, fpl), c.phi);
                }
            }
        } else if (tt < pg.n_t + pg.n_halo) {
            // x halo element of this row (full-x windows): columns 1-Hx..0 and Nx+1..Nx+Hx
            const int h = tt - pg.n_t;                          // 0 .. 2Hx-1
            const int I = h < Hx ? 1 - Hx + h : Nx + 1 + (h - Hx);
            const size_t o = row + (size_t)(I + Hx - 1);
#pragma unroll 1
            for (int loc = 0; loc < 4; ++loc) {
                double lam, phi;
                coord_at(tb, g, fpl, loc, I, J, lam, phi);
                store_coord<FT>(a, loc, o, lam, phi);
            }
        }
        return;
    }
    // rows outside 1..Ny-1: one padded element per thread through the source map
    const int t = tt;
    if (t >= a.w.wnx) return;
    const int I = a.w.I0 + t;
#pragma unroll 1
    for (int loc = 0; loc < 4; ++loc) {
        double lam, phi;
        coord_at(tb, g, fpl, loc, I, J, lam, phi);
        store_coord<FT>(a, loc, row + t, lam, phi);
    }
}

// ------------------------------------------------------------------ K2 (tile)
// Tile of TX x TY cells.  Each staggering needs its coordinate on a (TX+1) x (TY+1) patch
// (tripolar_grid_utils.jl:13-43):
//     FC: i..i+1, j-1..j      CC: i-1..i, j-1..j      FF: i..i+1, j..j+1      CF: i-1..i, j..j+1
// The patch is staged in shared memory once together with everything that depends on one point
// only -- deg2rad(phi), cos(deg2rad(phi)) for the haversines and, for CC and FF, the unit vector
// of lat_lon_to_cartesian -- so that the four (resp. up to six) cells sharing a point do not
// repeat its trigonometry.
#ifndef OSSG_USE_TMA
#define OSSG_USE_TMA 0
#endif
#ifndef OSSG_FRAME_DEFAULT
#define OSSG_FRAME_DEFAULT 2          /* measured best on the whole grid and on the slabs of 2..8 ranks (profiles/r02g_slabs.txt) */
#endif
#ifndef OSSG_TX
#define OSSG_TX 32
#endif
#ifndef OSSG_TY
#define OSSG_TY 8
#endif
constexpr int TX = OSSG_TX, TY = OSSG_TY, NT = TX * TY;
constexpr int TILE_CTAS = 1024 / NT;                       // 32 warps of tile CTAs per SM
constexpr int QY = TY + 1;
// Patch geometry.  cp.async staging: (TX+1) x QY points, dense.  TMA staging: the box is (TX+2) x QY
// (the inner box extent must be a multiple of 16 bytes; the extra column is fetched and ignored) and
// every patch starts on a 128-byte boundary (TMA destination alignment): stride 320 doubles.
template <bool TMA> struct TileGeom {
    static constexpr int QX = TMA ? TX + 2 : TX + 1;
    static constexpr int QPTS = QX * QY;                 // logical points of a patch (TMA: incl. the pad column)
    static constexpr int PS = TMA ? 320 : QPTS;          // patch stride in shared memory, doubles
    static constexpr size_t SMEM_BYTES = (size_t)18 * PS * sizeof(double) + (TMA ? 128 : 0);
};
// shared arrays: lam, phr, cosphi per staggering (12) + unit vectors of CC and FF (6)
struct TileMaps { CUtensorMap m[8]; };                  // lam cc,fc,cf,ff, phi cc,fc,cf,ff of the window (TMA staging)

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned phase)
{
    const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    unsigned done = 0;
    const long long t0 = clock64();
    while (!done) {
        asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(phase) : "memory");
        if (!done && clock64() - t0 > 2000000000ll) __trap();        // a lost TMA transaction must not hang the GPU
    }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int x, int y, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(reinterpret_cast<unsigned long long>(map)),
                    "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(x), "r"(y) : "memory");
}

struct FastRect { int I0, J0, nx, ny; };   // global indices of the first fast cell, extents

// patch origin (offset of patch point (0,0) from tile cell (0,0)) per staggering
__device__ __forceinline__ int patch_di(int loc) { return (loc == CC || loc == CF) ? -1 : 0; }
__device__ __forceinline__ int patch_dj(int loc) { return (loc == CC || loc == FC) ? -1 : 0; }

__device__ __forceinline__ void cp_async_8(double* smem_dst, const double* gmem_src)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(sa), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

template <typename FT>
__device__ __forceinline__ void store_metric(const BuildArgs& a, int id, size_t o, double v)
{
    static_cast<FT*>(a.metrics_out[id])[o] = (FT)v;
}

// The three phases of the tile kernel as inline device functions (shared by metrics_tile_kernel and
// the mirror-pair variant metrics_tile_pair_kernel below).
#ifndef OSSG_PC_SOA
#define OSSG_PC_SOA 1            /* 0: paired layout below; measured neutral (tile kernel -1.7 us, build +1.6 us, profiles/r02l_variants.txt) */
#endif
struct PatchPoint { double lam, hphi, cosphi; };           // what a haversine needs of one point
template <bool TMA>
struct TileSmem {
    static constexpr int PS = TileGeom<TMA>::PS;
    // (deg2rad(phi)/2, cos(deg2rad(phi))) of a point as ONE 16-byte element: one LDS.128 per haversine operand instead
    // of two LDS.64 (-DOSSG_PC_SOA=0; the default keeps two arrays).  Not with TMA staging (the box load writes phi densely).
    static constexpr bool PAIRED = !TMA && !OSSG_PC_SOA;
    double* lam;      // [4][PS]  lambda (degrees)
    double* phr;      // [4][PS]  phi (degrees) during staging, then deg2rad(phi)/2       (PAIRED: [4][PS][2] with cosp)
    double* cosp;     // [4][PS]  cos(deg2rad(phi))
    __device__ __forceinline__ double& phi(int t) const { return PAIRED ? phr[2 * t] : phr[t]; }
    __device__ __forceinline__ double& cosphi(int t) const { return PAIRED ? phr[2 * t + 1] : cosp[t]; }
    __device__ __forceinline__ void set_point(int t, double hphi, double c) const
    {
        if (PAIRED) reinterpret_cast<double2*>(phr)[t] = make_double2(hphi, c);
        else { phr[t] = hphi; cosp[t] = c; }
    }
    __device__ __forceinline__ PatchPoint point(int t) const
    {
        if (PAIRED) { const double2 q = reinterpret_cast<const double2*>(phr)[t]; return PatchPoint{ lam[t], q.x, q.y }; }
        return PatchPoint{ lam[t], phr[t], cosp[t] };
    }
    double2* uxy;     // [2][PS]  unit vectors, (x, y) as one 16-byte element (one LDS.128 / STS.128): slot 0 = CC, slot 1 = FF
    double* uz;       // [2][PS]
    __device__ __forceinline__ explicit TileSmem(double* smem)
        : lam(smem), phr(smem + 4 * PS), cosp(smem + 8 * PS), uxy(reinterpret_cast<double2*>(smem + 12 * PS)), uz(smem + 16 * PS) {}
};

// ---- stage the four coordinate patches: global -> shared, asynchronously (LDGSTS), no registers.
// Thread (tx, ty) copies patch point (tx, ty) of all four patches (one offset computation, the
// staggering shifts are constants); the first QY + TX threads then copy column TX and row TY.
// Ends with the wait for the thread's own copies; the caller synchronises the block.
// CHECK = false: the whole (TX+1) x (TY+1) patch of every staggering lies inside the window (true for every tile
// that is not cut by the right / top edge of the fast rectangle, which itself stays one cell inside the window).
template <bool CHECK>
__device__ __forceinline__ void tile_stage_points(const BuildArgs& a, const Window& w, int It0, int Jt0, int tid, const TileSmem<false>& S)
{
    constexpr int QX = TileGeom<false>::QX, PS = TileGeom<false>::PS;
    auto stage_point = [&](int pi, int pj) {
        const int iw = It0 + pi - w.I0, jw = Jt0 + pj - w.J0;
        const long long o = (long long)jw * w.wnx + iw;
        const int p = pj * QX + pi;
#pragma unroll
        for (int loc = 0; loc < 4; ++loc) {
            const int di = patch_di(loc), dj = patch_dj(loc);
            const int i2 = iw + di, j2 = jw + dj;
            if (!CHECK || (i2 >= 0 && i2 < w.wnx && j2 >= 0 && j2 < w.wny)) {
                const long long o2 = o + (long long)dj * w.wnx + di;
                cp_async_8(&S.lam[loc * PS + p], a.coords64[loc] + o2);
                cp_async_8(&S.phi(loc * PS + p), a.coords64[4 + loc] + o2);
            } else {
                S.lam[loc * PS + p] = 0.0;
                S.phi(loc * PS + p) = 0.0;
            }
        }
    };
    stage_point(threadIdx.x, threadIdx.y);
    if (tid < QY) stage_point(TX, tid);
    else if (tid < QY + TX) stage_point(tid - QY, TY);
}

__device__ __forceinline__ void tile_stage_cp_async(const BuildArgs& a, const Window& w, int It0, int Jt0, int tid, const TileSmem<false>& S)
{
    const bool inside = It0 - 1 >= w.I0 && It0 + TX < w.I0 + w.wnx && Jt0 - 1 >= w.J0 && Jt0 + TY < w.J0 + w.wny;   // (block-uniform)
    if (inside) tile_stage_points<false>(a, w, It0, Jt0, tid, S);
    else tile_stage_points<true>(a, w, It0, Jt0, tid, S);
    cp_async_wait_all();
}

// ---- point stage: everything that depends on one point only, once per patch point, in ONE
// pass (every task reads and writes its own point only):
//   heavy task (CC and FF points, 2 QPTS of them): unit vector + deg2rad(phi)/2 + cos(deg2rad(phi))
//   light task (FC and CF points, 2 QPTS of them): deg2rad(phi)/2 + cos(deg2rad(phi))
// 32-task chunks are dealt statically so that the warps finish together: heavy chunks round-robin;
// the warps that got one heavy chunk fewer share the light chunks (a heavy task costs ~4 light ones).
template <bool TMA>
__device__ __forceinline__ void tile_point_stage(const TileSmem<TMA>& S, int tid)
{
    constexpr int QPTS = TileGeom<TMA>::QPTS, PS = TileGeom<TMA>::PS;
    constexpr int NW = NT / 32, HC = (2 * QPTS + 31) / 32, LC = HC;
    constexpr int W0 = HC % NW;                        // warps below W0 carry an extra heavy chunk
    constexpr int NLW = (W0 == 0) ? NW : NW - W0;      // warps sharing the light chunks
    const int warp = tid >> 5, lane = tid & 31;
#pragma unroll 1
    for (int h = warp; h < HC; h += NW) {
        const int u = h * 32 + lane;                  // index into the unit-vector arrays
        if (u < 2 * QPTS) {
            const int t = u < QPTS ? u : u - QPTS + 3 * PS;   // CC patch, then FF patch
            const double lat = S.phi(t), lon = S.lam[t];
            bool bad = false;
            Vec3 v = lat_lon_to_cartesian_spec(lat, lon, bad);
            if (bad) v = lat_lon_to_cartesian_ool(lat, lon);
            const int us = u < QPTS ? u : u - QPTS + PS;
            S.uxy[us] = make_double2(v.x, v.y);
            S.uz[us] = v.z;
            const double phr = lat * D2R;
            S.set_point(t, 0.5 * phr, m::cos_rad_inl(phr));   // the haversines want deg2rad(phi)/2
        }
    }
    if (W0 == 0 || warp >= W0) {
#pragma unroll 1
        for (int l = (W0 == 0 ? warp : warp - W0); l < LC; l += NLW) {
            const int v = l * 32 + lane;              // FC patch, then CF patch
            const int t = v < QPTS ? PS + v : 2 * PS + (v - QPTS);
            if (v < 2 * QPTS) {
                const double phr = S.phi(t) * D2R;
                S.set_point(t, 0.5 * phr, m::cos_rad_inl(phr));
            }
        }
    }
}

// ---- cell stage, speculative straight-line form: the 8 haversines and 2 quadrilaterals are
// evaluated through their hot sequences in ONE basic block (branch-free sqrt / division, short
// atan), so ptxas interleaves the independent chains; `bad` collects every violated
// precondition (spacing above ~1.8 degrees, pole neighbourhoods, degenerate cells, NaN).  A bad
// cell is recomputed through the general functions below; its speculative stores are
// overwritten by the same thread.
//   o        element offset of the cell in the window arrays
//   shC/shF  TMA staging only: a patch whose first column is odd sits one column to the right in its box
__device__ __forceinline__ Vec3 unit_vector(double2 xy, double z) { return Vec3{ xy.x, xy.y, z }; }

template <typename FT, bool TMA>
__device__ __forceinline__ void tile_cell_stage(const BuildArgs& a, const TileSmem<TMA>& S, int tx, int ty, size_t o, int shC, int shF)
{
    constexpr int QX = TileGeom<TMA>::QX, PS = TileGeom<TMA>::PS;
    const double2* const s_uxy = S.uxy; double* const s_uz = S.uz;
    const double R = a.radius;
    // patch index of this cell's own point per staggering (index = loc)
    const int own_cc = (ty + 1) * QX + tx + 1 + shC, own_fc = (ty + 1) * QX + tx + shF;
    const int own_cf = ty * QX + tx + 1 + shC, own_ff = ty * QX + tx + shF;

    const double R2 = 2 * R, RR = R * R;
    SpecRange rg;
#define PT(L, p) S.point((L) * PS + (p))
    {
        const double dx_fc = haversine_spec(PT(CC, own_cc), PT(CC, own_cc - 1), R2, rg);        // utils.jl:14
        const double dy_fc = haversine_spec(PT(FF, own_ff + QX), PT(FF, own_ff), R2, rg);       // :19
        store_metric<FT>(a, DX_FC, o, dx_fc);
        store_metric<FT>(a, DY_FC, o, dy_fc);
        store_metric<FT>(a, AZ_FC, o, dy_fc * dx_fc);                                            // :34
        const double dy_cf = haversine_spec(PT(CC, own_cc), PT(CC, own_cc - QX), R2, rg);       // :20
        const double dx_cf = haversine_spec(PT(FF, own_ff + 1), PT(FF, own_ff), R2, rg);        // :15
        store_metric<FT>(a, DY_CF, o, dy_cf);
        store_metric<FT>(a, DX_CF, o, dx_cf);
        store_metric<FT>(a, AZ_CF, o, dy_cf * dx_cf);                                            // :35
        store_metric<FT>(a, DY_FF, o, haversine_spec(PT(FC, own_fc), PT(FC, own_fc - QX), R2, rg));   // :21
        store_metric<FT>(a, DY_CC, o, haversine_spec(PT(CF, own_cf + QX), PT(CF, own_cf), R2, rg));   // :18
        store_metric<FT>(a, DX_FF, o, haversine_spec(PT(CF, own_cf), PT(CF, own_cf - 1), R2, rg));    // :16
        store_metric<FT>(a, DX_CC, o, haversine_spec(PT(FC, own_fc + 1), PT(FC, own_fc), R2, rg));    // :13
    }
#define UV(q, p) unit_vector(s_uxy[(q) * PS + (p)], s_uz[(q) * PS + (p)])
    {   // az_ff from CC[i-1,j-1], CC[i,j-1], CC[i,j], CC[i-1,j] (:38-43); az_cc from FF[i,j], FF[i+1,j], FF[i+1,j+1], FF[i,j+1] (:23-28)
        const int b0 = own_cc - QX - 1, b1 = own_ff;
        store_metric<FT>(a, AZ_FF, o, spherical_area_quadrilateral_spec(UV(0, b0), UV(0, b0 + 1), UV(0, b0 + QX + 1), UV(0, b0 + QX), rg) * RR);
        store_metric<FT>(a, AZ_CC, o, spherical_area_quadrilateral_spec(UV(1, b1), UV(1, b1 + 1), UV(1, b1 + QX + 1), UV(1, b1 + QX), rg) * RR);
    }
    if (!rg.violated()) return;

    // ---- general path (rare): rolled, out-of-line functions
    //   it 0: dx_fc = hav(CC[i,j], CC[i-1,j])   dy_fc = hav(FF[i,j+1], FF[i,j])   az_fc = dy_fc * dx_fc
    //   it 1: dy_cf = hav(CC[i,j], CC[i,j-1])   dx_cf = hav(FF[i+1,j], FF[i,j])   az_cf = dy_cf * dx_cf
    //   it 2: dy_ff = hav(FC[i,j], FC[i,j-1])   dy_cc = hav(CF[i,j+1], CF[i,j])
    //   it 3: dx_ff = hav(CF[i,j], CF[i-1,j])   dx_cc = hav(FC[i+1,j], FC[i,j])
#pragma unroll 1
    for (int it = 0; it < 4; ++it) {
        const int locA = (it < 2) ? CC : (it == 2) ? FC : CF;
        const int ownA = (it < 2) ? own_cc : (it == 2) ? own_fc : own_cf;
        const int stepA = (it == 0 || it == 3) ? 1 : QX;
        const int locB = (it < 2) ? FF : (it == 2) ? CF : FC;
        const int ownB = (it < 2) ? own_ff : (it == 2) ? own_cf : own_fc;
        const int stepB = (it == 0 || it == 2) ? QX : 1;
        const double hA = haversine_pre_ool(PT(locA, ownA), PT(locA, ownA - stepA), R);
        const double hB = haversine_pre_ool(PT(locB, ownB + stepB), PT(locB, ownB), R);
        const int idA = (it == 0) ? DX_FC : (it == 1) ? DY_CF : (it == 2) ? DY_FF : DX_FF;
        const int idB = (it == 0) ? DY_FC : (it == 1) ? DX_CF : (it == 2) ? DY_CC : DX_CC;
        store_metric<FT>(a, idA, o, hA);
        store_metric<FT>(a, idB, o, hB);
        if (it < 2) store_metric<FT>(a, it == 0 ? AZ_FC : AZ_CF, o, it == 0 ? hB * hA : hA * hB);
    }
#pragma unroll 1
    for (int q = 0; q < 2; ++q) {
        const int b = q ? own_ff : own_cc - QX - 1;                   // lower-left corner of the quadrilateral
        store_metric<FT>(a, q ? AZ_CC : AZ_FF, o,
                         spherical_area_quadrilateral_ool(UV(q, b), UV(q, b + 1), UV(q, b + QX + 1), UV(q, b + QX)) * RR);
    }
#undef PT
#undef UV
}

// ------------------------------------------------------------------ K2f (frame)
// Coordinate accessor of the frame kernel: read the window's coordinate arrays where the point lies
// inside the window, otherwise evaluate the coordinate at its source from the tables (the two
// give the same bits; the second is needed on the ring of a tile window, whose stencil leaves it).
struct FrameAcc {
    const double* const* c; Window w; Tables tb; GridDims g; double fpl;
    __device__ void operator()(int loc, int I, int J, double& lam, double& phi) const
    {
        const int iw = I - w.I0, jw = J - w.J0;
        if (iw >= 0 && iw < w.wnx && jw >= 0 && jw < w.wny) {
            const size_t o = (size_t)jw * w.wnx + iw;
            lam = __ldg(c[loc] + o);
            phi = __ldg(c[4 + loc] + o);
        } else {
            coord_at(tb, g, fpl, loc, I, J, lam, phi);
        }
    }
};

// frame = window minus fast rectangle, enumerated as four rectangles
struct Frame {
    int jlo, jhi, ilo, ihi;     // fast rectangle in WINDOW coordinates: [ilo, ihi) x [jlo, jhi)
    long long n_south, n_north, n_west, n_east;
};

__host__ __device__ inline Frame make_frame(const Window& w, const FastRect& fr)
{
    Frame f;
    if (fr.nx <= 0 || fr.ny <= 0) { f.ilo = f.ihi = 0; f.jlo = f.jhi = w.wny; }
    else { f.ilo = fr.I0 - w.I0; f.ihi = f.ilo + fr.nx; f.jlo = fr.J0 - w.J0; f.jhi = f.jlo + fr.ny; }
    f.n_south = (long long)f.jlo * w.wnx;
    f.n_north = (long long)(w.wny - f.jhi) * w.wnx;
    f.n_west = (long long)(f.jhi - f.jlo) * f.ilo;
    f.n_east = (long long)(f.jhi - f.jlo) * (w.wnx - f.ihi);
    return f;
}

// One thread per (frame element, metric), metric-major: a warp serves 32 consecutive elements of
// ONE metric, evaluates it at its class's M2 source and stores it coalesced.  (The former
// one-thread-per-element form evaluated up to four sources and twelve metrics in sequence: a 40 us
// latency chain holding 128 registers next to the tile kernel.)  t = mtr * n + element.
template <typename FT>
__device__ __noinline__ void frame_element(const BuildArgs& a, const Frame& f, long long t)
{
    const long long n = f.n_south + f.n_north + f.n_west + f.n_east;
    const Window w = a.w;
    int mtr, iw, jw;
    if (N_METRIC * n < 0x7fffffffLL) {
        // (every realistic frame: 32-bit index arithmetic -- a 64-bit division costs ~100 instructions, and the decode
        // would be a third of this function)
        const unsigned n32 = (unsigned)n, ns = (unsigned)f.n_south, nn = (unsigned)f.n_north, nw = (unsigned)f.n_west;
        unsigned u = (unsigned)t;
        mtr = (int)(u / n32);
        u -= (unsigned)mtr * n32;
        if (u < ns) { jw = (int)(u / (unsigned)w.wnx); iw = (int)(u - (unsigned)jw * (unsigned)w.wnx); }
        else if ((u -= ns) < nn) { jw = (int)(u / (unsigned)w.wnx); iw = (int)(u - (unsigned)jw * (unsigned)w.wnx); jw += f.jhi; }
        else if ((u -= nn) < nw) { jw = (int)(u / (unsigned)f.ilo); iw = (int)(u - (unsigned)jw * (unsigned)f.ilo); jw += f.jlo; }
        else {
            u -= nw;
            const unsigned ne = (unsigned)(w.wnx - f.ihi);
            jw = (int)(u / ne); iw = (int)(u - (unsigned)jw * ne) + f.ihi; jw += f.jlo;
        }
    } else {
        mtr = (int)(t / n);
        t -= (long long)mtr * n;
        if (t < f.n_south) { jw = (int)(t / w.wnx); iw = (int)(t - (long long)jw * w.wnx); }
        else if ((t -= f.n_south) < f.n_north) { jw = (int)(t / w.wnx); iw = (int)(t - (long long)jw * w.wnx); jw += f.jhi; }
        else if ((t -= f.n_north) < f.n_west) { jw = (int)(t / f.ilo); iw = (int)(t - (long long)jw * f.ilo); jw += f.jlo; }
        else {
            t -= f.n_west;
            const int ne = w.wnx - f.ihi;
            jw = (int)(t / ne); iw = (int)(t - (long long)jw * ne) + f.ihi; jw += f.jlo;
        }
    }
    const GridDims g = a.g;
    const Tables tb = make_tables(a.tables, g.Nx, g.Ny);
    const int I = w.I0 + iw, J = w.J0 + jw;
    const int loc = metric_loc(mtr);
    double v;
    if (J <= 1) {                                            // M3: lat-lon continuation rows
        const double* row = tb.ll + (size_t)(J + g.Hy - 1) * LL_NCOL;
        const int col = mtr == DX_CC ? LL_DX_CC : mtr == DX_FC ? LL_DX_FC : mtr == DX_CF ? LL_DX_CF : mtr == DX_FF ? LL_DX_FF
                      : mtr == AZ_CC ? LL_AZ_CC : mtr == AZ_FC ? LL_AZ_FC : mtr == AZ_CF ? LL_AZ_CF : mtr == AZ_FF ? LL_AZ_FF : LL_DY;
        v = row[col];
    } else {                                                 // M1 at the class's M2 source
        int si, sj; bool wrapped;
        halo_source(loc_x(loc), loc_y(loc), I, J, g.Nx, g.Ny, si, sj, wrapped);
        const FrameAcc acc{ a.coords64, w, tb, g, a.first_pole_longitude };
        v = metric_one(acc, mtr, si, sj, a.radius);
    }
    static_cast<FT*>(a.metrics_out[mtr])[(size_t)jw * w.wnx + iw] = (FT)v;
}

// The frame on its own grid (see launch_metrics for when).
template <typename FT>
__global__ void __launch_bounds__(128) metrics_frame_kernel(const __grid_constant__ BuildArgs a, const Frame f, const int sync_first)
{
    // sync_first = 0 (launched with programmatic serialisation behind the TILE kernel, every block of which has waited
    // for the coordinates before it let this grid go): compute at once, next to the tile kernel's last wave, and wait for
    // that grid LAST -- this grid must not complete before it, the stream's next operation depends on both.
    if (sync_first) cudaGridDependencySynchronize();
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < N_METRIC * (f.n_south + f.n_north + f.n_west + f.n_east)) frame_element<FT>(a, f, t);
    cudaGridDependencySynchronize();
}

// RIDE: the halo frame (M2/M3, 0.55 % of the elements, one long latency chain per element) rides in this grid:
// block columns >= `tile_cols` of the first `frame.rows` block rows are frame blocks, so they are scheduled
// interleaved with the tile blocks (their latency hides behind the tiles' arithmetic, as on a second stream) at the
// cost of neither a launch nor an event pair.  RIDE = false: the frame runs on its own grid (launch_metrics).
struct FrameRide { Frame frame; int tile_cols, cols, rows; };

template <typename FT, bool TMA, bool RIDE>
__global__ void __launch_bounds__(NT, TILE_CTAS) metrics_tile_kernel(const __grid_constant__ BuildArgs a, const FastRect fr,
                                                             const __grid_constant__ TileMaps tm, const __grid_constant__ FrameRide ride)
{
    // Launched with programmatic stream serialisation behind the coordinate kernel: the block may be resident while
    // that grid drains.  A tile needs the ten coordinate rows of its patches only: with the row counters on
    // (a.rows_done) it waits for those, so the first tile rows run next to the coordinate kernel's last blocks;
    // otherwise, and in every frame block (whose sources lie anywhere), it waits for that grid's completion and
    // flush.  (The frame blocks' wait also makes this grid's completion imply the coordinate grid's.)
    const bool row_deps = RIDE && a.rows_done != nullptr && (int)blockIdx.x < ride.tile_cols;
    if (!row_deps) cudaGridDependencySynchronize();
    if constexpr (RIDE) {
        if ((int)blockIdx.x >= ride.tile_cols) {
            if ((int)blockIdx.y < ride.rows) {
                const long long t = ((long long)blockIdx.y * ride.cols + ((int)blockIdx.x - ride.tile_cols)) * NT + threadIdx.y * TX + threadIdx.x;
                const Frame& f = ride.frame;
                if (t < N_METRIC * (f.n_south + f.n_north + f.n_west + f.n_east)) frame_element<FT>(a, f, t);
            }
            return;
        }
    }
    // block rows from the north: the rows next to the poles (cells that leave the speculative path) are the slowest
    // tiles, scheduled first they run next to ordinary tiles instead of forming the grid's tail
    const int by = (int)(gridDim.y - 1 - blockIdx.y);
    constexpr int QPTS = TileGeom<TMA>::QPTS, PS = TileGeom<TMA>::PS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* smem = reinterpret_cast<double*>(TMA ? (reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127
                                                 : reinterpret_cast<uintptr_t>(smem_raw));
    const TileSmem<TMA> S(smem);

    const int tid = threadIdx.y * TX + threadIdx.x;
    const int It0 = fr.I0 + blockIdx.x * TX, Jt0 = fr.J0 + by * TY;   // global index of tile cell (0,0)
    const Window w = a.w;
    if (row_deps) {
        // patch rows of the four staggerings: window rows Jt0 - 1 .. Jt0 + TY (clipped); one thread per row polls its
        // counter with acquire semantics, the barrier hands the visibility to the block.  A counter that does not
        // arrive (it always does: every coordinate block is resident or done when this grid starts) ends in the
        // grid-wide wait instead of a hang.
        const int tid0 = threadIdx.y * TX + threadIdx.x;
        if (tid0 < TY + 2) {
            const int r = min(max(Jt0 - 1 - w.J0 + tid0, 0), w.wny - 1);
            const unsigned int* c = a.rows_done + r;
            unsigned v, spins = 0;
            do {                                                  // relaxed polls (an acquire load invalidates the SM's L1 every time)
                asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(c) : "memory");
            } while (v < a.rows_target && ++spins < (1u << 22));
            if (v < a.rows_target) cudaGridDependencySynchronize();
            else asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(c) : "memory");
        }
        __syncthreads();
    }
    cudaTriggerProgrammaticLaunchCompletion();

    if constexpr (TMA) {
        // ---- stage the four coordinate patches with the Tensor Memory Accelerator: one thread issues
        // eight 2-D box loads ((TX+2) x QY doubles each, zero-filled outside the window) that complete
        // on an mbarrier; no address arithmetic, no load instructions in the other 255 threads.
        __shared__ __align__(8) unsigned long long s_bar;
        if (tid == 0) mbar_init(&s_bar, 1);
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(&s_bar, 8u * QPTS * (unsigned)sizeof(double));
#pragma unroll
            for (int loc = 0; loc < 4; ++loc) {
                // the box must start on a 16-byte boundary of the row (measured: an odd FP64 column
                // raises an illegal-instruction fault): start at the even column at or below the
                // patch; the (TX+2)-wide box still covers the TX+1 columns the patch needs
                const int x0 = (It0 + patch_di(loc) - w.I0) & ~1, y0 = Jt0 + patch_dj(loc) - w.J0;
                tma_load_2d(&S.lam[loc * PS], &tm.m[loc], x0, y0, &s_bar);
                tma_load_2d(&S.phr[loc * PS], &tm.m[4 + loc], x0, y0, &s_bar);
            }
        }
        mbar_wait(&s_bar, 0);
    } else {
        tile_stage_cp_async(a, w, It0, Jt0, tid, S);
        __syncthreads();
    }

    tile_point_stage<TMA>(S, tid);
    __syncthreads();

    const int ci = blockIdx.x * TX + threadIdx.x, cj = by * TY + threadIdx.y;
    if (ci >= fr.nx || cj >= fr.ny) return;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const size_t o = (size_t)(Jt0 + ty - w.J0) * w.wnx + (It0 + tx - w.I0);
    const int shC = TMA ? ((It0 - 1 - w.I0) & 1) : 0, shF = TMA ? ((It0 - w.I0) & 1) : 0;
    tile_cell_stage<FT, TMA>(a, S, tx, ty, o, shC, shF);
}

// ------------------------------------------------------------------ K2 (mirror pairs; opt-in, OSSG_TILE_PAIRS=1)
// The map is mirror symmetric about the pole meridians: with c = Nx/2 + 1 the x-Center points of
// post-shift columns i and c - i (mod Nx) and the x-Face points of i and c + 1 - i carry the same
// latitude bit for bit (generate_tripolar_coordinates.jl:67-78: x changes sign, y does not), in every
// row including the fold row Ny.  About 70 % of the point stage depends on the latitude only
// (sind/cosd(phi) of the unit vectors, cos(deg2rad(phi)) of the haversines), so a tile A and its
// mirror tile B need those values only once.  The two tiles form a thread-block cluster of two CTAs:
// each computes the latitude part for its own lower half of the shared patch columns and stores the
// results into BOTH shared memories (distributed shared memory), after checking that the partner
// really holds the same latitude bits (ossg_tripolar_metrics accepts any coordinates); a mismatch --
// or a point outside the straight-line preconditions -- raises a flag and the CTA falls back to the
// ordinary point stage.  The longitude part, the products and the cell stage are unchanged.
//   x-Center patches (CC, CF: columns It0-1 .. It0+nt-1 -> px = 0 .. nt):  px <-> nt + 1 - px  (px = 0 has no partner)
//   x-Face   patches (FF, FC: columns It0   .. It0+nt   -> px = 0 .. nt):  px <-> nt - px
// Requires a window that spans x and Nx % 4 == 0 (each quarter of the columns is a run of A tiles
// whose mirrors fill the next quarter).
__device__ __forceinline__ long long f64_bits(double x) { return __double_as_longlong(x); }

template <typename FT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NT, TILE_CTAS)
metrics_tile_pair_kernel(const __grid_constant__ BuildArgs a, const FastRect fr)
{
    constexpr int QX = TileGeom<false>::QX, QPTS = TileGeom<false>::QPTS, PS = TileGeom<false>::PS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const TileSmem<false> S(reinterpret_cast<double*>(smem_raw));
    __shared__ int s_flag;

    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int member = (int)cluster.block_rank();             // 0: tile A, 1: its mirror tile B
    const int tid = threadIdx.y * TX + threadIdx.x;
    const int Nx = a.g.Nx, q4 = Nx / 4;
    const int nts = (q4 + TX - 1) / TX;                       // A tiles per quarter
    const int pair = blockIdx.x >> 1;
    const int seg = pair / nts, k = pair - seg * nts;
    const int acol = seg * (Nx / 2) + 1 + k * TX;             // first column of tile A
    const int nt = min(TX, q4 - k * TX);                      // columns of this pair's tiles
    const int bcol = seg * Nx + Nx / 2 + 1 - (acol + nt - 1); // first column of tile B: column acol+t <-> bcol+nt-1-t
    const int It0 = member ? bcol : acol, Jt0 = fr.J0 + blockIdx.y * TY;
    const Window w = a.w;

    if (tid == 0) s_flag = 0;
    tile_stage_cp_async(a, w, It0, Jt0, tid, S);
    cluster.sync();                                           // both patches staged, both flags cleared

    // ---- latitude part.  This CTA computes columns 0 .. hC of its x-Center patches and 0 .. gF of its
    // x-Face patches (in its own indexing); the partner computes the rest and stores it here.
    {
        const int hC = member ? nt / 2 : (nt + 1) / 2;
        const int gF = member ? (nt - 1) / 2 : nt / 2;
        const int nC = hC + 1, nF = gF + 1;
        const int ntask = (nC + nF) * QY;                     // heavy tasks (CC, FF); as many light ones (CF, FC)
        const unsigned other = (unsigned)(member ^ 1);
        TileSmem<false> R = S;                                // the partner's arrays (same layout)
        R.phr = cluster.map_shared_rank(S.phr, other);
        R.cosp = cluster.map_shared_rank(S.cosp, other);
        double2* const r_uxy = cluster.map_shared_rank(S.uxy, other);
        double* const r_uz = cluster.map_shared_rank(S.uz, other);
        int* const r_flag = cluster.map_shared_rank(&s_flag, other);
        // task u -> patch point: x-Center patch first (nC columns per row), then the x-Face patch
        auto locate = [&](int u, bool& face, int& idx, int& idxr, bool& shared) {
            face = u >= nC * QY;
            const int v = face ? u - nC * QY : u, ncol = face ? nF : nC;
            const int py = v / ncol, px = v - py * ncol;
            idx = py * QX + px;
            shared = face || px >= 1;
            idxr = py * QX + (face ? nt - px : nt + 1 - px);
        };
        const int NW = NT / 32, nchunk = (ntask + 31) / 32, W0 = nchunk % NW;
        const int NLW = (W0 == 0) ? NW : NW - W0;
        const int warp = tid >> 5, lane = tid & 31;
#pragma unroll 1
        for (int h = warp; h < nchunk; h += NW) {             // heavy: sind/cosd(phi) and cos(deg2rad(phi)) of a CC / FF point
            const int u = h * 32 + lane;
            if (u < ntask) {
                bool face, shared; int idx, idxr;
                locate(u, face, idx, idxr, shared);
                const int t = (face ? FF : CC) * PS + idx, us = (face ? PS : 0) + idx;
                const double lat = S.phi(t);
                bool bad = false;
                double sn, cs;
                m::sincosd_lat_spec(lat, sn, cs, bad);
                const double cphi = m::cos_rad_inl(lat * D2R);
                S.uz[us] = sn;
                S.uxy[us].x = cs;                             // (cosd(phi) until the longitude part multiplies it)
                S.cosphi(t) = cphi;
                if (bad) s_flag = 1;
                if (shared) {
                    const int tr = (face ? FF : CC) * PS + idxr, ur = (face ? PS : 0) + idxr;
                    if (!bad && f64_bits(R.phi(tr)) == f64_bits(lat)) { r_uz[ur] = sn; r_uxy[ur].x = cs; R.cosphi(tr) = cphi; }
                    else *r_flag = 1;
                }
            }
        }
        if (W0 == 0 || warp >= W0) {
#pragma unroll 1
            for (int l = (W0 == 0 ? warp : warp - W0); l < nchunk; l += NLW) {   // light: cos(deg2rad(phi)) of a CF / FC point
                const int u = l * 32 + lane;
                if (u < ntask) {
                    bool face, shared; int idx, idxr;
                    locate(u, face, idx, idxr, shared);
                    const int t = (face ? FC : CF) * PS + idx;
                    const double lat = S.phi(t);
                    const double cphi = m::cos_rad_inl(lat * D2R);
                    S.cosphi(t) = cphi;
                    if (shared) {
                        const int tr = (face ? FC : CF) * PS + idxr;
                        if (f64_bits(R.phi(tr)) == f64_bits(lat)) R.cosphi(tr) = cphi;
                        else *r_flag = 1;
                    }
                }
            }
        }
    }
    cluster.sync();                                           // the partner's half has arrived; no remote access after this point

    // ---- longitude part and products, every patch point of this CTA
#pragma unroll 1
    for (int u = tid; u < 2 * QPTS; u += NT) {
        const int t = u < QPTS ? u : u - QPTS + 3 * PS;       // CC patch, then FF patch
        const int us = u < QPTS ? u : u - QPTS + PS;
        const int t2 = u < QPTS ? PS + u : 2 * PS + (u - QPTS);   // the FC / CF point with the same index
        bool bad = false;
        double sl, cl;
        m::sincosd_lon_spec(S.lam[t], sl, cl, bad);
        const double cs = S.uxy[us].x;
        S.uxy[us] = make_double2(cl * cs, sl * cs);
        S.phi(t) = 0.5 * (S.phi(t) * D2R);                    // the haversines want deg2rad(phi)/2
        S.phi(t2) = 0.5 * (S.phi(t2) * D2R);
        if (bad) s_flag = 1;
    }
    __syncthreads();
    if (s_flag) {                                             // (block-uniform) the ordinary point stage on a fresh patch
        tile_stage_cp_async(a, w, It0, Jt0, tid, S);
        __syncthreads();
        tile_point_stage<false>(S, tid);
        __syncthreads();
    }

    const int tx = threadIdx.x, ty = threadIdx.y;
    if (tx >= nt || blockIdx.y * TY + ty >= fr.ny) return;
    const size_t o = (size_t)(Jt0 + ty - w.J0) * w.wnx + (It0 + tx - w.I0);
    tile_cell_stage<FT, false>(a, S, tx, ty, o, 0, 0);
}

// ------------------------------------------------------------------ host side
static int check_params(const ossg_tripolar_params* p, ossg_tripolar_params& q)
{
    if (!p) return set_error(OSSG_EINVAL, "params is NULL");
    q = *p;
    if (q.Nx < 2 || q.Ny < 2) return set_error(OSSG_EINVAL, "size must be at least (2, 2), got (%d, %d)", q.Nx, q.Ny);
    if (q.Nx % 2 != 0)
        return set_error(OSSG_EODD_NX, "The number of cells in the longitude dimension should be even!");
    if (q.Hx < 1 || q.Hy < 1) return set_error(OSSG_EHALO, "halo must be at least (1, 1), got (%d, %d)", q.Hx, q.Hy);
    if (q.Hy > q.Ny - 1 || q.Hx > q.Nx)
        return set_error(OSSG_EHALO, "halo (%d, %d) too large for size (%d, %d): need Hx <= Nx, Hy <= Ny-1", q.Hx, q.Hy, q.Nx, q.Ny);
    if (q.nx == 0) { q.i0 = 1; q.nx = q.Nx; }
    if (q.ny == 0) { q.j0 = 1; q.ny = q.Ny; }
    if (q.i0 < 1 || q.nx < 1 || q.i0 + q.nx - 1 > q.Nx || q.j0 < 1 || q.ny < 1 || q.j0 + q.ny - 1 > q.Ny)
        return set_error(OSSG_EWINDOW, "tile i0=%d nx=%d j0=%d ny=%d outside the %dx%d grid", q.i0, q.nx, q.j0, q.ny, q.Nx, q.Ny);
    if (!(q.radius > 0.0)) return set_error(OSSG_EINVAL, "radius must be positive");
    if (q.flags & ~OSSG_LAMBDA1D_F32) return set_error(OSSG_EINVAL, "unknown flags 0x%x", q.flags);
    return 0;
}

static size_t table_bytes(const ossg_tripolar_params& q)
{
    size_t n = TABLE_HDR_WORDS + tables_doubles(q.Nx, q.Ny, q.Hy);
    n = (n + 1) & ~(size_t)1;            // keep what follows 16-byte aligned
    return n * sizeof(double);
}
static size_t window_elems(const ossg_tripolar_params& q)
{
    return (size_t)(q.nx + 2 * q.Hx) * (size_t)(q.ny + 2 * q.Hy);
}
// workspace layout: [header + tables | FP64 coordinates of a Float32 build | one row counter per window row]
static size_t row_counter_offset(const ossg_tripolar_params& q, int out_type)
{
    size_t b = table_bytes(q);
    if (out_type == OSSG_F32) b += 8 * window_elems(q) * sizeof(double);
    return b;
}
static size_t row_counter_bytes(const ossg_tripolar_params& q)
{
    return (((size_t)(q.ny + 2 * q.Hy) * sizeof(unsigned int)) + 15) & ~(size_t)15;
}
// Row-band dependency between the coordinate and the tile kernel: opt-in with OSSG_ROW_DEPS=1 (read once per process).
// Measured slower than the grid-wide wait: 0.8325 against 0.8233 ms per build of the 1/12 degree grid, 0.1353 against
// 0.1339 ms on the slowest slab of 8 ranks (profiles/r02m_rowdeps.txt) -- the poll, its L1 invalidation and the extra
// barrier at the head of every tile cost more than the few microseconds of overlap at the kernel boundary give back.
static bool row_deps_enabled()
{
    static const bool on = [] { const char* e = getenv("OSSG_ROW_DEPS"); return e && strcmp(e, "1") == 0; }();
    return on;
}

static void fill_common(BuildArgs& a, const ossg_tripolar_params& q, void* workspace)
{
    a.g = GridDims{ q.Nx, q.Ny, q.Hx, q.Hy };
    a.flags = q.flags;
    a.w = Window{ q.i0 - q.Hx, q.j0 - q.Hy, q.nx + 2 * q.Hx, q.ny + 2 * q.Hy };
    a.south = q.southernmost_latitude;
    a.north_poles_latitude = q.north_poles_latitude;
    a.first_pole_longitude = q.first_pole_longitude;
    a.radius = q.radius;
    a.stamp = static_cast<unsigned long long*>(workspace);
    a.tables = static_cast<double*>(workspace) + TABLE_HDR_WORDS;
    a.rows_done = nullptr;
    a.rows_target = 0;
}

// Launch with programmatic stream serialisation: the kernel may start while its predecessor in the stream (the
// table kernel, which triggers its dependents at once) is still running; it synchronises on that grid itself.
// The table kernel is a stamp check in the steady state, so this removes one full kernel boundary per build.
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl_smem(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, bool programmatic, cudaStream_t s, Args&&... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = programmatic ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t s, Args&&... args)
{
    return launch_pdl_smem(kernel, grid, block, 0, true, s, std::forward<Args>(args)...);
}

static int launch_tables(const BuildArgs& a, cudaStream_t s)
{
    const int n = 2 * a.g.Nx + 2 * a.g.Ny + a.g.Hy + 1;
    OSSG_CUDA_TRY(launch_pdl(tables_kernel, dim3((unsigned)ceil_div(max(n, a.w.wny), 128)), dim3(128), s, a.g, a.flags, a.south, a.north_poles_latitude,
                             a.radius, a.tables, a.stamp, a.rows_done, a.w.wny));
    OSSG_LAUNCH_CHECK("tables_kernel");
    return 0;
}

template <typename FT>
static int launch_coords(BuildArgs& a, cudaStream_t s)        // sets a.rows_target (what a complete row counts to)
{
    dim3 block(256), grid(ceil_div(a.w.wnx, 256), a.w.wny);
    const bool full_x = a.w.I0 == 1 - a.g.Hx && a.w.wnx == a.g.Nx + 2 * a.g.Hx;
    bool split_x = !full_x && a.w.wnx <= a.g.Nx;               // no column has two images inside the window
    PairGrid pg;
    memset(&pg, 0, sizeof(pg));
    const int Nx = a.g.Nx, half = Nx / 2;
    if (full_x) {
        pg.t_lo = 0; pg.n_t = half + 1; pg.n_halo = 2 * a.g.Hx;
    } else if (split_x) {
        // bounding range of the pair threads with an image inside the window (kernel's own maps, on the host)
        auto inside = [&](int c) { int wc = c - a.w.I0; if (wc >= a.w.wnx) wc -= Nx; if (wc < 0) wc += Nx; return wc >= 0 && wc < a.w.wnx; };
        auto shifted = [&](int ip) { int i = ip + Nx / 4; return i > Nx ? i - Nx : i; };
        int lo = half + 1, hi = -1;
        for (int t = 0; t <= half; ++t) {
            const int ip = 1 + t;
            const bool any = inside(shifted(ip)) || (t != 0 && t != half && inside(shifted(Nx + 2 - ip))) || (t < half && inside(shifted(Nx + 1 - ip)));
            if (any) { lo = min(lo, t); hi = max(hi, t); }
        }
        pg.t_lo = hi < 0 ? 0 : lo; pg.n_t = hi < 0 ? 0 : hi - lo + 1; pg.n_halo = 0;
        // worth it only where evaluations are shared: with one stored image per evaluation (Rx >= 4) the
        // per-element kernel measured 1-3 % faster (profiles/r02j_tiles.txt)
        if (4 * pg.n_t > 3 * a.w.wnx) split_x = false;
    }
    if ((full_x && a.g.Hx <= a.g.Nx / 2) || split_x) {
        pg.jw_lo = max(0, 1 - a.w.J0);
        const int jw_hi = min(a.w.wny - 1, a.g.Ny - 1 - a.w.J0);
        pg.n_pair_rows = max(0, jw_hi - pg.jw_lo + 1);
        if (pg.n_pair_rows == 0) pg.jw_lo = 0;
        pg.chunks_pair = max(1, ceil_div(pg.n_t + pg.n_halo, OSSG_COORD_BLOCK));
        pg.chunks_full = ceil_div(a.w.wnx, OSSG_COORD_BLOCK);
        a.rows_target = (unsigned)pg.chunks_pair * (unsigned)pg.chunks_full;
        const long long nb = (long long)pg.n_pair_rows * pg.chunks_pair + (long long)(a.w.wny - pg.n_pair_rows) * pg.chunks_full;
#ifdef OSSG_COORD_MAGIC_DIV
        pg.magic_pair = ((1ull << 40) + pg.chunks_pair - 1) / pg.chunks_pair;     // (nb < 2^31 and chunks < 2^9 keep b * chunks below 2^40)
        pg.magic_full = ((1ull << 40) + pg.chunks_full - 1) / pg.chunks_full;
#endif
        if (split_x) OSSG_CUDA_TRY(launch_pdl(coords_pair_kernel<FT, true>, dim3((unsigned)nb), dim3(OSSG_COORD_BLOCK), s, a, pg));
        else OSSG_CUDA_TRY(launch_pdl(coords_pair_kernel<FT, false>, dim3((unsigned)nb), dim3(OSSG_COORD_BLOCK), s, a, pg));
        OSSG_LAUNCH_CHECK("coords_pair_kernel");
    } else {
        a.rows_target = grid.x;
        OSSG_CUDA_TRY(launch_pdl(coords_kernel<FT>, grid, block, s, a));
        OSSG_LAUNCH_CHECK("coords_kernel");
    }
    return 0;
}

// Fork/join resources of the current device: the halo-frame kernel depends on the coordinates only,
// so it runs on a side stream next to the tile kernel instead of after it.  One {stream, fork event,
// join event} set per device, created on first use and reused by every call (no per-call
// cudaEventCreate/Destroy).  The mutex is held from the fork record to the join wait: an event
// record/wait pair captures the event's state at the time of the call, so calls from several host
// threads (each on its own stream) may share the set as long as their pairs do not interleave.
struct ForkJoin {
    cudaStream_t side = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};
static std::mutex g_fork_mu;
static ForkJoin g_fork[64];

static int fork_join(ForkJoin** out)                       // call with g_fork_mu held
{
    int dev = 0;
    OSSG_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return set_error(OSSG_EINVAL, "device ordinal %d not supported", dev);
    ForkJoin& fj = g_fork[dev];
    if (!fj.side) OSSG_CUDA_TRY(cudaStreamCreateWithFlags(&fj.side, cudaStreamNonBlocking));
    if (!fj.fork) OSSG_CUDA_TRY(cudaEventCreateWithFlags(&fj.fork, cudaEventDisableTiming));
    if (!fj.join) OSSG_CUDA_TRY(cudaEventCreateWithFlags(&fj.join, cudaEventDisableTiming));
    *out = &fj;
    return 0;
}

// the tile kernels ask for more than 48 KB of dynamic shared memory: opt in once per device
static int tile_attributes(int which /*0: double, 1: float*/, int kernel /*0: cp.async, 1: TMA, 2: pairs, 3: cp.async + frame*/, const void* fn, size_t bytes)
{
    static std::mutex mu;
    static unsigned long long done[2][4] = {};
    int dev = 0;
    OSSG_CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    if (dev >= 0 && dev < 64 && ((done[which][kernel] >> dev) & 1ull)) return 0;
    OSSG_CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    if (dev >= 0 && dev < 64) done[which][kernel] |= 1ull << dev;
    return 0;
}

// tile-kernel form, from the environment once per process: 0 = cp.async staging
// (default, the measured best), 1 = TMA box loads (OSSG_TILE_STAGING=tma), 2 = mirror-pair clusters (OSSG_TILE_PAIRS=1)
static int tile_form()
{
    static const int form = [] {
        const char* staging = getenv("OSSG_TILE_STAGING");
        const char* pairs = getenv("OSSG_TILE_PAIRS");
        if (OSSG_USE_TMA || (staging && strcmp(staging, "tma") == 0)) return 1;
        if (pairs && strcmp(pairs, "1") == 0) return 2;
        return 0;
    }();
    return form;
}

// Where the halo frame runs, from the environment once per process (OSSG_FRAME = fork | tail | ride):
//   0 fork  its own grid on a side stream next to the tile kernel (event pair around it)
//   1 tail  its own grid behind the tile kernel with programmatic serialisation: it starts when the tile kernel's last
//           wave is resident and fills the slots that wave leaves
//   2 ride  inside the tile kernel's grid (extra block columns, metrics_tile_kernel<.., RIDE = true>)
static int frame_mode()
{
    static const int mode = [] {
        const char* e = getenv("OSSG_FRAME");
        if (e && strcmp(e, "fork") == 0) return 0;
        if (e && strcmp(e, "tail") == 0) return 1;
        if (e && strcmp(e, "ride") == 0) return 2;
        return OSSG_FRAME_DEFAULT;
    }();
    return mode;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled encode_tiled_fn()
{
    static PFN_encodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(sym);
        else
            (void)cudaGetLastError();
    }
    return fn;
}

// Tensor maps of the eight FP64 coordinate arrays of the window, box (TX+2) x QY.  Returns false
// when TMA cannot describe them (odd row pitch -> stride not a multiple of 16 bytes, unaligned
// base, no driver entry point); the caller then uses the cp.async staging.
static bool make_tile_maps(const BuildArgs& a, TileMaps& tm)
{
    PFN_encodeTiled enc = encode_tiled_fn();
    if (!enc || (a.w.wnx & 1) || a.w.wnx < TileGeom<true>::QX || a.w.wny < QY) return false;
    const cuuint64_t dims[2] = { (cuuint64_t)a.w.wnx, (cuuint64_t)a.w.wny };
    const cuuint64_t strides[1] = { (cuuint64_t)a.w.wnx * sizeof(double) };
    const cuuint32_t box[2] = { (cuuint32_t)TileGeom<true>::QX, (cuuint32_t)QY };
    const cuuint32_t estr[2] = { 1, 1 };
    for (int k = 0; k < 8; ++k) {
        if (reinterpret_cast<uintptr_t>(a.coords64[k]) & 15) return false;
        if (enc(&tm.m[k], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, a.coords64[k], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
    }
    return true;
}

template <typename FT>
static int launch_metrics(const BuildArgs& a, bool whole_grid, cudaStream_t s, cudaEvent_t between = nullptr)
{
    (void)whole_grid;
    const Window& w = a.w;
    // fast rectangle: global interior rows 2..Ny-1, columns 1..Nx, one cell inside the window ring
    FastRect fr;
    const int Ilo = max(1, w.I0 + 1), Ihi = min(a.g.Nx, w.I0 + w.wnx - 2);
    const int Jlo = max(2, w.J0 + 1), Jhi = min(a.g.Ny - 1, w.J0 + w.wny - 2);
    fr.I0 = Ilo; fr.J0 = Jlo; fr.nx = Ihi - Ilo + 1; fr.ny = Jhi - Jlo + 1;
    const bool have_tile = fr.nx > 0 && fr.ny > 0;
    const Frame f = make_frame(w, fr);
    const long long nframe = f.n_south + f.n_north + f.n_west + f.n_east;

    // table kernel -> coordinate kernel -> tile kernel, each launched with programmatic stream serialisation behind the
    // one it depends on; the frame as frame_mode() says.  The profiled build records events between the kernels,
    // which serialises them fully, and keeps the frame on its own grid behind the tile kernel so that the kernel
    // times can be told apart.
    const bool chain = !between;
    const int form = tile_form();
    const bool want_tma = form == 1;
    const bool want_pairs = form == 2 && fr.I0 == 1 && fr.nx == a.g.Nx && a.g.Nx % 4 == 0;
    int mode = (chain && have_tile && nframe > 0) ? frame_mode() : 1;
    if (mode == 2 && (want_pairs || want_tma)) mode = 1;
    if (mode == 1 && want_pairs) mode = 0;               // (the pair kernel does not trigger its dependents early)
    const long long nfb128 = ceil_div((long long)N_METRIC * nframe, 128);
    std::unique_lock<std::mutex> fork_lock(g_fork_mu, std::defer_lock);
    ForkJoin* fj = nullptr;
    if (mode == 0) {
        fork_lock.lock();
        int rc = fork_join(&fj);
        if (rc) return rc;
        OSSG_CUDA_TRY(cudaEventRecord(fj->fork, s));
        OSSG_CUDA_TRY(cudaStreamWaitEvent(fj->side, fj->fork, 0));
        metrics_frame_kernel<FT><<<(unsigned)nfb128, 128, 0, fj->side>>>(a, f, 1);
        OSSG_LAUNCH_CHECK("metrics_frame_kernel");
        OSSG_CUDA_TRY(cudaEventRecord(fj->join, fj->side));
    }
    if (have_tile) {
        const int which = sizeof(FT) == 4;
        dim3 block(TX, TY), grid(ceil_div(fr.nx, TX), ceil_div(fr.ny, TY));
        TileMaps tm;
        FrameRide ride;
        memset(&ride, 0, sizeof(ride));
        // Staging of the coordinate patches: cp.async by default.  The TMA form (eight 2-D box loads
        // issued by one thread) is complete and parity-tested but measured 3 % slower on the 1/12
        // degree grid (0.650 vs 0.631 ms: 3 % more patch points because of the 16-byte box
        // alignment, and the patch arrives later than 256 threads' own loads); the mirror-pair
        // variant (clusters of two CTAs sharing the latitude part of the point stage) measured 17 %
        // slower (0.724 vs 0.621 ms).  Both stay opt-in, see tile_form().
        int rc = 0;
        if (want_pairs) {
            if ((rc = tile_attributes(which, 2, (const void*)metrics_tile_pair_kernel<FT>, TileGeom<false>::SMEM_BYTES))) return rc;
            const int nts = ceil_div(a.g.Nx / 4, TX);
            dim3 pgrid(4 * nts, ceil_div(fr.ny, TY));            // 2 quarters of A tiles x nts pairs x 2 members
            metrics_tile_pair_kernel<FT><<<pgrid, block, TileGeom<false>::SMEM_BYTES, s>>>(a, fr);
        } else if (want_tma && make_tile_maps(a, tm)) {
            if ((rc = tile_attributes(which, 1, (const void*)metrics_tile_kernel<FT, true, false>, TileGeom<true>::SMEM_BYTES))) return rc;
            OSSG_CUDA_TRY(launch_pdl_smem(metrics_tile_kernel<FT, true, false>, grid, block, TileGeom<true>::SMEM_BYTES, chain, s, a, fr, tm, ride));
        } else if (mode == 2) {
            // frame blocks in the first 3/4 of the block rows only: a frame block is one ~25 us latency chain, in the
            // last rows it would outlast the tile blocks next to it
            const long long nfb = ceil_div((long long)N_METRIC * nframe, (long long)NT);
            ride.frame = f;
            ride.tile_cols = (int)grid.x;
            ride.rows = max(1, (int)(grid.y * 3 / 4));
            ride.cols = (int)ceil_div(nfb, (long long)ride.rows);
            grid.x += ride.cols;
            if ((rc = tile_attributes(which, 3, (const void*)metrics_tile_kernel<FT, false, true>, TileGeom<false>::SMEM_BYTES))) return rc;
            memset(&tm, 0, sizeof(tm));
            OSSG_CUDA_TRY(launch_pdl_smem(metrics_tile_kernel<FT, false, true>, grid, block, TileGeom<false>::SMEM_BYTES, chain, s, a, fr, tm, ride));
        } else {
            if ((rc = tile_attributes(which, 0, (const void*)metrics_tile_kernel<FT, false, false>, TileGeom<false>::SMEM_BYTES))) return rc;
            memset(&tm, 0, sizeof(tm));
            OSSG_CUDA_TRY(launch_pdl_smem(metrics_tile_kernel<FT, false, false>, grid, block, TileGeom<false>::SMEM_BYTES, chain, s, a, fr, tm, ride));
        }
        OSSG_LAUNCH_CHECK("metrics_tile_kernel");
    }
    if (between) OSSG_CUDA_TRY(cudaEventRecord(between, s));
    if (mode == 0) {
        OSSG_CUDA_TRY(cudaStreamWaitEvent(s, fj->join, 0));
    } else if (mode == 1 && nframe > 0) {
        // behind the tile kernel (which lets it go early) the wait comes last; behind the coordinate kernel itself first
        OSSG_CUDA_TRY(launch_pdl_smem(metrics_frame_kernel<FT>, dim3((unsigned)nfb128), dim3(128), 0, chain, s, a, f, have_tile && !want_pairs ? 0 : 1));
        OSSG_LAUNCH_CHECK("metrics_frame_kernel");
    }
    return 0;
}

static bool is_whole_grid(const ossg_tripolar_params& q)
{
    return q.i0 == 1 && q.j0 == 1 && q.nx == q.Nx && q.ny == q.Ny;
}

static int check_ptrs(void* const* v, int n, size_t align, const char* what)
{
    if (!v) return set_error(OSSG_EINVAL, "%s is NULL", what);
    for (int k = 0; k < n; ++k) {
        if (!v[k]) return set_error(OSSG_EINVAL, "%s[%d] is NULL", what, k);
        if (reinterpret_cast<uintptr_t>(v[k]) % align) return set_error(OSSG_EALIGN, "%s[%d] is not %zu-byte aligned", what, k, align);
    }
    return 0;
}

}  // namespace ossg

using namespace ossg;

namespace ossg {
template <typename FT>
__global__ void __launch_bounds__(256) cartesian_nodes_kernel(const double* __restrict__ lam, const double* __restrict__ phi, size_t n,
                                                              double radius, FT* __restrict__ x, FT* __restrict__ y, FT* __restrict__ z)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    double sin_lat, cos_lat, sin_lon, cos_lon;
    m::sincosd(__ldg(phi + t), sin_lat, cos_lat);
    m::sincosd(__ldg(lam + t), sin_lon, cos_lon);
    x[t] = (FT)(radius * cos_lon * cos_lat);
    y[t] = (FT)(radius * sin_lon * cos_lat);
    z[t] = (FT)(radius * sin_lat);
}
}  // namespace ossg

extern "C" int ossg_cartesian_nodes(const void* lam, const void* phi, size_t n, double radius, int out_type,
                                    void* x, void* y, void* z, void* stream)
{
    if (!lam || !phi || !x || !y || !z) return set_error(OSSG_EINVAL, "NULL array");
    if (out_type != OSSG_F64 && out_type != OSSG_F32) return set_error(OSSG_EINVAL, "out_type must be OSSG_F64 or OSSG_F32");
    if (n == 0) return 0;
    if (n > (size_t)0x7fffffff * 256) return set_error(OSSG_EINVAL, "too many nodes for one launch");
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (out_type == OSSG_F64)
        cartesian_nodes_kernel<double><<<blocks, 256, 0, as_stream(stream)>>>(static_cast<const double*>(lam), static_cast<const double*>(phi), n, radius,
                                                                              static_cast<double*>(x), static_cast<double*>(y), static_cast<double*>(z));
    else
        cartesian_nodes_kernel<float><<<blocks, 256, 0, as_stream(stream)>>>(static_cast<const double*>(lam), static_cast<const double*>(phi), n, radius,
                                                                             static_cast<float*>(x), static_cast<float*>(y), static_cast<float*>(z));
    OSSG_LAUNCH_CHECK("cartesian_nodes_kernel");
    return 0;
}

extern "C" size_t ossg_tripolar_workspace_bytes(const ossg_tripolar_params* p, int out_type)
{
    ossg_tripolar_params q;
    if (check_params(p, q)) return 0;
    return row_counter_offset(q, out_type) + row_counter_bytes(q);
}

extern "C" int ossg_tripolar_workspace_init(void* workspace, size_t workspace_bytes, void* stream)
{
    if (!workspace || reinterpret_cast<uintptr_t>(workspace) % 16) return set_error(OSSG_EALIGN, "workspace must be a 16-byte aligned device pointer");
    if (workspace_bytes < TABLE_HDR_WORDS * sizeof(unsigned long long)) return set_error(OSSG_EWORKSPACE, "workspace has %zu bytes, the header alone needs %zu", workspace_bytes, TABLE_HDR_WORDS * sizeof(unsigned long long));
    OSSG_CUDA_TRY(cudaMemsetAsync(workspace, 0, TABLE_HDR_WORDS * sizeof(unsigned long long), as_stream(stream)));
    return 0;
}

extern "C" size_t ossg_tripolar_table_bytes(const ossg_tripolar_params* p)
{
    ossg_tripolar_params q;
    if (check_params(p, q)) return 0;
    return table_bytes(q);
}

static int build_impl(const ossg_tripolar_params* p, int out_type, void* const out[20],
                      void* workspace, size_t workspace_bytes, void* stream, cudaEvent_t* ev /*[5] or NULL*/)
{
    ossg_tripolar_params q;
    int rc = check_params(p, q);
    if (rc) return rc;
    if (out_type != OSSG_F64 && out_type != OSSG_F32) return set_error(OSSG_EINVAL, "out_type must be OSSG_F64 or OSSG_F32");
    const size_t esz = out_type == OSSG_F64 ? 8 : 4;
    if ((rc = check_ptrs(out, 20, esz, "out"))) return rc;
    if (!workspace || reinterpret_cast<uintptr_t>(workspace) % 16) return set_error(OSSG_EALIGN, "workspace must be a 16-byte aligned device pointer");
    if (workspace_bytes < ossg_tripolar_workspace_bytes(p, out_type))
        return set_error(OSSG_EWORKSPACE, "workspace has %zu bytes, need %zu", workspace_bytes, ossg_tripolar_workspace_bytes(p, out_type));
    BuildArgs a;
    fill_common(a, q, workspace);
    for (int k = 0; k < 8; ++k) a.coords_out[k] = out[k];
    for (int k = 0; k < 12; ++k) a.metrics_out[k] = out[8 + k];
    if (out_type == OSSG_F64) {
        for (int k = 0; k < 8; ++k) a.coords64[k] = static_cast<double*>(out[k]);
    } else {
        double* scratch = reinterpret_cast<double*>(static_cast<char*>(workspace) + table_bytes(q));
        for (int k = 0; k < 8; ++k) a.coords64[k] = scratch + (size_t)k * window_elems(q);
    }
    cudaStream_t s = as_stream(stream);
    const bool whole = is_whole_grid(q);
    if (!ev && row_deps_enabled() && frame_mode() == 2 && tile_form() == 0)
        a.rows_done = reinterpret_cast<unsigned int*>(static_cast<char*>(workspace) + row_counter_offset(q, out_type));
    if (ev) OSSG_CUDA_TRY(cudaEventRecord(ev[0], s));
    if ((rc = launch_tables(a, s))) return rc;
    if (ev) OSSG_CUDA_TRY(cudaEventRecord(ev[1], s));
    rc = out_type == OSSG_F64 ? launch_coords<double>(a, s) : launch_coords<float>(a, s);
    if (rc) return rc;
    if (ev) OSSG_CUDA_TRY(cudaEventRecord(ev[2], s));
    rc = out_type == OSSG_F64 ? launch_metrics<double>(a, whole, s, ev ? ev[3] : nullptr)
                              : launch_metrics<float>(a, whole, s, ev ? ev[3] : nullptr);
    if (rc) return rc;
    if (ev) OSSG_CUDA_TRY(cudaEventRecord(ev[4], s));
    return 0;
}

extern "C" int ossg_tripolar_build(const ossg_tripolar_params* p, int out_type, void* const out[20],
                                   void* workspace, size_t workspace_bytes, void* stream)
{
    return build_impl(p, out_type, out, workspace, workspace_bytes, stream, nullptr);
}

extern "C" int ossg_tripolar_build_profiled(const ossg_tripolar_params* p, int out_type, void* const out[20],
                                            void* workspace, size_t workspace_bytes, void* stream, float ms[4])
{
    if (!ms) return set_error(OSSG_EINVAL, "ms is NULL");
    cudaEvent_t ev[5];
    for (int k = 0; k < 5; ++k) OSSG_CUDA_TRY(cudaEventCreate(&ev[k]));
    int rc = build_impl(p, out_type, out, workspace, workspace_bytes, stream, ev);
    if (rc == 0) {
        cudaError_t e = cudaEventSynchronize(ev[4]);
        if (e != cudaSuccess) rc = cuda_fail(e, "cudaEventSynchronize");
        for (int k = 0; k < 4 && rc == 0; ++k) {
            e = cudaEventElapsedTime(&ms[k], ev[k], ev[k + 1]);
            if (e != cudaSuccess) rc = cuda_fail(e, "cudaEventElapsedTime");
        }
    }
    for (int k = 0; k < 5; ++k) cudaEventDestroy(ev[k]);
    return rc;
}

extern "C" int ossg_tripolar_coords(const ossg_tripolar_params* p, void* const coords[8],
                                    void* workspace, size_t workspace_bytes, void* stream)
{
    ossg_tripolar_params q;
    int rc = check_params(p, q);
    if (rc) return rc;
    if ((rc = check_ptrs(coords, 8, 8, "coords"))) return rc;
    if (!workspace || reinterpret_cast<uintptr_t>(workspace) % 16) return set_error(OSSG_EALIGN, "workspace must be a 16-byte aligned device pointer");
    if (workspace_bytes < table_bytes(q)) return set_error(OSSG_EWORKSPACE, "workspace has %zu bytes, need %zu", workspace_bytes, table_bytes(q));
    BuildArgs a;
    fill_common(a, q, workspace);
    for (int k = 0; k < 8; ++k) { a.coords64[k] = static_cast<double*>(coords[k]); a.coords_out[k] = coords[k]; }
    for (int k = 0; k < 12; ++k) a.metrics_out[k] = nullptr;
    cudaStream_t s = as_stream(stream);
    if ((rc = launch_tables(a, s))) return rc;
    return launch_coords<double>(a, s);
}

extern "C" int ossg_tripolar_metrics(const ossg_tripolar_params* p, const void* const coords[8],
                                     int metrics_type, void* const metrics[12],
                                     void* workspace, size_t workspace_bytes, void* stream)
{
    ossg_tripolar_params q;
    int rc = check_params(p, q);
    if (rc) return rc;
    if (metrics_type != OSSG_F64 && metrics_type != OSSG_F32) return set_error(OSSG_EINVAL, "metrics_type must be OSSG_F64 or OSSG_F32");
    if ((rc = check_ptrs(const_cast<void* const*>(coords), 8, 8, "coords"))) return rc;
    if ((rc = check_ptrs(metrics, 12, metrics_type == OSSG_F64 ? 8 : 4, "metrics"))) return rc;
    if (!workspace || reinterpret_cast<uintptr_t>(workspace) % 16) return set_error(OSSG_EALIGN, "workspace must be a 16-byte aligned device pointer");
    if (workspace_bytes < table_bytes(q)) return set_error(OSSG_EWORKSPACE, "workspace has %zu bytes, need %zu", workspace_bytes, table_bytes(q));
    BuildArgs a;
    fill_common(a, q, workspace);
    for (int k = 0; k < 8; ++k) { a.coords64[k] = static_cast<double*>(const_cast<void*>(coords[k])); a.coords_out[k] = nullptr; }
    for (int k = 0; k < 12; ++k) a.metrics_out[k] = metrics[k];
    cudaStream_t s = as_stream(stream);
    if ((rc = launch_tables(a, s))) return rc;     // lat-lon rows (and closed-form tables for tile windows)
    if (metrics_type == OSSG_F64) return launch_metrics<double>(a, is_whole_grid(q), s);
    return launch_metrics<float>(a, is_whole_grid(q), s);
}
