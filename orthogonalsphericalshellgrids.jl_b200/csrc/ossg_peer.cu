// csrc/ossg_peer.cu -- the fold across an x-partitioned top row as ONE kernel over NVLink peer memory.
//
// Reference: src/zipper_boundary_condition.jl:70-155 is the operation; the reference never splits x
// (src/distributed_tripolar_grid.jl:28-31), so the fold never leaves a rank there.  With an (x, y)
// partition the fold source of global column i is column Nx-i+1 (x-Center) or Nx-i+2 (x-Face), which
// lives on the mirror rank of the top row (and, for x-Face, one column of its neighbour).  The
// pack -> ncclSend/ncclRecv -> unpack exchange (ossg_fold_pack/unpack) moves every element three
// times and pays two kernel launches plus the NCCL p2p latency.  On an NVSwitch node every GPU can
// load from every peer at NVLink speed, so here each rank's zipper kernel simply READS its source
// rows from the owner's field through a peer-mapped pointer (CUDA IPC) and writes its own halo: one
// launch, every element crosses NVLink once and is written once, no staging buffer.
//
// Ordering between ranks (no host synchronisation, no NCCL call on this path).  Every rank owns a flag
// block in device memory that its partners map through CUDA IPC: ready[r] and done[r] (epochs), r = partner rank.
//   start: block 0 stores epoch -> ready[me] on every partner (st.release.sys; no fence in front of it: the
//          rank's earlier kernels are complete by stream order).  Every warp prefetches the local rows of its
//          block's first work item, then waits until ready[r] >= epoch for all partners r (ld.acquire.sys on
//          LOCAL memory).  "Kernel started" == "interior ready, halo free to be overwritten".
//   end:   the last block to finish (local arrival counter) fences at system scope, stores epoch -> done[me] on
//          every partner and waits for done[r] >= epoch from all of them before it exits: work queued behind
//          the kernel then sees a complete halo (push) and cannot overwrite rows a partner still reads (pull).
// Epochs only grow, nothing is ever reset.  A partner that never shows up (a rank that did not make the
// matching call) is a programming error: after OSSG_PEER_TIMEOUT_MS (default 120 s, 0 = wait forever) the
// waiting kernel raises the error word of its flag block and TRAPS -- the next CUDA call of this process
// fails loudly instead of continuing on stale halos.
#include <mutex>
#include <vector>
#include <cstring>
#include <cstdlib>

#include "ossg_common.cuh"

namespace ossg {

constexpr int PB = 256;
#ifndef PEER_BLOCKS_PER_SM
#define PEER_BLOCKS_PER_SM 4
#endif

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Wait until *flag >= target.  A missing partner must not hang the GPU for ever and must not pass
// silently either: on timeout the error word is raised (for post-mortem inspection) and the kernel traps.
__device__ __forceinline__ void wait_flag(const unsigned long long* flag, unsigned long long target, unsigned int* error_word,
                                          unsigned long long timeout_ns)
{
    if (ld_acquire_sys(flag) >= target) return;
    const unsigned long long t0 = global_timer_ns();
    unsigned spins = 0;
    while (ld_acquire_sys(flag) < target) {
        if (timeout_ns && (++spins & 1023u) == 0 && global_timer_ns() - t0 > timeout_ns) {
            atomicExch(error_word, 1u);
            __threadfence_system();
            __trap();
        }
    }
}

struct PeerSync {
    unsigned long long* mine;            // this rank's flag block: ready[MAX_PEERS], done[MAX_PEERS], {unused, error word}
    void* const* peer;                   // the partners' flag blocks
    int me, nranks;
    unsigned partners;
    unsigned long long epoch, timeout_ns;
    __device__ __forceinline__ unsigned int* error_word() const { return reinterpret_cast<unsigned int*>(mine + 2 * OSSG_MAX_PEERS) + 1; }
    __device__ __forceinline__ bool is_partner(int r) const { return r < nranks && r != me && ((partners >> r) & 1u); }

    // block 0 announces this rank (call before the prefetch of the first item)
    __device__ __forceinline__ void announce() const
    {
        const int t = threadIdx.x;
        if (blockIdx.x == 0 && is_partner(t))
            st_release_sys(static_cast<unsigned long long*>(peer[t]) + me, epoch);
    }
    // every warp waits for all partners' announcements (lane r polls partner r)
    __device__ __forceinline__ void wait_ready() const
    {
        const int lane = threadIdx.x & 31;
        if (is_partner(lane)) wait_flag(mine + lane, epoch, error_word(), timeout_ns);
        __syncwarp();
    }
    // end of the kernel: the LAST block to finish (a local arrival counter) fences at system scope, tells every
    // partner "rank me has finished exchange `epoch`" and waits for the same from all of them before it exits.
    // (Measured alternative, dropped: every block signalling its partners itself with red.release.sys.add -- 592
    // remote atomics on one address serialise at the partner's L2 and cost 7-13 us per exchange on 2 GPUs.)
    __device__ __forceinline__ void finish() const
    {
        __shared__ int s_last;
        __syncthreads();                                              // the block's stores / peer loads are issued
        if (threadIdx.x == 0) {
            unsigned int* counter = reinterpret_cast<unsigned int*>(mine + 2 * OSSG_MAX_PEERS);
            __threadfence();                                          // release this block's stores ...
            s_last = atomicAdd(counter, 1u) == gridDim.x - 1;
            if (s_last) {
                __threadfence();                                      // ... and acquire everybody else's
                atomicExch(counter, 0u);                              // ready for the next launch (stream-ordered)
            }
        }
        __syncthreads();
        const int t = threadIdx.x;
        if (s_last && t < 32) {
            __threadfence_system();
            if (is_partner(t)) {
                st_release_sys(static_cast<unsigned long long*>(peer[t]) + OSSG_MAX_PEERS + me, epoch);
                wait_flag(mine + OSSG_MAX_PEERS + t, epoch, error_word(), timeout_ns);
            }
        }
    }
};

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

// loads of peer memory: system-scope relaxed, so that no non-coherent cache serves an older epoch
__device__ __forceinline__ double peer_load(const double* p)
{
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float peer_load(const float* p)
{
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

struct PeerBatch {
    ossg_peer_field pf[OSSG_MAX_PEER_BATCH];
    ossg_peer_fold fold;
    int n;
    unsigned int partners;       // bit r: rank r owns fold sources / destinations of this rank's columns
    int chunks, levels;          // work items: n fields x levels x chunks of PB columns (bulk push: of bulk_cw columns)
    int bulk_cw;
    long long nitems;
    unsigned long long timeout_ns;
};

// one thread per (local column i, level): all Hy source rows are loaded (over NVLink) before any
// store, as in the single-GPU zipper_column.
template <typename T>
__device__ __forceinline__ void peer_column(const ossg_peer_field& p, const ossg_peer_fold& fd, int i /*1..nx local*/, int kk)
{
    const ossg_field_desc& f = p.f;
    const int gNx = fd.gNx, Ny = f.Ny, Hy = f.Hy, Hx = f.Hx;
    const int ig = fd.i0[fd.me] + i - 1;                              // global destination column
    double s = f.sign;
    int ipg = f.locx ? gNx - ig + 2 : gNx - ig + 1;                   // zipper_boundary_condition.jl:73,90,110,125
    if (ipg > gNx) { s = fabs(s); ipg -= gNx; }                       // :74-75, :91-92
    int r = 0;
#pragma unroll 1
    for (; r < fd.nranks - 1; ++r)
        if (ipg < fd.i0[r] + fd.nx[r]) break;                         // owner of the source column
    const int li = ipg - fd.i0[r] + 1;                                // its local column there
    const size_t sxr = (size_t)(fd.nx[r] + 2 * Hx), sxyr = sxr * (size_t)(Ny + 2 * Hy);
    const T* srcb = static_cast<const T*>(p.peer[r]) + (size_t)(f.k0 + kk) * sxyr + (li + Hx - 1);
    const size_t sx = (size_t)(f.Nx + 2 * Hx), sxy = sx * (size_t)(Ny + 2 * Hy);
    T* c = static_cast<T*>(f.data) + (size_t)(f.k0 + kk) * sxy;
    // halo row j = q + 1: dst row Ny + j, src row Ny - j (y-Center) or Ny - j + 1 (y-Face)
    const T* src = srcb + (size_t)((f.locy ? Ny : Ny - 1) + Hy - 1) * sxr;
    T* dst = c + (size_t)(Ny + Hy) * sx + (i + Hx - 1);
    const bool fix_row = (f.locy == OSSG_CENTER) && (ig > gNx / 2);   // :102, :135 (global test)
    T vrow = T(0);
    if (fix_row) vrow = peer_load(srcb + (size_t)(Ny + Hy - 1) * sxr);
    int q = 0;
    for (; q + 4 <= Hy; q += 4) {
        const T v0 = peer_load(src - (ptrdiff_t)((q + 0) * sxr)), v1 = peer_load(src - (ptrdiff_t)((q + 1) * sxr));
        const T v2 = peer_load(src - (ptrdiff_t)((q + 2) * sxr)), v3 = peer_load(src - (ptrdiff_t)((q + 3) * sxr));
        dst[(size_t)(q + 0) * sx] = apply_sign(v0, s);
        dst[(size_t)(q + 1) * sx] = apply_sign(v1, s);
        dst[(size_t)(q + 2) * sx] = apply_sign(v2, s);
        dst[(size_t)(q + 3) * sx] = apply_sign(v3, s);
    }
    for (; q < Hy; ++q) dst[(size_t)q * sx] = apply_sign(peer_load(src - (ptrdiff_t)(q * sxr)), s);
    if (fix_row) c[(size_t)(Ny + Hy - 1) * sx + (i + Hx - 1)] = apply_sign(vrow, s);
}

// push form: the thread owns SOURCE column i of this rank and stores the folded values into the
// halo of the rank that owns the destination column (posted NVLink writes instead of round-trip
// reads).  The fold map is an involution on 1..Nx (i <-> Nx-i+1, resp. Nx-i+2 with column 1 fixed),
// so every destination receives exactly one value.
template <typename T>
__device__ __forceinline__ void peer_column_push(const ossg_peer_field& p, const ossg_peer_fold& fd, int i /*1..nx local*/, int kk)
{
    const ossg_field_desc& f = p.f;
    const int gNx = fd.gNx, Ny = f.Ny, Hy = f.Hy, Hx = f.Hx;
    const int isg = fd.i0[fd.me] + i - 1;                             // global source column
    double s = f.sign;
    int idg = f.locx ? gNx - isg + 2 : gNx - isg + 1;                 // destination column
    if (idg > gNx) { s = fabs(s); idg -= gNx; }                       // x-Face column 1 folds onto itself with |s|
    int r = 0;
#pragma unroll 1
    for (; r < fd.nranks - 1; ++r)
        if (idg < fd.i0[r] + fd.nx[r]) break;                         // owner of the destination column
    const int li = idg - fd.i0[r] + 1;
    const size_t sxr = (size_t)(fd.nx[r] + 2 * Hx), sxyr = sxr * (size_t)(Ny + 2 * Hy);
    T* dstb = static_cast<T*>(const_cast<void*>(p.peer[r])) + (size_t)(f.k0 + kk) * sxyr + (li + Hx - 1);
    const size_t sx = (size_t)(f.Nx + 2 * Hx), sxy = sx * (size_t)(Ny + 2 * Hy);
    const T* c = static_cast<const T*>(f.data) + (size_t)(f.k0 + kk) * sxy + (i + Hx - 1);
    const T* src = c + (size_t)((f.locy ? Ny : Ny - 1) + Hy - 1) * sx;    // q = 0, then -sx per row
    T* dst = dstb + (size_t)(Ny + Hy) * sxr;                                // q = 0, then +sxr per row
    const bool fix_row = (f.locy == OSSG_CENTER) && (idg > gNx / 2);
    T vrow = T(0);
    if (fix_row) vrow = c[(size_t)(Ny + Hy - 1) * sx];
    int q = 0;
    for (; q + 4 <= Hy; q += 4) {
        const T v0 = src[-(ptrdiff_t)((q + 0) * sx)], v1 = src[-(ptrdiff_t)((q + 1) * sx)];
        const T v2 = src[-(ptrdiff_t)((q + 2) * sx)], v3 = src[-(ptrdiff_t)((q + 3) * sx)];
        dst[(size_t)(q + 0) * sxr] = apply_sign(v0, s);
        dst[(size_t)(q + 1) * sxr] = apply_sign(v1, s);
        dst[(size_t)(q + 2) * sxr] = apply_sign(v2, s);
        dst[(size_t)(q + 3) * sxr] = apply_sign(v3, s);
    }
    for (; q < Hy; ++q) dst[(size_t)q * sxr] = apply_sign(src[-(ptrdiff_t)(q * sx)], s);
    if (fix_row) dstb[(size_t)(Ny + Hy - 1) * sxr] = apply_sign(vrow, s);
}

// ---- push form with bulk stores (fold.push == 2).  The block stages the folded rows of one
// (field, level, 256 source columns) item in shared memory in DESTINATION order (the reversal and the
// sign happen on the way in), then a few elected threads hand every destination row segment to the
// bulk-copy engine (cp.async.bulk shared -> global, the 1-D TMA store): the segment crosses NVLink as
// one 2 KB transfer per row instead of 64 SM-issued 32-byte-per-lane stores.  A bulk copy needs 16-byte
// aligned addresses and sizes, so the staging position of an element has the parity of its destination
// address, and the (at most A-1) leading / trailing elements of a segment that fall outside the aligned
// run -- as well as the wrapped x-Face column 1 -- are stored directly by their threads.  A chunk's
// destinations may touch up to BULK_NREG ranks (the x-Face map is shifted by one column against the
// partition); elements beyond that are stored directly too.
constexpr int BULK_ROWS = 5;            // rows staged per pass (Hy = 4 plus the row-Ny overwrite: one pass)
constexpr int BULK_NREG = 2;            // destination ranks per chunk served by bulk copies
constexpr int BULK_CW = 544;            // widest chunk, columns: a 540-column tile (1/12 degree on 8 GPUs) is ONE chunk,
                                        // i.e. one 4.3 KB transfer per row; wider tiles are cut into equal chunks
constexpr int BULK_RS = BULK_CW + 4;    // staging row stride in elements (alignment slack: up to 3 elements)

__device__ __forceinline__ void bulk_store_s2g(void* gdst, const void* ssrc, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gdst), "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}

template <typename T>
__device__ __forceinline__ void peer_chunk_push_bulk(const ossg_peer_field& p, const ossg_peer_fold& fd, int chunk, int cw, int kk, unsigned char* sraw)
{
    constexpr long long A = 16 / sizeof(T);                           // elements per 16 bytes
    T* sb = reinterpret_cast<T*>(sraw);
    const ossg_field_desc& f = p.f;
    const int tx = threadIdx.x;
    const int gNx = fd.gNx, Ny = f.Ny, Hy = f.Hy, Hx = f.Hx;
    const int c0 = chunk * cw;
    const int n = min(cw, f.Nx - c0);                                 // source columns c0+1 .. c0+n of this rank
    const int sg_lo = fd.i0[fd.me] + c0, sg_hi = sg_lo + n - 1;       // their global indices
    const int cst = f.locx ? gNx + 2 : gNx + 1;                       // destination column = cst - source column
    const int dlo = cst - sg_hi;                                      // destination range, ascending, without the
    const int dhi = min(cst - sg_lo, gNx);                            // wrapped x-Face column (source 1 -> Nx+1 -> 1)
    int r_lo = 0;
#pragma unroll 1
    for (; r_lo < fd.nranks - 1; ++r_lo)
        if (dlo < fd.i0[r_lo] + fd.nx[r_lo]) break;                   // owner of the lowest destination column
    const int half = gNx / 2;
    const bool center = f.locy == OSSG_CENTER;
    const int nrows = Hy + (center ? 1 : 0);                          // row q < Hy: halo row q; q == Hy: row Ny (:102, :135)
    const size_t sx = (size_t)(f.Nx + 2 * Hx), sxy = sx * (size_t)(Ny + 2 * Hy);

#pragma unroll 1
    for (int q0 = 0; q0 < nrows; q0 += BULK_ROWS) {
        const int nq = min(BULK_ROWS, nrows - q0);
        if (tx < BULK_NREG * BULK_ROWS) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging buffer free again
        __syncthreads();
#pragma unroll 1
        for (int col = tx; col < n; col += PB) {                      // this thread's source columns and their destinations
            const int isg = sg_lo + col;
            double s = f.sign;
            int idg = cst - isg;
            bool wrapped = false;
            if (idg > gNx) { s = fabs(s); idg -= gNx; wrapped = true; }
            int r = 0;
#pragma unroll 1
            for (; r < fd.nranks - 1; ++r)
                if (idg < fd.i0[r] + fd.nx[r]) break;
            const size_t sxr = (size_t)(fd.nx[r] + 2 * Hx), sxyr = sxr * (size_t)(Ny + 2 * Hy);
            T* dstb = static_cast<T*>(const_cast<void*>(p.peer[r])) + (size_t)(f.k0 + kk) * sxyr + (idg - fd.i0[r] + Hx);
            const T* c = static_cast<const T*>(f.data) + (size_t)(f.k0 + kk) * sxy + (c0 + col + Hx);
            const int reg = r - r_lo;
            const int chi = min(dhi, fd.i0[r] + fd.nx[r] - 1);        // this column's segment: [clo, chi] on rank r
            const int clo_halo = max(dlo, fd.i0[r]);
#pragma unroll 1
            for (int qq = 0; qq < nq; ++qq) {
                const int q = q0 + qq;
                const bool fix = q == Hy;
                if (fix && idg <= half) continue;                     // row Ny is overwritten for the right half only
                const T v = apply_sign(fix ? c[(size_t)(Ny + Hy - 1) * sx] : c[(size_t)((f.locy ? Ny : Ny - 1) + Hy - 1 - q) * sx], s);
                T* d = dstb + (fix ? (size_t)(Ny + Hy - 1) : (size_t)(Ny + Hy + q)) * sxr;
                const int clo = fix ? max(clo_halo, half + 1) : clo_halo;
                bool direct = wrapped || reg >= BULK_NREG;
                if (!direct) {
                    const long long E = (long long)(reinterpret_cast<uintptr_t>(d) / sizeof(T));   // destination element index
                    const long long E_lo = E - (idg - clo), E_end = E + (chi - idg) + 1;
                    const long long Ea = (E_lo + A - 1) & ~(A - 1), Eb = E_end & ~(A - 1);       // 16-byte aligned run
                    if (E >= Ea && E < Eb) sb[(size_t)(reg * BULK_ROWS + qq) * BULK_RS + (size_t)(E - (E_lo & ~(A - 1)))] = v;
                    else direct = true;
                }
                if (direct) *d = v;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // staged values visible to the bulk-copy engine
        __syncthreads();
        if (tx < BULK_NREG * nq) {
            const int g = tx / nq, qq = tx - g * nq, q = q0 + qq, rr = r_lo + g;
            if (rr < fd.nranks) {
                const bool fix = q == Hy;
                int clo = max(dlo, fd.i0[rr]);
                if (fix) clo = max(clo, half + 1);
                const int ch = min(dhi, fd.i0[rr] + fd.nx[rr] - 1);
                if (ch >= clo) {
                    const size_t sxq = (size_t)(fd.nx[rr] + 2 * Hx), sxyq = sxq * (size_t)(Ny + 2 * Hy);
                    T* d = static_cast<T*>(const_cast<void*>(p.peer[rr])) + (size_t)(f.k0 + kk) * sxyq + (clo - fd.i0[rr] + Hx)
                           + (fix ? (size_t)(Ny + Hy - 1) : (size_t)(Ny + Hy + q)) * sxq;
                    const long long E_lo = (long long)(reinterpret_cast<uintptr_t>(d) / sizeof(T)), E_end = E_lo + (ch - clo) + 1;
                    const long long Ea = (E_lo + A - 1) & ~(A - 1), Eb = E_end & ~(A - 1);
                    if (Eb > Ea)
                        bulk_store_s2g(reinterpret_cast<void*>(static_cast<uintptr_t>(Ea) * sizeof(T)),
                                       &sb[(size_t)(g * BULK_ROWS + qq) * BULK_RS + (size_t)(Ea - (E_lo & ~(A - 1)))],
                                       (unsigned)((Eb - Ea) * sizeof(T)));
                }
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
}

// local source rows of the block's first work item, requested from DRAM while the announcements travel
__device__ __forceinline__ void prefetch_first_item(const PeerBatch& b)
{
    if ((long long)blockIdx.x >= b.nitems || !b.fold.push || (threadIdx.x & 7)) return;
    const int chunk = (int)(blockIdx.x % b.chunks);
    const long long rest = blockIdx.x / b.chunks;
    const int kk = (int)(rest % b.levels), fi = (int)(rest / b.levels);
    const ossg_field_desc& f = b.pf[fi].f;
    if (kk >= f.nk) return;
    const size_t esz = f.elem_type == OSSG_F64 ? 8 : 4;
    const size_t sx = (size_t)(f.Nx + 2 * f.Hx), sxy = sx * (size_t)(f.Ny + 2 * f.Hy);
    const int cw = b.fold.push == 2 ? b.bulk_cw : PB;
    const char* base = static_cast<const char*>(f.data) + ((size_t)(f.k0 + kk) * sxy + (size_t)f.Hx) * esz;
    const int rows = f.Hy + (f.locy == OSSG_CENTER ? 1 : 0);
    for (int col = chunk * cw + threadIdx.x; col < min((chunk + 1) * cw, f.Nx); col += PB)
        for (int q = 0; q < rows; ++q) {
            const int row = q == f.Hy ? f.Ny + f.Hy - 1 : (f.locy ? f.Ny : f.Ny - 1) + f.Hy - 1 - q;
            prefetch_l2(base + ((size_t)row * sx + col) * esz);
        }
}

__global__ void __launch_bounds__(PB, PEER_BLOCKS_PER_SM) zipper_peer_kernel(const __grid_constant__ PeerBatch b)
{
    const ossg_peer_fold& fd = b.fold;
    __shared__ __align__(128) unsigned char s_bulk[BULK_NREG * BULK_ROWS * BULK_RS * sizeof(double)];   // staging rows of the bulk push form
    const PeerSync sync{ static_cast<unsigned long long*>(fd.flags_local), fd.flags_peer, fd.me, fd.nranks, b.partners, fd.epoch, b.timeout_ns };
    sync.announce();
    prefetch_first_item(b);
    sync.wait_ready();

    // persistent blocks: a block-stride loop over the work items (field, level, chunk of PB columns)
    // -- few waves, many NVLink transfers in flight per SM
    for (long long item = blockIdx.x; item < b.nitems; item += gridDim.x) {
        const int chunk = (int)(item % b.chunks);
        const long long rest = item / b.chunks;
        const int kk = (int)(rest % b.levels), fi = (int)(rest / b.levels);
        const ossg_peer_field& p = b.pf[fi];
        const int i = chunk * PB + threadIdx.x + 1;
        if (fd.push == 2) {                                           // (block-uniform)
            if (chunk * b.bulk_cw < p.f.Nx && kk < p.f.nk) {
                if (p.f.elem_type == OSSG_F64) peer_chunk_push_bulk<double>(p, fd, chunk, b.bulk_cw, kk, s_bulk);
                else peer_chunk_push_bulk<float>(p, fd, chunk, b.bulk_cw, kk, s_bulk);
            }
            continue;
        }
        if (i <= p.f.Nx && kk < p.f.nk) {
            if (fd.push) {
                if (p.f.elem_type == OSSG_F64) peer_column_push<double>(p, fd, i, kk);
                else peer_column_push<float>(p, fd, i, kk);
            } else {
                if (p.f.elem_type == OSSG_F64) peer_column<double>(p, fd, i, kk);
                else peer_column<float>(p, fd, i, kk);
            }
        }
    }

    if (fd.push == 2) {                                               // bulk stores of this thread complete (async proxy) and
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // ordered before the generic-proxy signal in finish()
        asm volatile("fence.proxy.async;" ::: "memory");
    }
    sync.finish();
}

// ------------------------------------------------------------------ ordinary halos (push)
struct HaloBatch {
    ossg_peer_halo_field hf[OSSG_MAX_PEER_BATCH];
    ossg_peer_halo halo;
    int n;
    unsigned int partners;
    long long items_per_field;           // phase 0: 2 directions x Hy rows x parent levels x column chunks
    int chunks;                          // phase 0: ceil((Nx + 2Hx) / PB); phase 1: ceil(rows / PB)
    unsigned long long timeout_ns;
};

template <typename T>
__device__ __forceinline__ void halo_rows_push(const ossg_peer_halo_field& h, long long item, int chunks)
{
    // item -> (direction, halo row r, parent level, column chunk)
    const ossg_field_desc& f = h.f;
    const int sx = f.Nx + 2 * f.Hx, levels = f.Nz + 2 * f.Hz;
    const int chunk = (int)(item % chunks);
    long long rest = item / chunks;
    const int lev = (int)(rest % levels);
    rest /= levels;
    const int r = (int)(rest % f.Hy), dir = (int)(rest / f.Hy);     // dir 0: to the north neighbour, 1: to the south neighbour
    const int col = chunk * PB + threadIdx.x;
    if (col >= sx) return;
    T* peer = static_cast<T*>(dir ? h.south : h.north);
    if (!peer) return;
    const int ny_peer = dir ? h.ny_south : h.ny_north;
    const size_t sxy = (size_t)sx * (size_t)(f.Ny + 2 * f.Hy), sxy_peer = (size_t)sx * (size_t)(ny_peer + 2 * f.Hy);
    const int src_row = dir ? f.Hy + r : f.Ny + r;                    // bottom / top Hy interior rows (parent index)
    const int dst_row = dir ? ny_peer + f.Hy + r : r;                 // the neighbour's north / south halo
    peer[(size_t)lev * sxy_peer + (size_t)dst_row * sx + col] = static_cast<const T*>(f.data)[(size_t)lev * sxy + (size_t)src_row * sx + col];
}

// staged form of the column exchange: push = local strided read, contiguous remote write;
// unpack (phase 2) = contiguous local read, strided local write
template <typename T>
__device__ __forceinline__ void halo_cols_stage(const ossg_peer_halo_field& h, const ossg_peer_halo& hd, int fi, long long item, int chunks, bool unpack)
{
    const ossg_field_desc& f = h.f;
    const int chunk = (int)(item % chunks), side = (int)(item / chunks);   // push: side 0 -> east neighbour; unpack: side 0 = west halo
    const long long rows = (long long)(f.Ny + 2 * f.Hy) * (f.Nz + 2 * f.Hz);
    const long long row = (long long)chunk * PB + threadIdx.x;
    if (row >= rows) return;
    const size_t piece = ((size_t)fi * rows + row) * f.Hx;
    if (!unpack) {
        T* stage = static_cast<T*>(side ? hd.stage_to_west : hd.stage_to_east);
        if (!(side ? h.west : h.east)) return;
        const T* src = static_cast<const T*>(f.data) + (size_t)row * (size_t)(f.Nx + 2 * f.Hx) + (side ? f.Hx : f.Nx);
        copy_halo_piece(stage + piece, src, f.Hx);
    } else {
        const T* stage = static_cast<const T*>(side ? hd.stage_from_east : hd.stage_from_west);
        if (!(side ? h.east : h.west)) return;
        T* dst = static_cast<T*>(f.data) + (size_t)row * (size_t)(f.Nx + 2 * f.Hx) + (side ? f.Nx + f.Hx : 0);
        copy_halo_piece(dst, stage + piece, f.Hx);
    }
}

template <typename T>
__device__ __forceinline__ void halo_cols_push(const ossg_peer_halo_field& h, long long item, int chunks)
{
    // item -> (side, chunk of PB parent rows); a thread moves the Hx elements of one (row, side)
    const ossg_field_desc& f = h.f;
    const int chunk = (int)(item % chunks), side = (int)(item / chunks);   // side 0: to the east neighbour, 1: to the west neighbour
    const long long rows = (long long)(f.Ny + 2 * f.Hy) * (f.Nz + 2 * f.Hz);
    const long long row = (long long)chunk * PB + threadIdx.x;
    if (row >= rows) return;
    T* peer = static_cast<T*>(side ? h.west : h.east);
    if (!peer) return;
    const int nx_peer = side ? h.nx_west : h.nx_east;
    const T* src = static_cast<const T*>(f.data) + (size_t)row * (size_t)(f.Nx + 2 * f.Hx) + (side ? f.Hx : f.Nx);
    T* dst = peer + (size_t)row * (size_t)(nx_peer + 2 * f.Hx) + (side ? nx_peer + f.Hx : 0);
    copy_halo_piece(dst, src, f.Hx);
}

__global__ void __launch_bounds__(PB, PEER_BLOCKS_PER_SM) halo_peer_kernel(const __grid_constant__ HaloBatch b)
{
    const ossg_peer_halo& hd = b.halo;
    const PeerSync sync{ static_cast<unsigned long long*>(hd.flags_local), hd.flags_peer, hd.me, hd.nranks, b.partners, hd.epoch, b.timeout_ns };
    sync.announce();
    sync.wait_ready();                                                // (see zipper_peer_kernel)

    const long long nitems = b.items_per_field * b.n;
    for (long long item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int fi = (int)(item / b.items_per_field);
        const long long it = item - (long long)fi * b.items_per_field;
        const ossg_peer_halo_field& h = b.hf[fi];
        if (hd.phase == 0) {
            if (h.f.elem_type == OSSG_F64) halo_rows_push<double>(h, it, b.chunks);
            else halo_rows_push<float>(h, it, b.chunks);
        } else if (hd.phase == 2 || hd.stage_to_east) {
            if (h.f.elem_type == OSSG_F64) halo_cols_stage<double>(h, hd, fi, it, b.chunks, hd.phase == 2);
            else halo_cols_stage<float>(h, hd, fi, it, b.chunks, hd.phase == 2);
        } else {
            if (h.f.elem_type == OSSG_F64) halo_cols_push<double>(h, it, b.chunks);
            else halo_cols_push<float>(h, it, b.chunks);
        }
    }
    sync.finish();
}

// how long a kernel waits for a partner before it traps (OSSG_PEER_TIMEOUT_MS, read once; 0 = for ever)
static unsigned long long peer_timeout_ns()
{
    static const int ms = env_int("OSSG_PEER_TIMEOUT_MS", 120000);
    return ms <= 0 ? 0ull : (unsigned long long)ms * 1000000ull;
}

// ------------------------------------------------------------------ CUDA IPC plumbing
struct OpenedHandle { unsigned char h[64]; void* base; int refs; };
static std::mutex g_ipc_mu;
static std::vector<OpenedHandle> g_opened;

typedef int (*PFN_cuMemGetAddressRange)(unsigned long long*, size_t*, unsigned long long);

static int address_range(const void* ptr, void** base, size_t* size)
{
    static PFN_cuMemGetAddressRange fn = nullptr;
    if (!fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        OSSG_CUDA_TRY(cudaGetDriverEntryPoint("cuMemGetAddressRange", &sym, cudaEnableDefault, &q));
        if (!sym || q != cudaDriverEntryPointSuccess) return set_error(OSSG_EINVAL, "cuMemGetAddressRange is not available from the driver");
        fn = reinterpret_cast<PFN_cuMemGetAddressRange>(sym);
    }
    unsigned long long b = 0;
    size_t n = 0;
    const int rc = fn(&b, &n, (unsigned long long)reinterpret_cast<uintptr_t>(ptr));
    if (rc != 0) return set_error(OSSG_EINVAL, "cuMemGetAddressRange failed with CUresult %d (not a device allocation?)", rc);
    *base = reinterpret_cast<void*>(static_cast<uintptr_t>(b));
    *size = n;
    return 0;
}

}  // namespace ossg

using namespace ossg;

static_assert(sizeof(cudaIpcMemHandle_t) == OSSG_IPC_HANDLE_BYTES, "ossg_ipc_handle must hold a cudaIpcMemHandle_t");

extern "C" int ossg_ipc_export(const void* device_ptr, ossg_ipc_handle* handle, uint64_t* offset)
{
    if (!device_ptr || !handle || !offset) return set_error(OSSG_EINVAL, "NULL argument");
    void* base = nullptr;
    size_t size = 0;
    int rc = address_range(device_ptr, &base, &size);
    if (rc) return rc;
    cudaIpcMemHandle_t h;
    OSSG_CUDA_TRY(cudaIpcGetMemHandle(&h, base));
    std::memcpy(handle->bytes, &h, sizeof(h));
    *offset = (uint64_t)(static_cast<const char*>(device_ptr) - static_cast<const char*>(base));
    return 0;
}

extern "C" int ossg_ipc_open(const ossg_ipc_handle* handle, uint64_t offset, void** peer_ptr)
{
    if (!handle || !peer_ptr) return set_error(OSSG_EINVAL, "NULL argument");
    std::lock_guard<std::mutex> lock(g_ipc_mu);
    for (OpenedHandle& o : g_opened)
        if (std::memcmp(o.h, handle->bytes, 64) == 0) {
            ++o.refs;
            *peer_ptr = static_cast<char*>(o.base) + offset;
            return 0;
        }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle->bytes, sizeof(h));
    void* base = nullptr;
    OSSG_CUDA_TRY(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    OpenedHandle o;
    std::memcpy(o.h, handle->bytes, 64);
    o.base = base;
    o.refs = 1;
    g_opened.push_back(o);
    *peer_ptr = static_cast<char*>(base) + offset;
    return 0;
}

extern "C" int ossg_ipc_close(const ossg_ipc_handle* handle)
{
    if (!handle) return set_error(OSSG_EINVAL, "NULL argument");
    std::lock_guard<std::mutex> lock(g_ipc_mu);
    for (size_t k = 0; k < g_opened.size(); ++k)
        if (std::memcmp(g_opened[k].h, handle->bytes, 64) == 0) {
            if (--g_opened[k].refs > 0) return 0;
            void* base = g_opened[k].base;
            g_opened.erase(g_opened.begin() + (long)k);
            OSSG_CUDA_TRY(cudaIpcCloseMemHandle(base));
            return 0;
        }
    return set_error(OSSG_EINVAL, "handle is not open");
}

extern "C" int ossg_zipper_fill_peer(const ossg_peer_field* fields, int nfields, const ossg_peer_fold* fold, void* stream)
{
    if (!fields || !fold || nfields < 1) return set_error(OSSG_EINVAL, "fields/fold is NULL or nfields < 1");
    if (nfields > OSSG_MAX_PEER_BATCH) return set_error(OSSG_EINVAL, "at most %d fields per peer launch, got %d", OSSG_MAX_PEER_BATCH, nfields);
    const ossg_peer_fold& fd = *fold;
    if (fd.nranks < 1 || fd.nranks > OSSG_MAX_PEERS || fd.me < 0 || fd.me >= fd.nranks)
        return set_error(OSSG_EINVAL, "bad rank layout: me = %d of %d (max %d)", fd.me, fd.nranks, OSSG_MAX_PEERS);
    if (fd.gNx < 2 || fd.gNx % 2) return set_error(OSSG_EODD_NX, "global Nx must be even, got %d", fd.gNx);
    int next = 1;
    for (int r = 0; r < fd.nranks; ++r) {
        if (fd.i0[r] != next || fd.nx[r] < 1) return set_error(OSSG_EWINDOW, "top-row tiles must cover 1..Nx in rank order (rank %d: i0 = %d, nx = %d)", r, fd.i0[r], fd.nx[r]);
        next += fd.nx[r];
        if (r != fd.me && !fd.flags_peer[r]) return set_error(OSSG_EINVAL, "flags_peer[%d] is NULL", r);
    }
    if (next != fd.gNx + 1) return set_error(OSSG_EWINDOW, "top-row tiles cover %d columns, global Nx = %d", next - 1, fd.gNx);
    if (!fd.flags_local || reinterpret_cast<uintptr_t>(fd.flags_local) % 8) return set_error(OSSG_EALIGN, "flags_local must be an 8-byte aligned device pointer");
    if (fd.epoch == 0) return set_error(OSSG_EINVAL, "epoch must start at 1 and grow by one per call");
    if (fd.push < 0 || fd.push > 2) return set_error(OSSG_EINVAL, "push must be 0 (pull), 1 (push) or 2 (bulk push), got %d", fd.push);
    PeerBatch b;
    b.fold = fd;
    b.n = nfields;
    int gx = 0, gy = 0;
    for (int k = 0; k < nfields; ++k) {
        b.pf[k] = fields[k];
        ossg_field_desc& f = b.pf[k].f;
        if (!f.data) return set_error(OSSG_EINVAL, "field %d: data is NULL", k);
        if (f.elem_type != OSSG_F64 && f.elem_type != OSSG_F32) return set_error(OSSG_EINVAL, "field %d: bad elem_type", k);
        if (f.Nx != fd.nx[fd.me]) return set_error(OSSG_EWINDOW, "field %d: local Nx (%d) differs from this rank's tile width (%d)", k, f.Nx, fd.nx[fd.me]);
        if (f.Hy < 1 || f.Hy > f.Ny - 1 || f.Hx < 0) return set_error(OSSG_EHALO, "field %d: bad halo (%d, %d)", k, f.Hx, f.Hy);
        if ((f.locx != OSSG_CENTER && f.locx != OSSG_FACE) || (f.locy != OSSG_CENTER && f.locy != OSSG_FACE))
            return set_error(OSSG_ELOCATION, "field %d: location codes must be OSSG_CENTER or OSSG_FACE", k);
        if (f.nk == 0) { f.k0 = f.Hz; f.nk = f.Nz; }
        if (f.k0 < 0 || f.nk < 1 || f.k0 + f.nk > f.Nz + 2 * f.Hz) return set_error(OSSG_EINVAL, "field %d: bad k-range", k);
        for (int r = 0; r < fd.nranks; ++r)
            if (!b.pf[k].peer[r]) return set_error(OSSG_EINVAL, "field %d: peer[%d] is NULL", k, r);
        gx = max(gx, ceil_div(f.Nx, PB));
        gy = max(gy, f.nk);
    }
    int dev = 0, sms = 0;
    OSSG_CUDA_TRY(cudaGetDevice(&dev));
    OSSG_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long resident = (long long)sms * PEER_BLOCKS_PER_SM;
    b.bulk_cw = PB;
    if (fd.push == 2) {
        // equal chunks (all fields share this rank's Nx): as wide as the staging rows allow -- few, large
        // transfers -- but at least ~3 work items per resident block so that the persistent blocks finish
        // together, as long as that keeps a chunk at 256 columns or more
        const int nx = fd.nx[fd.me];
        const int nch_min = ceil_div(nx, BULK_CW);
        const int nch_cap = max(nch_min, nx / PB);
        static const int env_waves = env_int("OSSG_PEER_BULK_WAVES", 3);  // (read once; 0: always the widest chunks)
        const int waves = fd.bulk_waves > 0 ? fd.bulk_waves : fd.bulk_waves < 0 ? 0 : (env_waves < 0 ? 0 : env_waves);
        const int nch_bal = ceil_div((long long)waves * resident, (long long)nfields * gy);
        const int nch = min(max(nch_bal, nch_min), nch_cap);
        b.bulk_cw = ceil_div(nx, nch);
        gx = ceil_div(nx, b.bulk_cw);
    }
    // partner ranks: owners of the images of this rank's columns under the fold map (an involution,
    // so the relation is symmetric: r is my partner exactly when I am r's)
    b.partners = 0;
    {
        const int lo = fd.i0[fd.me], hi = lo + fd.nx[fd.me] - 1;
        for (int k = 0; k < nfields; ++k) {
            const int c = b.pf[k].f.locx ? 2 : 1;
            for (int r = 0; r < fd.nranks; ++r) {
                const int rlo = fd.i0[r], rhi = rlo + fd.nx[r] - 1;
                int a0 = fd.gNx - hi + c, a1 = fd.gNx - lo + c;               // image range, ascending
                bool hit = a0 <= rhi && (a1 > fd.gNx ? fd.gNx : a1) >= rlo;
                if (a1 > fd.gNx && rlo == 1) hit = true;                     // x-Face: column Nx+1 wraps onto column 1
                if (hit) b.partners |= 1u << r;
            }
        }
    }
    b.chunks = gx;
    b.levels = gy;
    b.nitems = (long long)gx * gy * nfields;
    b.timeout_ns = peer_timeout_ns();
    const int grid = (int)(b.nitems < resident ? b.nitems : resident);
    zipper_peer_kernel<<<grid, PB, 0, as_stream(stream)>>>(b);
    OSSG_LAUNCH_CHECK("zipper_peer_kernel");
    return 0;
}

extern "C" int ossg_halo_exchange_peer(const ossg_peer_halo_field* fields, int nfields, const ossg_peer_halo* halo, void* stream)
{
    if (!fields || !halo || nfields < 1) return set_error(OSSG_EINVAL, "fields/halo is NULL or nfields < 1");
    if (nfields > OSSG_MAX_PEER_BATCH) return set_error(OSSG_EINVAL, "at most %d fields per peer launch, got %d", OSSG_MAX_PEER_BATCH, nfields);
    const ossg_peer_halo& hd = *halo;
    if (hd.nranks < 1 || hd.nranks > OSSG_MAX_PEERS || hd.me < 0 || hd.me >= hd.nranks)
        return set_error(OSSG_EINVAL, "bad rank layout: me = %d of %d (max %d)", hd.me, hd.nranks, OSSG_MAX_PEERS);
    if (hd.phase < 0 || hd.phase > 2) return set_error(OSSG_EINVAL, "phase must be 0 (rows), 1 (columns) or 2 (unpack staged columns)");
    const bool staged = hd.stage_to_east || hd.stage_to_west || hd.stage_from_east || hd.stage_from_west;
    if (staged && !(hd.stage_to_east && hd.stage_to_west && hd.stage_from_east && hd.stage_from_west))
        return set_error(OSSG_EINVAL, "staging needs all four buffers");
    if (hd.phase == 2 && !staged) return set_error(OSSG_EINVAL, "phase 2 needs the staging buffers");
    if (!hd.flags_local || reinterpret_cast<uintptr_t>(hd.flags_local) % 8) return set_error(OSSG_EALIGN, "flags_local must be an 8-byte aligned device pointer");
    if (hd.epoch == 0) return set_error(OSSG_EINVAL, "epoch must start at 1 and grow by one per call");
    const int nb[4] = { hd.north, hd.south, hd.east, hd.west };
    HaloBatch b;
    b.halo = hd;
    b.n = nfields;
    b.partners = 0;
    for (int k = (hd.phase ? 2 : 0); k < (hd.phase ? 4 : 2) && hd.phase != 2; ++k) {
        if (nb[k] < -1 || nb[k] >= hd.nranks) return set_error(OSSG_EINVAL, "neighbour id %d out of range", nb[k]);
        if (nb[k] >= 0 && nb[k] != hd.me) {
            if (!hd.flags_peer[nb[k]]) return set_error(OSSG_EINVAL, "flags_peer[%d] is NULL", nb[k]);
            b.partners |= 1u << nb[k];
        }
    }
    long long items = 0;
    int chunks = 0;
    for (int k = 0; k < nfields; ++k) {
        b.hf[k] = fields[k];
        const ossg_field_desc& f = b.hf[k].f;
        if (!f.data) return set_error(OSSG_EINVAL, "field %d: data is NULL", k);
        if (f.elem_type != OSSG_F64 && f.elem_type != OSSG_F32) return set_error(OSSG_EINVAL, "field %d: bad elem_type", k);
        if (f.Nx < 1 || f.Ny < 1 || f.Nz < 1 || f.Hx < 0 || f.Hy < 0 || f.Hz < 0) return set_error(OSSG_EINVAL, "field %d: bad sizes", k);
        if (f.Hy > f.Ny || f.Hx > f.Nx) return set_error(OSSG_EHALO, "field %d: halo (%d, %d) larger than the tile (%d, %d)", k, f.Hx, f.Hy, f.Nx, f.Ny);
        if (k && (f.Nx != b.hf[0].f.Nx || f.Ny != b.hf[0].f.Ny || f.Nz != b.hf[0].f.Nz || f.Hx != b.hf[0].f.Hx || f.Hy != b.hf[0].f.Hy || f.Hz != b.hf[0].f.Hz))
            return set_error(OSSG_EINVAL, "field %d: all fields of a batch must share size and halo", k);
        if (hd.phase == 0) {
            if ((hd.north >= 0 && !b.hf[k].north) || (hd.south >= 0 && !b.hf[k].south)) return set_error(OSSG_EINVAL, "field %d: neighbour pointer is NULL", k);
            if (hd.north < 0) b.hf[k].north = nullptr;
            if (hd.south < 0) b.hf[k].south = nullptr;
        } else {
            if ((hd.east >= 0 && !b.hf[k].east) || (hd.west >= 0 && !b.hf[k].west)) return set_error(OSSG_EINVAL, "field %d: neighbour pointer is NULL", k);
            if (hd.east < 0) b.hf[k].east = nullptr;
            if (hd.west < 0) b.hf[k].west = nullptr;
        }
    }
    {
        const ossg_field_desc& f = b.hf[0].f;
        if (hd.phase == 0) {
            chunks = ceil_div(f.Nx + 2 * f.Hx, PB);
            items = 2LL * f.Hy * (f.Nz + 2 * f.Hz) * chunks;
        } else {
            const long long rows = (long long)(f.Ny + 2 * f.Hy) * (f.Nz + 2 * f.Hz);
            chunks = (int)((rows + PB - 1) / PB);
            items = 2LL * chunks;
        }
    }
    if (items == 0) return 0;
    b.items_per_field = items;
    b.chunks = chunks;
    b.timeout_ns = peer_timeout_ns();
    int dev = 0, sms = 0;
    OSSG_CUDA_TRY(cudaGetDevice(&dev));
    OSSG_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long resident = (long long)sms * PEER_BLOCKS_PER_SM, total = items * nfields;
    halo_peer_kernel<<<(int)(total < resident ? total : resident), PB, 0, as_stream(stream)>>>(b);
    OSSG_LAUNCH_CHECK("halo_peer_kernel");
    return 0;
}
