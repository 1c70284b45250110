// csrc/ossg_nccl.cu -- fold-partner exchange and ordinary halo exchange of an (x, y)-partitioned
// tripolar field through NCCL send/recv (north_star: "ordinary halos and the fold-partner exchange
// between mirror ranks in the top row go through NCCL send/recv over NVLink, replacing MPI").
//
// Reference: the fold is src/zipper_boundary_condition.jl:70-155; the reference itself never splits x
// (src/distributed_tripolar_grid.jl:28-31) and leaves the south/north slabs to Oceananigans' MPI
// halo passing.  Here one exchange is
//     ONE pack kernel  ->  ONE ncclGroup of ncclSend/ncclRecv (one message per partner)  ->  ONE unpack kernel
// on the caller's stream, with the caller's communicator and a caller-provided persistent buffer: no
// allocation, no host synchronisation, nothing but plain device pointers across the ABI.
//
// NCCL is resolved at run time (dlopen of the libnccl.so.2 already in the process, e.g. the one
// NCCL.jl or torch loaded), so libossg_b200.so itself has no link-time dependency on it.
#include <dlfcn.h>
#include <mutex>
#include <cstring>

#include "ossg_common.cuh"

namespace ossg {

// ------------------------------------------------------------------ NCCL entry points (minimal declarations)
typedef struct ncclComm* nccl_comm_t;
typedef struct { char internal[128]; } nccl_unique_id;
struct NcclApi {
    int (*GetUniqueId)(nccl_unique_id*);
    int (*CommInitRank)(nccl_comm_t*, int, nccl_unique_id, int);
    int (*CommDestroy)(nccl_comm_t);
    int (*GroupStart)();
    int (*GroupEnd)();
    int (*Send)(const void*, size_t, int, int, nccl_comm_t, cudaStream_t);
    int (*Recv)(void*, size_t, int, int, nccl_comm_t, cudaStream_t);
    const char* (*GetErrorString)(int);
    bool ok;
};

static const NcclApi* nccl_api()
{
    static NcclApi api = {};
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);        // the copy the host program already uses
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        api.GetUniqueId = (int (*)(nccl_unique_id*))dlsym(h, "ncclGetUniqueId");
        api.CommInitRank = (int (*)(nccl_comm_t*, int, nccl_unique_id, int))dlsym(h, "ncclCommInitRank");
        api.CommDestroy = (int (*)(nccl_comm_t))dlsym(h, "ncclCommDestroy");
        api.GroupStart = (int (*)())dlsym(h, "ncclGroupStart");
        api.GroupEnd = (int (*)())dlsym(h, "ncclGroupEnd");
        api.Send = (int (*)(const void*, size_t, int, int, nccl_comm_t, cudaStream_t))dlsym(h, "ncclSend");
        api.Recv = (int (*)(void*, size_t, int, int, nccl_comm_t, cudaStream_t))dlsym(h, "ncclRecv");
        api.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.GroupStart && api.GroupEnd && api.Send && api.Recv;
    });
    return api.ok ? &api : nullptr;
}

static int nccl_fail(const NcclApi* n, int rc, const char* what)
{
    return set_error(OSSG_ENCCL, "%s: ncclResult_t %d (%s)", what, rc, n && n->GetErrorString ? n->GetErrorString(rc) : "?");
}
#define OSSG_NCCL_TRY(n, expr) do { int _r = (expr); if (_r != 0) return nccl_fail(n, _r, #expr); } while (0)

// ------------------------------------------------------------------ partition arithmetic (host)
// Equal partition, the last rank takes the remainder (external: Oceananigans local_size;
// src/distributed_tripolar_grid.jl:41-49).  Rank r sits at (rx, ry) = (r / Ry, r % Ry).
static inline int tile_n(int N, int R, int r) { return r == R - 1 ? N - (N / R) * (R - 1) : N / R; }
static inline int tile_0(int N, int R, int r) { return 1 + (N / R) * r; }

constexpr int MAX_XFIELDS = OSSG_MAX_PEER_BATCH;     // fields per exchange call
constexpr int MAX_SEGS = 4 * MAX_XFIELDS;            // a field meets at most 3 partners per direction (mirror rank, its neighbour, the wrap)

// One contiguous run of fold source columns travelling from rank `src` to rank `dst`.
// Buffer layout of a segment (as ossg_fold_pack): [level][row r = 0..rows-1][column ascending],
// row r < Hy: the source row of halo row r+1, row r == Hy: row Ny (the y-Center overwrite, only
// when the segment serves destination columns > Nx/2).
struct FoldSeg {
    int field, src, dst;           // field index, x-indices of the two ranks in the top row
    int src_i0, n;                 // global source columns [src_i0, src_i0 + n)
    int rows, wrapped;             // rows per level; wrapped: the x-Face column Nx+1 -> 1 (sign |s|, destination = source column)
    unsigned long long off;        // byte offset of the segment in its message area
};

// All segments of one exchange in the canonical order (field, destination rank, range, source rank);
// identical on every rank.  Returns the count (<= cap) or -1.
static int fold_segments(const ossg_field_desc* f, int nf, int gNx, int Rx, FoldSeg* out, int cap)
{
    int ns = 0;
    for (int k = 0; k < nf; ++k) {
        const int locx = f[k].locx;
        for (int d = 0; d < Rx; ++d) {
            const int lo = tile_0(gNx, Rx, d), hi = lo + tile_n(gNx, Rx, d) - 1;           // destination columns
            const int a = gNx - hi + 1 + locx, b = gNx - lo + 1 + locx;                    // source columns, ascending
            int ra[2], rb[2], wr[2], nr = 0;
            if (a <= gNx) { ra[nr] = a; rb[nr] = b < gNx ? b : gNx; wr[nr] = 0; ++nr; }
            if (b > gNx) { ra[nr] = 1; rb[nr] = b - gNx; wr[nr] = 1; ++nr; }               // periodic wrap of the x-Face fold
            for (int q = 0; q < nr; ++q)
                for (int s = 0; s < Rx; ++s) {
                    const int slo = tile_0(gNx, Rx, s), shi = slo + tile_n(gNx, Rx, s) - 1;
                    const int c0 = ra[q] > slo ? ra[q] : slo, c1 = rb[q] < shi ? rb[q] : shi;
                    if (c0 > c1) continue;
                    if (ns >= cap) return -1;
                    FoldSeg& g = out[ns++];
                    g.field = k; g.src = s; g.dst = d; g.src_i0 = c0; g.n = c1 - c0 + 1; g.wrapped = wr[q];
                    // largest destination column served: source c0 (the map reverses); row Ny only for the right half
                    const int dmax = wr[q] ? c1 : gNx + 1 + locx - c0;
                    g.rows = f[k].Hy + ((f[k].locy == OSSG_CENTER && dmax > gNx / 2) ? 1 : 0);
                    g.off = 0;
                }
        }
    }
    return ns;
}

static inline size_t seg_bytes(const FoldSeg& g, const ossg_field_desc& f)
{
    return (size_t)f.nk * g.rows * g.n * (f.elem_type == OSSG_F64 ? 8 : 4);
}
static inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

// ------------------------------------------------------------------ device: fold pack / unpack over a segment list
struct FoldSegArgs {
    ossg_field_desc f[MAX_XFIELDS];
    FoldSeg seg[MAX_SEGS];
    int nseg, gNx, i0_me;
    char* area;                    // message area the offsets refer to
};

template <typename T>
__device__ __forceinline__ void seg_pack(const ossg_field_desc& f, const FoldSeg& g, int i0_me, T* buf, long long e)
{
    const int c = (int)(e % g.n);
    const long long rest = e / g.n;
    const int r = (int)(rest % g.rows), kk = (int)(rest / g.rows);
    const size_t sx = (size_t)(f.Nx + 2 * f.Hx), sxy = sx * (size_t)(f.Ny + 2 * f.Hy);
    const int li = g.src_i0 + c - i0_me + 1;                                   // local 1-based source column
    const int js = (r < f.Hy) ? (f.locy ? f.Ny - r : f.Ny - r - 1) : f.Ny;     // zipper_boundary_condition.jl:79,96,113,129
    buf[e] = static_cast<const T*>(f.data)[(size_t)(f.k0 + kk) * sxy + (size_t)(js + f.Hy - 1) * sx + (li + f.Hx - 1)];
}

template <typename T>
__device__ __forceinline__ void seg_unpack(const ossg_field_desc& f, const FoldSeg& g, int gNx, int i0_me, const T* buf, long long e)
{
    const int c = (int)(e % g.n);
    const long long rest = e / g.n;
    const int r = (int)(rest % g.rows), kk = (int)(rest / g.rows);
    const int cs = g.src_i0 + c;                                               // global source column
    const int i = g.wrapped ? cs : gNx + 1 + f.locx - cs;                      // global destination column (:73-75, :90-92, :110, :125)
    const double s = g.wrapped ? fabs(f.sign) : f.sign;
    int jd;
    if (r < f.Hy) jd = f.Ny + r + 1;
    else { if (i <= gNx / 2) return; jd = f.Ny; }                              // :102, :135
    const size_t sx = (size_t)(f.Nx + 2 * f.Hx), sxy = sx * (size_t)(f.Ny + 2 * f.Hy);
    const int li = i - i0_me + 1;
    static_cast<T*>(f.data)[(size_t)(f.k0 + kk) * sxy + (size_t)(jd + f.Hy - 1) * sx + (li + f.Hx - 1)] = apply_sign(buf[e], s);
}

template <bool PACK>
__global__ void __launch_bounds__(256) fold_segments_kernel(const __grid_constant__ FoldSegArgs a)
{
    const FoldSeg& g = a.seg[blockIdx.y];
    const ossg_field_desc& f = a.f[g.field];
    const long long total = (long long)f.nk * g.rows * g.n;
    char* buf = a.area + g.off;
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        if (f.elem_type == OSSG_F64) {
            if (PACK) seg_pack<double>(f, g, a.i0_me, reinterpret_cast<double*>(buf), e);
            else seg_unpack<double>(f, g, a.gNx, a.i0_me, reinterpret_cast<const double*>(buf), e);
        } else {
            if (PACK) seg_pack<float>(f, g, a.i0_me, reinterpret_cast<float*>(buf), e);
            else seg_unpack<float>(f, g, a.gNx, a.i0_me, reinterpret_cast<const float*>(buf), e);
        }
    }
}

// ------------------------------------------------------------------ device: box pack / unpack (ordinary halos)
// A box is [i0, i0+ni) x [j0, j0+nj) x all parent levels of a field (parent indices); its buffer
// image is dense, i fastest.
struct Box {
    int field, i0, j0, ni, nj;
    unsigned long long off;
};
constexpr int MAX_BOXES = 2 * MAX_XFIELDS;
struct BoxArgs {
    ossg_field_desc f[MAX_XFIELDS];
    Box box[MAX_BOXES];
    int nbox;
    char* area;
};

template <bool PACK>
__global__ void __launch_bounds__(256) boxes_kernel(const __grid_constant__ BoxArgs a)
{
    const Box& b = a.box[blockIdx.y];
    const ossg_field_desc& f = a.f[b.field];
    const int sz = f.Nz + 2 * f.Hz;
    const size_t sx = (size_t)(f.Nx + 2 * f.Hx), sxy = sx * (size_t)(f.Ny + 2 * f.Hy);
    const long long per_level = (long long)b.ni * b.nj, total = per_level * sz;
    char* buf = a.area + b.off;
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const int k = (int)(e / per_level);
        const long long r = e - (long long)k * per_level;
        const int j = (int)(r / b.ni), i = (int)(r - (long long)j * b.ni);
        const size_t o = (size_t)k * sxy + (size_t)(b.j0 + j) * sx + (size_t)(b.i0 + i);
        if (f.elem_type == OSSG_F64) {
            if (PACK) reinterpret_cast<double*>(buf)[e] = static_cast<const double*>(f.data)[o];
            else static_cast<double*>(f.data)[o] = reinterpret_cast<const double*>(buf)[e];
        } else {
            if (PACK) reinterpret_cast<float*>(buf)[e] = static_cast<const float*>(f.data)[o];
            else static_cast<float*>(f.data)[o] = reinterpret_cast<const float*>(buf)[e];
        }
    }
}

// local periodic west/east copy (Rx == 1), whole parent extent
__global__ void __launch_bounds__(256) periodic_local_kernel(const __grid_constant__ BoxArgs a)
{
    const ossg_field_desc& f = a.f[blockIdx.y];
    const long long rows = (long long)(f.Ny + 2 * f.Hy) * (f.Nz + 2 * f.Hz);
    const size_t sx = (size_t)(f.Nx + 2 * f.Hx);
    for (long long t = (long long)blockIdx.x * 256 + threadIdx.x; t < 2 * rows; t += (long long)gridDim.x * 256) {
        const long long row = t >> 1;
        const int side = (int)(t & 1);
        if (f.elem_type == OSSG_F64) {
            double* c = static_cast<double*>(f.data) + (size_t)row * sx;
            copy_halo_piece(side ? c + f.Nx + f.Hx : c, side ? c + f.Hx : c + f.Nx, f.Hx);
        } else {
            float* c = static_cast<float*>(f.data) + (size_t)row * sx;
            copy_halo_piece(side ? c + f.Nx + f.Hx : c, side ? c + f.Hx : c + f.Nx, f.Hx);
        }
    }
}

// ------------------------------------------------------------------ host: argument checks shared by the entry points
static int check_exchange(const ossg_field_desc* fields, int nfields, const ossg_exchange_layout* lay, ossg_field_desc* f, bool need_data = true)
{
    if (!fields || !lay || nfields < 1) return set_error(OSSG_EINVAL, "fields/layout is NULL or nfields < 1");
    if (nfields > MAX_XFIELDS) return set_error(OSSG_EINVAL, "at most %d fields per exchange, got %d", MAX_XFIELDS, nfields);
    if (lay->Rx < 1 || lay->Ry < 1 || lay->rank < 0 || lay->rank >= lay->Rx * lay->Ry)
        return set_error(OSSG_EINVAL, "bad layout: rank %d of a %d x %d partition", lay->rank, lay->Rx, lay->Ry);
    if (lay->gNx < 2 || lay->gNx % 2) return set_error(OSSG_EODD_NX, "global Nx must be even, got %d", lay->gNx);
    if (lay->gNx / lay->Rx < 1 || lay->gNy / lay->Ry < 1) return set_error(OSSG_EWINDOW, "more ranks than cells in a direction");
    const int rx = lay->rank / lay->Ry, ry = lay->rank % lay->Ry;
    const int nx = tile_n(lay->gNx, lay->Rx, rx), ny = tile_n(lay->gNy, lay->Ry, ry);
    for (int k = 0; k < nfields; ++k) {
        f[k] = fields[k];
        if (need_data && !f[k].data) return set_error(OSSG_EINVAL, "field %d: data is NULL", k);
        if (f[k].elem_type != OSSG_F64 && f[k].elem_type != OSSG_F32) return set_error(OSSG_EINVAL, "field %d: bad elem_type", k);
        if (f[k].Nx != nx || f[k].Ny != ny)
            return set_error(OSSG_EWINDOW, "field %d: local size (%d, %d) differs from this rank's tile (%d, %d)", k, f[k].Nx, f[k].Ny, nx, ny);
        if (f[k].Nz < 1 || f[k].Hx < 0 || f[k].Hy < 0 || f[k].Hz < 0) return set_error(OSSG_EINVAL, "field %d: bad sizes", k);
        if ((f[k].locx != OSSG_CENTER && f[k].locx != OSSG_FACE) || (f[k].locy != OSSG_CENTER && f[k].locy != OSSG_FACE))
            return set_error(OSSG_ELOCATION, "field %d: location codes must be OSSG_CENTER or OSSG_FACE", k);
        if (f[k].nk == 0) { f[k].k0 = f[k].Hz; f[k].nk = f[k].Nz; }
        if (f[k].k0 < 0 || f[k].nk < 1 || f[k].k0 + f[k].nk > f[k].Nz + 2 * f[k].Hz) return set_error(OSSG_EINVAL, "field %d: bad k-range", k);
    }
    return 0;
}

// message areas of the fold: per partner, the segments in canonical order; returns the bytes of the
// send and of the receive area (each partner's message 16-byte aligned) and fills seg[].off
struct FoldPlan {
    FoldSeg send[MAX_SEGS], recv[MAX_SEGS], self[MAX_SEGS];
    int nsend, nrecv, nself;
    size_t peer_send_off[OSSG_MAX_PEERS + 1], peer_recv_off[OSSG_MAX_PEERS + 1];   // message of partner p: [off[p], off[p+1])
    size_t send_bytes, recv_bytes, self_bytes;
};

static int plan_fold(const ossg_field_desc* f, int nf, const ossg_exchange_layout& lay, FoldPlan& P)
{
    if (lay.Rx > OSSG_MAX_PEERS) return set_error(OSSG_EINVAL, "at most %d ranks in x, got %d", OSSG_MAX_PEERS, lay.Rx);
    static thread_local FoldSeg all[MAX_SEGS * OSSG_MAX_PEERS];
    const int ns = fold_segments(f, nf, lay.gNx, lay.Rx, all, MAX_SEGS * OSSG_MAX_PEERS);
    if (ns < 0) return set_error(OSSG_EINVAL, "too many fold segments");
    const int me = lay.rank / lay.Ry;
    P.nsend = P.nrecv = P.nself = 0;
    size_t so = 0, ro = 0, fo = 0;
    for (int p = 0; p < lay.Rx; ++p) {                       // partner-major, canonical order inside
        P.peer_send_off[p] = so; P.peer_recv_off[p] = ro;
        for (int q = 0; q < ns; ++q) {
            const FoldSeg& g = all[q];
            if (g.src == me && g.dst == me) {
                if (p != me) continue;
                if (P.nself >= MAX_SEGS) return set_error(OSSG_EINVAL, "too many fold segments");
                P.self[P.nself] = g; P.self[P.nself].off = fo; fo += align16(seg_bytes(g, f[g.field])); ++P.nself;
            } else if (g.src == me && g.dst == p) {
                if (P.nsend >= MAX_SEGS) return set_error(OSSG_EINVAL, "too many fold segments");
                P.send[P.nsend] = g; P.send[P.nsend].off = so; so += align16(seg_bytes(g, f[g.field])); ++P.nsend;
            } else if (g.dst == me && g.src == p) {
                if (P.nrecv >= MAX_SEGS) return set_error(OSSG_EINVAL, "too many fold segments");
                P.recv[P.nrecv] = g; P.recv[P.nrecv].off = ro; ro += align16(seg_bytes(g, f[g.field])); ++P.nrecv;
            }
        }
    }
    P.peer_send_off[lay.Rx] = so; P.peer_recv_off[lay.Rx] = ro;
    P.send_bytes = so; P.recv_bytes = ro; P.self_bytes = fo;
    return 0;
}

template <bool PACK>
static int launch_segments(const ossg_field_desc* f, int nf, const FoldSeg* seg, int nseg, const ossg_exchange_layout& lay, char* area, cudaStream_t s)
{
    if (nseg == 0) return 0;
    FoldSegArgs a;
    for (int k = 0; k < nf; ++k) a.f[k] = f[k];
    long long most = 0;
    for (int q = 0; q < nseg; ++q) {
        a.seg[q] = seg[q];
        const long long e = (long long)f[seg[q].field].nk * seg[q].rows * seg[q].n;
        most = most > e ? most : e;
    }
    a.nseg = nseg; a.gNx = lay.gNx; a.i0_me = tile_0(lay.gNx, lay.Rx, lay.rank / lay.Ry); a.area = area;
    const long long want = (most + 255) / 256;
    const unsigned gx = (unsigned)(want < 1 ? 1 : want > 4096 ? 4096 : want);
    fold_segments_kernel<PACK><<<dim3(gx, nseg), 256, 0, s>>>(a);
    OSSG_LAUNCH_CHECK(PACK ? "fold_segments_kernel<pack>" : "fold_segments_kernel<unpack>");
    return 0;
}

}  // namespace ossg

using namespace ossg;

// ------------------------------------------------------------------ C ABI
extern "C" int ossg_nccl_get_unique_id(unsigned char id[128])
{
    if (!id) return set_error(OSSG_EINVAL, "id is NULL");
    const NcclApi* n = nccl_api();
    if (!n) return set_error(OSSG_ENCCL, "libnccl.so.2 could not be loaded");
    nccl_unique_id u;
    OSSG_NCCL_TRY(n, n->GetUniqueId(&u));
    std::memcpy(id, u.internal, 128);
    return 0;
}

extern "C" int ossg_nccl_comm_init_rank(void** comm, int nranks, const unsigned char id[128], int rank)
{
    if (!comm || !id) return set_error(OSSG_EINVAL, "NULL argument");
    const NcclApi* n = nccl_api();
    if (!n) return set_error(OSSG_ENCCL, "libnccl.so.2 could not be loaded");
    nccl_unique_id u;
    std::memcpy(u.internal, id, 128);
    nccl_comm_t c = nullptr;
    OSSG_NCCL_TRY(n, n->CommInitRank(&c, nranks, u, rank));
    *comm = c;
    return 0;
}

extern "C" int ossg_nccl_comm_destroy(void* comm)
{
    const NcclApi* n = nccl_api();
    if (!n) return set_error(OSSG_ENCCL, "libnccl.so.2 could not be loaded");
    if (comm) OSSG_NCCL_TRY(n, n->CommDestroy(static_cast<nccl_comm_t>(comm)));
    return 0;
}

extern "C" int ossg_fold_exchange_plan(const ossg_field_desc* fields, int nfields, const ossg_exchange_layout* lay,
                                       ossg_fold_segment* out, int capacity)
{
    ossg_field_desc f[MAX_XFIELDS];
    int rc = check_exchange(fields, nfields, lay, f, false);
    if (rc) return rc;
    static thread_local FoldSeg all[MAX_SEGS * OSSG_MAX_PEERS];
    const int ns = fold_segments(f, nfields, lay->gNx, lay->Rx, all, MAX_SEGS * OSSG_MAX_PEERS);
    if (ns < 0) return set_error(OSSG_EINVAL, "too many fold segments");
    for (int q = 0; q < ns && q < capacity && out; ++q) {
        out[q].field = all[q].field; out[q].src_rank = all[q].src * lay->Ry + lay->Ry - 1; out[q].dst_rank = all[q].dst * lay->Ry + lay->Ry - 1;
        out[q].src_i0 = all[q].src_i0; out[q].n = all[q].n; out[q].rows = all[q].rows; out[q].wrapped = all[q].wrapped; out[q].reserved = 0;
    }
    return ns;
}

// ---- fold-partner exchange ---------------------------------------------------------------------
static int fold_needs(const ossg_field_desc* f, int nf, const ossg_exchange_layout& lay)
{
    if (lay.rank % lay.Ry != lay.Ry - 1) return 0;                 // ranks below the top row take no part
    for (int k = 0; k < nf; ++k)
        if (f[k].Hy > f[k].Ny - 1) return set_error(OSSG_EHALO, "field %d: the fold needs Hy <= ny - 1 on the top row (Hy = %d, ny = %d)", k, f[k].Hy, f[k].Ny);
    return 1;
}

extern "C" size_t ossg_fold_exchange_workspace_bytes(const ossg_field_desc* fields, int nfields, const ossg_exchange_layout* lay)
{
    ossg_field_desc f[MAX_XFIELDS];
    if (check_exchange(fields, nfields, lay, f, false)) return 0;
    if (fold_needs(f, nfields, *lay) != 1) return 0;
    static thread_local FoldPlan P;
    if (plan_fold(f, nfields, *lay, P)) return 0;
    return align16(P.send_bytes) + align16(P.recv_bytes) + align16(P.self_bytes);
}

extern "C" int ossg_fold_exchange(void* nccl_comm, const ossg_field_desc* fields, int nfields, const ossg_exchange_layout* lay,
                                  void* workspace, size_t workspace_bytes, void* stream)
{
    ossg_field_desc f[MAX_XFIELDS];
    int rc = check_exchange(fields, nfields, lay, f);
    if (rc) return rc;
    rc = fold_needs(f, nfields, *lay);
    if (rc <= 0) return rc;
    static thread_local FoldPlan P;
    if ((rc = plan_fold(f, nfields, *lay, P))) return rc;
    const size_t need = align16(P.send_bytes) + align16(P.recv_bytes) + align16(P.self_bytes);
    if (need > 0 && (!workspace || reinterpret_cast<uintptr_t>(workspace) % 16)) return set_error(OSSG_EALIGN, "workspace must be a 16-byte aligned device pointer");
    if (workspace_bytes < need) return set_error(OSSG_EWORKSPACE, "workspace has %zu bytes, need %zu", workspace_bytes, need);
    const bool remote = P.nsend > 0 || P.nrecv > 0;
    const NcclApi* n = nullptr;
    if (remote) {
        if (!nccl_comm) return set_error(OSSG_EINVAL, "nccl_comm is NULL but the fold crosses ranks");
        if (!(n = nccl_api())) return set_error(OSSG_ENCCL, "libnccl.so.2 could not be loaded");
    }
    cudaStream_t s = as_stream(stream);
    char* send = static_cast<char*>(workspace);
    char* recv = send + align16(P.send_bytes);
    char* self = recv + align16(P.recv_bytes);
    // pack: what the partners need from me, and my own (self-mirrored) part
    if ((rc = launch_segments<true>(f, nfields, P.send, P.nsend, *lay, send, s))) return rc;
    if ((rc = launch_segments<true>(f, nfields, P.self, P.nself, *lay, self, s))) return rc;
    if (remote) {
        const int Ry = lay->Ry, top = Ry - 1;
        nccl_comm_t comm = static_cast<nccl_comm_t>(nccl_comm);
        OSSG_NCCL_TRY(n, n->GroupStart());
        for (int p = 0; p < lay->Rx; ++p) {
            const size_t sb = P.peer_send_off[p + 1] - P.peer_send_off[p], rb = P.peer_recv_off[p + 1] - P.peer_recv_off[p];
            if (sb) OSSG_NCCL_TRY(n, n->Send(send + P.peer_send_off[p], sb, /*ncclChar*/ 0, p * Ry + top, comm, s));
            if (rb) OSSG_NCCL_TRY(n, n->Recv(recv + P.peer_recv_off[p], rb, /*ncclChar*/ 0, p * Ry + top, comm, s));
        }
        OSSG_NCCL_TRY(n, n->GroupEnd());
    }
    if ((rc = launch_segments<false>(f, nfields, P.self, P.nself, *lay, self, s))) return rc;
    return launch_segments<false>(f, nfields, P.recv, P.nrecv, *lay, recv, s);
}

// ---- ordinary halos: south/north rows, then west/east columns over the whole y extent ---------------
struct HaloPlan {
    Box send[2][MAX_BOXES], recv[2][MAX_BOXES];       // [phase]
    int nsend[2], nrecv[2];
    int peer[2][2];                                   // [phase][0: north / east, 1: south / west]; -1: none
    size_t send_off[2][3], recv_off[2][3];            // message to / from peer[phase][d]: [off[d], off[d+1])
    size_t bytes;                                     // send + recv area of the larger phase
};

static int plan_halo(const ossg_field_desc* f, int nf, const ossg_exchange_layout& lay, HaloPlan& H)
{
    const int Rx = lay.Rx, Ry = lay.Ry, rx = lay.rank / Ry, ry = lay.rank % Ry;
    H.peer[0][0] = ry + 1 < Ry ? rx * Ry + ry + 1 : -1;              // north
    H.peer[0][1] = ry > 0 ? rx * Ry + ry - 1 : -1;                   // south
    H.peer[1][0] = Rx > 1 ? ((rx + 1) % Rx) * Ry + ry : -1;          // east (periodic)
    H.peer[1][1] = Rx > 1 ? ((rx - 1 + Rx) % Rx) * Ry + ry : -1;     // west
    size_t worst = 0;
    for (int ph = 0; ph < 2; ++ph) {
        H.nsend[ph] = H.nrecv[ph] = 0;
        size_t so = 0, ro = 0;
        for (int d = 0; d < 2; ++d) {
            H.send_off[ph][d] = so; H.recv_off[ph][d] = ro;
            if (H.peer[ph][d] < 0) continue;
            for (int k = 0; k < nf; ++k) {
                const ossg_field_desc& g = f[k];
                const int sx = g.Nx + 2 * g.Hx, sy = g.Ny + 2 * g.Hy, sz = g.Nz + 2 * g.Hz;
                const size_t esz = g.elem_type == OSSG_F64 ? 8 : 4;
                Box bs, br;
                bs.field = br.field = k;
                if (ph == 0) {                               // whole rows (x halos included): Hy interior rows out, Hy halo rows in
                    if (g.Hy > g.Ny) return set_error(OSSG_EHALO, "field %d: Hy = %d exceeds the tile's %d rows", k, g.Hy, g.Ny);
                    bs.i0 = br.i0 = 0; bs.ni = br.ni = sx; bs.nj = br.nj = g.Hy;
                    bs.j0 = d == 0 ? g.Ny : g.Hy;            // to the north: top interior rows; to the south: bottom interior rows
                    br.j0 = d == 0 ? g.Ny + g.Hy : 0;        // from the north: my north halo; from the south: my south halo
                } else {                                     // whole columns (y halos included, so the corners carry exchanged / folded rows)
                    if (g.Hx > g.Nx) return set_error(OSSG_EHALO, "field %d: Hx = %d exceeds the tile's %d columns", k, g.Hx, g.Nx);
                    bs.j0 = br.j0 = 0; bs.nj = br.nj = sy; bs.ni = br.ni = g.Hx;
                    bs.i0 = d == 0 ? g.Nx : g.Hx;            // to the east: east interior columns; to the west: west interior columns
                    br.i0 = d == 0 ? g.Nx + g.Hx : 0;        // from the east: my east halo; from the west: my west halo
                }
                const size_t nb = align16((size_t)bs.ni * bs.nj * sz * esz);
                bs.off = so; so += nb;
                br.off = ro; ro += nb;
                H.send[ph][H.nsend[ph]++] = bs;
                H.recv[ph][H.nrecv[ph]++] = br;
            }
        }
        H.send_off[ph][2] = so; H.recv_off[ph][2] = ro;
        worst = worst > so + ro ? worst : so + ro;
    }
    H.bytes = worst;
    return 0;
}

extern "C" size_t ossg_halo_exchange_workspace_bytes(const ossg_field_desc* fields, int nfields, const ossg_exchange_layout* lay)
{
    ossg_field_desc f[MAX_XFIELDS];
    if (check_exchange(fields, nfields, lay, f, false)) return 0;
    static thread_local HaloPlan H;
    if (plan_halo(f, nfields, *lay, H)) return 0;
    return H.bytes;
}

extern "C" int ossg_halo_exchange(void* nccl_comm, const ossg_field_desc* fields, int nfields, const ossg_exchange_layout* lay,
                                  void* workspace, size_t workspace_bytes, void* stream)
{
    ossg_field_desc f[MAX_XFIELDS];
    int rc = check_exchange(fields, nfields, lay, f);
    if (rc) return rc;
    static thread_local HaloPlan H;
    if ((rc = plan_halo(f, nfields, *lay, H))) return rc;
    if (H.bytes > 0 && (!workspace || reinterpret_cast<uintptr_t>(workspace) % 16)) return set_error(OSSG_EALIGN, "workspace must be a 16-byte aligned device pointer");
    if (workspace_bytes < H.bytes) return set_error(OSSG_EWORKSPACE, "workspace has %zu bytes, need %zu", workspace_bytes, H.bytes);
    cudaStream_t s = as_stream(stream);
    const NcclApi* n = nullptr;
    if (H.bytes > 0) {
        if (!nccl_comm) return set_error(OSSG_EINVAL, "nccl_comm is NULL but the halos cross ranks");
        if (!(n = nccl_api())) return set_error(OSSG_ENCCL, "libnccl.so.2 could not be loaded");
    }
    nccl_comm_t comm = static_cast<nccl_comm_t>(nccl_comm);
    BoxArgs a;
    for (int k = 0; k < nfields; ++k) a.f[k] = f[k];
    for (int ph = 0; ph < 2; ++ph) {
        if (ph == 1 && lay->Rx == 1) {                       // x not partitioned: the local periodic copy
            a.nbox = 0; a.area = nullptr;
            long long most = 0;
            for (int k = 0; k < nfields; ++k) {
                const long long rows = (long long)(f[k].Ny + 2 * f[k].Hy) * (f[k].Nz + 2 * f[k].Hz);
                most = most > rows ? most : rows;
            }
            bool any = false;
            for (int k = 0; k < nfields; ++k) any = any || f[k].Hx > 0;
            if (any) {
                periodic_local_kernel<<<dim3((unsigned)((2 * most + 255) / 256), nfields), 256, 0, s>>>(a);
                OSSG_LAUNCH_CHECK("periodic_local_kernel");
            }
            continue;
        }
        if (H.nsend[ph] == 0) continue;
        char* send = static_cast<char*>(workspace);
        char* recv = send + H.send_off[ph][2];
        auto launch = [&](const Box* boxes, int nb, char* area, bool pack) -> int {
            a.nbox = nb; a.area = area;
            long long most = 0;
            for (int q = 0; q < nb; ++q) {
                a.box[q] = boxes[q];
                const long long e = (long long)boxes[q].ni * boxes[q].nj * (f[boxes[q].field].Nz + 2 * f[boxes[q].field].Hz);
                most = most > e ? most : e;
            }
            const long long want = (most + 255) / 256;
            const unsigned gx = (unsigned)(want < 1 ? 1 : want > 4096 ? 4096 : want);
            if (pack) boxes_kernel<true><<<dim3(gx, nb), 256, 0, s>>>(a);
            else boxes_kernel<false><<<dim3(gx, nb), 256, 0, s>>>(a);
            OSSG_LAUNCH_CHECK("boxes_kernel");
            return 0;
        };
        if ((rc = launch(H.send[ph], H.nsend[ph], send, true))) return rc;
        OSSG_NCCL_TRY(n, n->GroupStart());
        // to the north / east first, then to the south / west; the receives mirror that order, so two
        // messages between the same pair of ranks (Rx == 2 or Ry == 2 ... periodic) match up: my
        // message "to the east" is the partner's message "from the west"
        for (int d = 0; d < 2; ++d) {
            const size_t sb = H.send_off[ph][d + 1] - H.send_off[ph][d];
            if (H.peer[ph][d] >= 0 && sb) OSSG_NCCL_TRY(n, n->Send(send + H.send_off[ph][d], sb, 0, H.peer[ph][d], comm, s));
        }
        for (int d = 1; d >= 0; --d) {
            const size_t rb = H.recv_off[ph][d + 1] - H.recv_off[ph][d];
            if (H.peer[ph][d] >= 0 && rb) OSSG_NCCL_TRY(n, n->Recv(recv + H.recv_off[ph][d], rb, 0, H.peer[ph][d], comm, s));
        }
        OSSG_NCCL_TRY(n, n->GroupEnd());
        if ((rc = launch(H.recv[ph], H.nrecv[ph], recv, false))) return rc;
    }
    return 0;
}
