/*
 * include/ossg_b200.h -- C ABI of libossg_b200.so
 *
 * B200 (sm_100a) implementation of the data-parallel hot path of
 * CliMA/OrthogonalSphericalShellGrids.jl v0.2.1: TripolarGrid coordinate + metric generation and
 * the zipper (north-fold) halo fill.  The reference has no FFI for this path (it is pure Julia
 * multiple dispatch); each entry point below names the reference method whose body it replaces
 * when called through `ccall` (see INTEGRATION.md for the Julia shim).
 *
 * Conventions
 *  - Every pointer to array data is a DEVICE pointer owned by the caller.  The library never
 *    allocates, frees, copies to or from host memory on the caller's behalf, and never
 *    synchronises the device; all work is enqueued on the `stream` argument (a cudaStream_t
 *    passed as void*; NULL = legacy default stream).
 *  - Arrays use the reference's memory layout: the parent array of the OffsetArray, column-major,
 *    i fastest.  A "padded" 2-D array has (nx+2Hx) x (ny+2Hy) elements, Julia index (i, j) at
 *    element (i+Hx-1) + (j+Hy-1)*(nx+2Hx).  A 3-D field parent has (nx+2Hx)(ny+2Hy)(nz+2Hz).
 *  - Every function returns 0 on success, a negative OSSG_E* code for an argument error
 *    (the Julia shim raises ArgumentError), or a positive cudaError_t.  ossg_last_error()
 *    returns a thread-local message.  Nothing aborts, exits or throws across the ABI.
 *  - There is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef OSSG_B200_H
#define OSSG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OSSG_VERSION_MAJOR 0
#define OSSG_VERSION_MINOR 1

/* argument errors */
#define OSSG_EINVAL      (-1)   /* null pointer / nonsensical size                               */
#define OSSG_EODD_NX     (-2)   /* "The number of cells in the longitude dimension should be even!"
                                   src/tripolar_grid.jl:81-83                                    */
#define OSSG_EHALO       (-3)   /* halo larger than the fold can serve (needs Hy <= Ny-1, Hx <= Nx) */
#define OSSG_EWINDOW     (-4)   /* tile window outside the global padded extent                  */
#define OSSG_EALIGN      (-5)   /* device pointer not aligned to its element size                */
#define OSSG_ELOCATION   (-6)   /* location code not 0 (Center) / 1 (Face)                       */
#define OSSG_EWORKSPACE  (-7)   /* workspace too small                                           */
#define OSSG_ENCCL       (-8)   /* libnccl.so.2 missing, or an NCCL call failed (message has the ncclResult_t) */
#define OSSG_ETIMEOUT    (-9)   /* a peer-memory exchange of this rank gave up waiting for a partner */

/* staggering codes: Oceananigans Center / Face */
#define OSSG_CENTER 0
#define OSSG_FACE   1

/* element types */
#define OSSG_F64 0
#define OSSG_F32 1

/* index of each array in the `out[20]` argument = positional order of the
 * OrthogonalSphericalShellGrid constructor call at src/tripolar_grid.jl:308-328 (z omitted) */
enum {
    OSSG_LAM_CC = 0, OSSG_LAM_FC, OSSG_LAM_CF, OSSG_LAM_FF,
    OSSG_PHI_CC, OSSG_PHI_FC, OSSG_PHI_CF, OSSG_PHI_FF,
    OSSG_DX_CC, OSSG_DX_FC, OSSG_DX_CF, OSSG_DX_FF,
    OSSG_DY_CC, OSSG_DY_CF, OSSG_DY_FC, OSSG_DY_FF,      /* sic: cc, cf, fc, ff (:321-324) */
    OSSG_AZ_CC, OSSG_AZ_FC, OSSG_AZ_CF, OSSG_AZ_FF,
    OSSG_N_ARRAYS
};

/* Keyword arguments of TripolarGrid(arch, FT; size, southernmost_latitude, halo, radius, z,
 * north_poles_latitude, first_pole_longitude)  (src/tripolar_grid.jl:59-66) plus the tile window
 * of a Distributed grid (src/distributed_tripolar_grid.jl:41-49, 112-120). */
typedef struct ossg_tripolar_params {
    int32_t Nx, Ny;                 /* global size[1], size[2]                                   */
    int32_t Hx, Hy;                 /* halo[1], halo[2]                                          */
    double  southernmost_latitude;  /* default -80                                               */
    double  north_poles_latitude;   /* default  55                                               */
    double  first_pole_longitude;   /* default  70                                               */
    double  radius;                 /* default R_Earth = 6371.0e3                                */
    /* Tile: interior cells i0 .. i0+nx-1, j0 .. j0+ny-1 (1-based, global); the arrays written
     * are the window [i0-Hx, i0+nx-1+Hx] x [j0-Hy, j0+ny-1+Hy] of the GLOBAL padded arrays, i.e.
     * exactly partition_tripolar_metric's slice.  Serial grid: i0 = j0 = 1, nx = Nx, ny = Ny.
     * Setting nx = 0 (or ny = 0) means "whole grid" for that direction.                          */
    int32_t i0, j0, nx, ny;
    /* OSSG_LAMBDA1D_F32: the 1-D longitudes are Float32 values (rounded once from the correctly rounded
     * Float64 ones) widened back to Float64 before the map -- one reading of
     * generate_coordinate(FT, ...) at src/tripolar_grid.jl:89 for FT = Float32; 0 (default): they stay
     * Float64 for every FT.  See DESIGN.md section 4.                                             */
    int32_t flags;
    int32_t reserved;
} ossg_tripolar_params;
#define OSSG_LAMBDA1D_F32 1

/* ---- housekeeping ------------------------------------------------------------------------- */
int         ossg_version(void);                 /* major*100 + minor                             */
const char* ossg_last_error(void);              /* thread-local, never NULL                      */
int         ossg_device_ok(void);               /* 0 iff a CUDA device of compute capability 10.x
                                                   is current; else a positive cudaError_t / -1  */

/* ---- TripolarGrid constructor ------------------------------------------------------------- *
 * Replaces the body of TripolarGrid(arch, FT; ...) between src/tripolar_grid.jl:95 and :300
 * (rows C0-C3, M1-M3 of SURVEY.md section 8a) and of TripolarGrid(arch::Distributed, ...)
 * src/distributed_tripolar_grid.jl:36-73.
 *
 * workspace: device scratch of at least ossg_tripolar_workspace_bytes() bytes, 16-byte aligned
 * (1-D tables; plus the 8 FP64 coordinate arrays when out_type = OSSG_F32, because the metrics
 * are always computed from FP64 coordinates and rounded once at the end,
 * src/tripolar_grid.jl:308-328; plus one 4-byte counter per window row for the optional row-band
 * dependency between the coordinate and the metric kernel, OSSG_ROW_DEPS=1).  Every launch of a
 * call goes to the caller's stream; the kernels of one build are chained with programmatic
 * dependent launch.                                                                             */
size_t ossg_tripolar_workspace_bytes(const ossg_tripolar_params* p, int out_type);

/* The workspace starts with a 128-byte header and the 1-D tables (ossg_tripolar_table_bytes() bytes in
 * all).  The header holds validity stamps: a later build / coords / metrics call with the same
 * workspace and the same (Nx, Ny, southernmost_latitude, north_poles_latitude) finds the x/y tables
 * in place and does not recompute them (likewise the lat-lon rows for the same Hy and radius) --
 * the fast path of with_halo and reconstruct_global_grid (src/with_halo.jl:5-44,
 * src/distributed_tripolar_grid.jl:201-226), which re-run the constructor with other halos.
 * Contract: call ossg_tripolar_workspace_init once after allocating a workspace (it zeroes the header,
 * asynchronously on `stream`); do not write into the first ossg_tripolar_table_bytes() bytes between
 * calls; one workspace serves one stream at a time.  To carry the tables over to a larger workspace
 * (new halo), copy its first ossg_tripolar_table_bytes() bytes device-to-device.                  */
int    ossg_tripolar_workspace_init(void* workspace, size_t workspace_bytes, void* stream);
size_t ossg_tripolar_table_bytes(const ossg_tripolar_params* p);

/* all 20 arrays; out[k] (k = OSSG_LAM_CC ...) has (nx+2Hx)(ny+2Hy) elements of out_type */
int ossg_tripolar_build(const ossg_tripolar_params* p, int out_type, void* const out[20],
                        void* workspace, size_t workspace_bytes, void* stream);

/* same launches with CUDA events between the kernels; synchronises the stream and returns the
 * device time of tables / coordinates / metrics (tile) / metrics (halo frame) in ms[0..3].
 * Measurement aid for bench.py's roofline; not meant for production callers.                    */
int ossg_tripolar_build_profiled(const ossg_tripolar_params* p, int out_type, void* const out[20],
                                 void* workspace, size_t workspace_bytes, void* stream, float ms[4]);

/* stage-wise entry points (same kernels; used for staged parity tests and by with_halo-style
 * callers that already hold coordinates).  coords[8]: lam cc,fc,cf,ff, phi cc,fc,cf,ff, always
 * FP64.  metrics[12]: dx cc,fc,cf,ff, dy cc,cf,fc,ff, az cc,fc,cf,ff of metrics_type.
 * ossg_tripolar_metrics reads the coordinate arrays it is given: with a whole-grid window they
 * may come from any source (e.g. the CPU oracle); with a tile window the halo ring, whose
 * stencil leaves the window, is evaluated from the closed form instead.                          */
int ossg_tripolar_coords(const ossg_tripolar_params* p, void* const coords[8],
                         void* workspace, size_t workspace_bytes, void* stream);
int ossg_tripolar_metrics(const ossg_tripolar_params* p, const void* const coords[8],
                          int metrics_type, void* const metrics[12],
                          void* workspace, size_t workspace_bytes, void* stream);

/* ---- Cartesian nodes --------------------------------------------------------------------- *
 * (external) Oceananigans.Grids.lat_lon_to_cartesian(lat, lon, radius) over n nodes, the kernel of
 * get_cartesian_nodes_and_vertices as used by test/test_tripolar_grid.jl:68 and
 * examples/visualize_tripolar_grid.jl:39-42:
 *     x = radius cosd(lon) cosd(lat),  y = radius sind(lon) cosd(lat),  z = radius sind(lat)
 * lam/phi are FP64 coordinate arrays (any staggering, any layout: n contiguous elements);
 * x, y, z are arrays of out_type.                                                               */
int ossg_cartesian_nodes(const void* lam, const void* phi, size_t n, double radius, int out_type,
                         void* x, void* y, void* z, void* stream);

/* ---- Zipper boundary condition ------------------------------------------------------------- *
 * Replaces the per-thread bodies fold_north_{center_center,face_center,center_face,face_face}!
 * (src/zipper_boundary_condition.jl:70-138) as dispatched by _fill_north_halo! (:146-155), for
 * all i in 1..Nx and the parent k-levels [k0, k0+nk):
 *     c[i, Ny+j, k] = s * c[i', Ny-j(+1), k],  j = 1..Hy ;  y-Center: c[i, Ny, k] = s*c[i', Ny, k] for i > Nx/2
 * `sign` is bc.condition (ZipperBoundaryCondition(sign), :52).                                  */
typedef struct ossg_field_desc {
    void*   data;                   /* parent array, device                                      */
    int32_t elem_type;              /* OSSG_F64 / OSSG_F32                                       */
    int32_t Nx, Ny, Nz;             /* interior size                                             */
    int32_t Hx, Hy, Hz;             /* halo                                                      */
    int32_t locx, locy;             /* OSSG_CENTER / OSSG_FACE                                   */
    int32_t k0, nk;                 /* parent k-levels to fill; nk = 0 means k0 = Hz, nk = Nz    */
    double  sign;                   /* bc.condition                                              */
} ossg_field_desc;

int ossg_zipper_fill(const ossg_field_desc* f, void* stream);
/* several fields in ONE launch (tupled_fill_halo_regions!, src/tripolar_grid_extensions.jl:17) */
int ossg_zipper_fill_batch(const ossg_field_desc* f, int nfields, void* stream);
/* zipper north fill followed by the periodic west/east fill over the whole (y,z) parent extent --
 * fill_halo_regions! of a field on a TripolarGrid whose other boundaries are `nothing`
 * (src/tripolar_grid.jl:147-186); one launch per batch.                                          */
int ossg_fill_halo_regions_batch(const ossg_field_desc* f, int nfields, void* stream);

/* ---- (x, y)-partitioned fold: pack / unpack around the mirror-rank exchange ----------------- *
 * New relative to the reference (which only supports Y slabs, src/distributed_tripolar_grid.jl:
 * 28-31).  A top-row rank owning global columns [i0, i0+nx-1] needs, for its north halo, the
 * source rows of global columns [Nx-(i0+nx-1)+1(+1) .. ] held by its mirror rank.
 * ossg_fold_pack gathers, from the LOCAL field, the rows another rank needs into a contiguous
 * buffer; ossg_fold_unpack writes a received buffer into the local north halo with the i-reversal
 * and sign.  The exchange itself (ncclSend/ncclRecv or peer copy) is the caller's.
 * Buffer layout: [k][row r = 0..nrows-1][global column ascending], nrows = Hy (+1 for y-Center:
 * the row-Ny overwrite), element type of the field.                                             */
typedef struct ossg_fold_window {
    int32_t gNx;                    /* global Nx                                                 */
    int32_t i0, nx;                 /* this rank's global first interior column and width        */
    int32_t src_i0, src_n;          /* global column range [src_i0, src_i0+src_n) to pack/unpack */
} ossg_fold_window;

size_t ossg_fold_buffer_elems(const ossg_field_desc* f, const ossg_fold_window* w);
int ossg_fold_pack(const ossg_field_desc* f, const ossg_fold_window* w, void* buffer, void* stream);
int ossg_fold_unpack(const ossg_field_desc* f, const ossg_fold_window* w, const void* buffer, void* stream);

/* ---- (x, y)-partitioned fold and ordinary halos through NCCL send/recv ----------------------- *
 * north_star: "ordinary halos and the fold-partner exchange between mirror ranks in the top row go
 * through NCCL send/recv over NVLink, replacing MPI".  Replaces, for a field whose north boundary is
 * a ZipperBoundaryCondition on a Distributed grid, the communication that Oceananigans' distributed
 * fill_halo_regions! performs with MPI (fold: src/zipper_boundary_condition.jl:70-155 applied across
 * ranks; partition: src/distributed_tripolar_grid.jl:41-49, generalised from Y slabs to (x, y)).
 *
 * `nccl_comm` is the caller's ncclComm_t spanning all Rx*Ry ranks, rank r at (rx, ry) = (r / Ry,
 * r % Ry) (Oceananigans' rank2index), tiles by the reference's rule (equal parts, the last rank takes
 * the remainder).  `fields` are the LOCAL fields (Nx, Ny = this rank's tile), at most
 * OSSG_MAX_PEER_BATCH per call.  One call enqueues on `stream`:
 *   ossg_fold_exchange:  one pack kernel, one ncclGroup with one ncclSend + ncclRecv per partner
 *                        rank (all fields and levels of the batch in one message), one unpack kernel
 *                        (i-reversal and sign applied on the way out).  Ranks below the top row return 0.
 *   ossg_halo_exchange:  south/north rows (whole rows), then west/east columns over the whole y extent,
 *                        periodic in x (a local copy when Rx == 1): pack, group, unpack per phase.
 * The workspace is caller-provided, persistent device memory of at least the *_workspace_bytes()
 * size (0 when the rank has nothing to exchange); the library allocates nothing and never
 * synchronises.  NCCL is resolved at run time from the libnccl.so.2 already loaded in the process.
 * Every rank must make the matching call with the same field list.                              */
typedef struct ossg_exchange_layout {
    int32_t gNx, gNy;               /* global size[1], size[2]                                   */
    int32_t Rx, Ry;                 /* Partition(x, y)                                           */
    int32_t rank;                   /* this rank in nccl_comm                                    */
    int32_t reserved;
} ossg_exchange_layout;

size_t ossg_fold_exchange_workspace_bytes(const ossg_field_desc* fields, int nfields, const ossg_exchange_layout* layout);
int    ossg_fold_exchange(void* nccl_comm, const ossg_field_desc* fields, int nfields, const ossg_exchange_layout* layout,
                          void* workspace, size_t workspace_bytes, void* stream);
size_t ossg_halo_exchange_workspace_bytes(const ossg_field_desc* fields, int nfields, const ossg_exchange_layout* layout);
int    ossg_halo_exchange(void* nccl_comm, const ossg_field_desc* fields, int nfields, const ossg_exchange_layout* layout,
                          void* workspace, size_t workspace_bytes, void* stream);

/* the messages of one fold exchange, identical on every rank (host only, no device needed): entry q
 * says that rank src_rank sends global source columns [src_i0, src_i0+n) of field `field`, `rows`
 * rows per level, to rank dst_rank.  Returns the number of messages (fills at most `capacity`).  */
typedef struct ossg_fold_segment {
    int32_t field, src_rank, dst_rank, src_i0, n, rows, wrapped, reserved;
} ossg_fold_segment;
int ossg_fold_exchange_plan(const ossg_field_desc* fields, int nfields, const ossg_exchange_layout* layout,
                            ossg_fold_segment* out, int capacity);

/* communicator helpers for hosts without their own NCCL binding (thin wrappers of ncclGetUniqueId /
 * ncclCommInitRank / ncclCommDestroy); a Julia host passes NCCL.jl's communicator handle instead */
int ossg_nccl_get_unique_id(unsigned char id[128]);
int ossg_nccl_comm_init_rank(void** comm, int nranks, const unsigned char id[128], int rank);
int ossg_nccl_comm_destroy(void* comm);

/* ---- (x, y)-partitioned fold over NVLink peer memory: ONE kernel per rank, no staging ------- *
 * B200-native form of the exchange above for ranks of one NVSwitch node.  Every top-row rank maps
 * the field arrays of the other top-row ranks (CUDA IPC, ossg_ipc_export / ossg_ipc_open) and its
 * zipper kernel loads the source rows straight from the owner (the mirror rank; for x-Face fields
 * also one column of its neighbour) while writing its own north halo -- each element crosses
 * NVLink once.  Ranks order themselves with release/acquire epoch flags in IPC-mapped device memory (an
 * "arrived" flag at kernel start, a "finished" flag written by the last block at kernel end, which then
 * waits for the partners' flags); the host never synchronises and no NCCL call is on this path.  A rank
 * that never makes the matching call is a programming error: the waiting kernels trap after
 * OSSG_PEER_TIMEOUT_MS and the next CUDA call of the process fails.  Same result as ossg_zipper_fill on the global
 * field (src/zipper_boundary_condition.jl:70-155), bit for bit.                                 */
#define OSSG_MAX_PEERS 8
#define OSSG_MAX_PEER_BATCH 16
#define OSSG_IPC_HANDLE_BYTES 64
typedef struct ossg_ipc_handle { unsigned char bytes[OSSG_IPC_HANDLE_BYTES]; } ossg_ipc_handle;
/* handle of the device allocation that contains device_ptr, and device_ptr's offset inside it   */
int ossg_ipc_export(const void* device_ptr, ossg_ipc_handle* handle, uint64_t* offset);
/* maps a peer's allocation into this process (once per handle; cached) and returns base+offset  */
int ossg_ipc_open(const ossg_ipc_handle* handle, uint64_t offset, void** peer_ptr);
/* one ossg_ipc_close per successful ossg_ipc_open; the mapping goes away with the last one (the
 * caller makes sure no kernel that uses the mapped pointers is still in flight)                  */
int ossg_ipc_close(const ossg_ipc_handle* handle);

typedef struct ossg_peer_fold {
    int32_t gNx;                            /* global Nx                                         */
    int32_t me, nranks;                     /* this rank's x-index in the top row, Rx (<= 8)     */
    int32_t i0[OSSG_MAX_PEERS];             /* global first interior column of top-row rank r    */
    int32_t nx[OSSG_MAX_PEERS];             /* its width                                         */
    void*   flags_local;                    /* this rank's flag block: ready[OSSG_MAX_PEERS] and
                                               done[OSSG_MAX_PEERS] epochs (uint64), then one uint64: lower
                                               half a block counter, upper half the error word (set to 1 when
                                               a partner did not show up within OSSG_PEER_TIMEOUT_MS,
                                               default 120 s; the kernel then TRAPS): 2*OSSG_MAX_PEERS+1
                                               uint64, zero-initialised once, never reset            */
    void*   flags_peer[OSSG_MAX_PEERS];     /* the peers' flag blocks, peer-mapped ([me] unused)  */
    uint64_t epoch;                         /* 1, 2, 3, ... growing with every call on this flag block, the same on all ranks */
    int32_t push;                           /* 0: each rank loads its sources from the owners;
                                               1: each rank stores its rows into the destinations'
                                               halos (posted NVLink writes);
                                               2: as 1, rows staged in shared memory in destination
                                               order and sent as bulk copies (cp.async.bulk, one
                                               transfer per row segment); same on all ranks        */
    int32_t bulk_waves;                     /* push == 2: work items per resident block the kernel aims for when
                                               it narrows its chunks; 0 = library default (3, or
                                               OSSG_PEER_BULK_WAVES), < 0 = always the widest chunks  */
} ossg_peer_fold;

typedef struct ossg_peer_field {
    ossg_field_desc f;                      /* the LOCAL field (Nx = this rank's width)          */
    const void* peer[OSSG_MAX_PEERS];       /* data pointer of the same field on top-row rank r
                                               (peer-mapped; the local pointer for r == me)      */
} ossg_peer_field;

/* north-fold fill of up to OSSG_MAX_PEER_BATCH fields of a top-row rank; every top-row rank must
 * make the matching call (same fields, same epoch) or the kernels wait for each other forever   */
int ossg_zipper_fill_peer(const ossg_peer_field* fields, int nfields, const ossg_peer_fold* fold, void* stream);

/* ---- ordinary halos of an (x, y)-partitioned field over NVLink peer memory ------------------ *
 * Replaces Oceananigans' MPI halo passing (Isend/Irecv of FieldBoundaryBuffers between j- and
 * i-adjacent ranks) for ranks of one NVSwitch node, with the same two ordered steps: phase 0
 * stores this rank's top / bottom Hy interior rows (whole rows, all parent levels) into the south
 * halo of the north neighbour and the north halo of the south neighbour; phase 1 stores its east /
 * west Hx interior columns over the WHOLE local y extent (so corners carry exchanged and folded
 * rows) into the west halo of the east neighbour and the east halo of the west neighbour, periodic
 * in x.  One kernel per phase, posted NVLink writes, the flag barriers of ossg_zipper_fill_peer
 * between neighbours only.  A neighbour id of -1 means "no neighbour" (domain edge in y); a
 * neighbour equal to `me` (Rx == 1) makes the copy local.                                       */
typedef struct ossg_peer_halo {
    int32_t me, nranks;                     /* rank id and world size (<= OSSG_MAX_PEERS)        */
    int32_t north, south, east, west;       /* neighbour rank ids                                */
    int32_t phase;                          /* 0: south/north rows, 1: west/east columns,
                                               2: unpack of the staged west/east columns (local)  */
    int32_t reserved;
    void*   flags_local;                    /* flag block of THIS phase (2*OSSG_MAX_PEERS+1 uint64)              */
    void*   flags_peer[OSSG_MAX_PEERS];     /* the same block on rank r, peer-mapped              */
    uint64_t epoch;
    /* Optional staging of phase 1: a halo column is a 32-byte piece per row, the least efficient
     * NVLink packet.  With staging, phase 1 stores the pieces CONTIGUOUSLY ([field][row][Hx]) into
     * a buffer on the neighbour (full-size packets) and phase 2, a local kernel, scatters the
     * received buffers into the halo columns.  NULL pointers select the direct form.            */
    void*   stage_to_east;                  /* east neighbour's "from west" buffer, peer-mapped  */
    void*   stage_to_west;                  /* west neighbour's "from east" buffer, peer-mapped  */
    void*   stage_from_west;                /* this rank's buffers (phase 2 reads them)          */
    void*   stage_from_east;
} ossg_peer_halo;

typedef struct ossg_peer_halo_field {
    ossg_field_desc f;                      /* the LOCAL field                                   */
    void*   north; void* south; void* east; void* west;   /* the field's data on the neighbours  */
    int32_t ny_north, ny_south;             /* interior rows of the north / south neighbour      */
    int32_t nx_east, nx_west;               /* interior columns of the east / west neighbour     */
} ossg_peer_halo_field;

int ossg_halo_exchange_peer(const ossg_peer_halo_field* fields, int nfields, const ossg_peer_halo* halo, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OSSG_B200_H */
