"""Distributed tripolar grid and halo exchange: one process per GPU, torch.distributed (NCCL over
NVLink on the GPU box, gloo in CPU tests) for the plumbing.

Mirrors:
  TripolarGrid(arch::Distributed, FT; halo, kwargs...)      src/distributed_tripolar_grid.jl:24-110
  partition_tripolar_metric                                 src/distributed_tripolar_grid.jl:112-120
  reconstruct_global_grid                                   src/distributed_tripolar_grid.jl:201-226
The reference supports Y slabs only (:28-31) and builds the whole global grid on every rank (:36)
before slicing.  Here every rank evaluates only its own window of the global padded arrays
(ossg_tripolar_build with a tile window; zero communication), and the partition may also split x:
the fold then crosses ranks and is served by a pack -> NCCL send/recv -> unpack exchange between
mirror ranks of the top row (new relative to the reference; its correctness spec is "every
rank's tile equals the same slice of the serial result").
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import ARRAY_NAMES, FoldWindow
from .grids import (GPU, Float64, OffsetArray, OrthogonalSphericalShellGrid, R_Earth, Tripolar,
                    _generate_z, _require_cuda, build_arrays, build_arrays_cached, make_params)


@dataclass(frozen=True)
class Partition:
    """Oceananigans.DistributedComputations.Partition(x, y): number of ranks per direction."""
    x: int = 1
    y: int = 1


def local_sizes(N: int, R: int) -> list:
    """Equal partition, the last rank takes the remainder (external: Oceananigans local_size;
    src/distributed_tripolar_grid.jl:41-49)."""
    base = N // R
    sizes = [base] * R
    sizes[-1] += N - base * R
    return sizes


class Distributed:
    """Oceananigans.Distributed(child_architecture; partition).  Rank r sits at
    (rx, ry) = (r // Ry, r % Ry) (Oceananigans rank2index: x slowest)."""

    def __init__(self, child_architecture=None, partition: Optional[Partition] = None,
                 rank: Optional[int] = None, world_size: Optional[int] = None, group=None,
                 y_slabs_only: bool = False):
        import torch.distributed as dist
        self.child_architecture = child_architecture if child_architecture is not None else GPU()
        self.group = group
        if rank is None:
            rank = dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0
        if world_size is None:
            world_size = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.local_rank = int(rank)
        self.world_size = int(world_size)
        self.partition = partition if partition is not None else Partition(1, self.world_size)
        if self.partition.x * self.partition.y != self.world_size:
            raise ValueError(f"partition {self.partition} does not match {self.world_size} ranks")
        if y_slabs_only and self.partition.x != 1:     # src/distributed_tripolar_grid.jl:30-31
            raise ValueError("The tripolar grid is supported only on a Y-partitioning configuration")
        self.Rx, self.Ry = self.partition.x, self.partition.y
        self.rx, self.ry = self.local_rank // self.Ry, self.local_rank % self.Ry

    def rank_of(self, rx: int, ry: int) -> int:
        return rx * self.Ry + ry

    def is_top_row(self) -> bool:
        return self.ry == self.Ry - 1

    def torch_device(self):
        return self.child_architecture.torch_device()

    def __repr__(self):
        return f"Distributed({self.child_architecture}, {self.partition}, rank {self.local_rank}/{self.world_size})"


def tile_of(Nx: int, Ny: int, Rx: int, Ry: int, rx: int, ry: int):
    """(i0, j0, nx, ny) of rank (rx, ry): jstart = 1 + sum(nlocal[1:rank]) (:47)."""
    sx, sy = local_sizes(Nx, Rx), local_sizes(Ny, Ry)
    return 1 + sum(sx[:rx]), 1 + sum(sy[:ry]), sx[rx], sy[ry]


def DistributedTripolarGrid(arch: Distributed, FT=Float64, *, size, southernmost_latitude=-80, halo=(4, 4, 4),
                            radius=R_Earth, z=(0, 1), north_poles_latitude=55, first_pole_longitude=70, _tables=None,
                            float32_longitudes: bool = False):
    Nx, Ny, Nz = (int(s) for s in size)
    Hx, Hy, Hz = (int(h) for h in halo)
    if Nx % 2 == 1:
        raise ValueError("The number of cells in the longitude dimension should be even!")
    i0, j0, nx, ny = tile_of(Nx, Ny, arch.Rx, arch.Ry, arch.rx, arch.ry)
    if ny - 1 < Hy and arch.is_top_row():
        raise ValueError(f"top-row tile has {ny} rows: the fold needs at least Hy+1 = {Hy + 1}")
    if (arch.Ry > 1 and ny < Hy) or (arch.Rx > 1 and nx < Hx):      # a neighbour's halo is filled from INTERIOR rows / columns
        raise ValueError(f"tile {nx}x{ny} of rank {arch.local_rank} is smaller than the halo ({Hx}, {Hy})")
    params = make_params(size, halo, southernmost_latitude, north_poles_latitude, first_pole_longitude, radius,
                         tile=(i0, j0, nx, ny), flags=_lib.OSSG_LAMBDA1D_F32 if float32_longitudes else 0)
    _require_cuda()
    device = arch.torch_device()
    parents, tables = build_arrays_cached(params, FT, device, tables=_tables)
    zc = _generate_z(z, Nz)
    LX = "Periodic" if arch.Rx == 1 else "FullyConnected"
    LY = "RightConnected" if arch.ry == 0 else "FullyConnected"     # :75
    return OrthogonalSphericalShellGrid(
        architecture=arch, Nx=nx, Ny=ny, Nz=Nz, Hx=Hx, Hy=Hy, Hz=Hz, Lz=zc.Lz, z=zc, radius=float(radius),
        conformal_mapping=Tripolar(north_poles_latitude, first_pole_longitude, southernmost_latitude),
        dtype=parents[0].dtype, topology=(LX, LY, "Bounded"),
        arrays={n: OffsetArray(t, Hx, Hy) for n, t in zip(ARRAY_NAMES, parents)},
        i0=i0, j0=j0, global_size=(Nx, Ny, Nz), tables=tables)


def reconstruct_global_grid(grid: OrthogonalSphericalShellGrid):
    """src/distributed_tripolar_grid.jl:201-226: the global grid on the child architecture (note:
    like the reference, `radius` is not forwarded)."""
    from .grids import TripolarGrid
    arch = grid.architecture
    cm = grid.conformal_mapping
    return TripolarGrid(arch.child_architecture, grid.dtype, halo=grid.halo, size=grid.global_size,
                        north_poles_latitude=cm.north_poles_latitude,
                        first_pole_longitude=cm.first_pole_longitude,
                        southernmost_latitude=cm.southernmost_latitude, z=grid.z.zf,
                        _tables=getattr(grid, "tables", None))      # a tile's 1-D tables are the global ones (same Nx, Ny)


# ---------------------------------------------------------------- fold communication plan
@dataclass(frozen=True)
class FoldMessage:
    """Rank `src` packs global columns [src_i0, src_i0 + n) of its fold source rows and sends
    them to rank `dst` (top row), which unpacks them into its north halo."""
    src: int
    dst: int
    src_i0: int
    n: int


def fold_plan(Nx: int, Rx: int, Ry: int, locx: int) -> list:
    """All messages of one fold exchange for fields at x-location `locx`, identical on every rank.
    Destination column i needs source column i' = Nx - i + 1 (+1 for x-Face, wrapping Nx+1 -> 1)
    (src/zipper_boundary_condition.jl:73-75,90-92,110,125)."""
    sx = local_sizes(Nx, Rx)
    starts = [1 + sum(sx[:r]) for r in range(Rx)]
    top = Ry - 1
    msgs = []
    for drx in range(Rx):
        lo, hi = starts[drx], starts[drx] + sx[drx] - 1          # destination columns
        a, b = Nx - hi + 1 + locx, Nx - lo + 1 + locx            # source columns, ascending
        ranges = [(a, min(b, Nx))] if a <= Nx else []
        if b > Nx:
            ranges.append((1, b - Nx))                           # periodic wrap of the x-Face fold
        for (ra, rb) in ranges:
            for srx in range(Rx):
                slo, shi = starts[srx], starts[srx] + sx[srx] - 1
                c0, c1 = max(ra, slo), min(rb, shi)
                if c0 <= c1:
                    msgs.append(FoldMessage(src=srx * Ry + top, dst=drx * Ry + top, src_i0=c0, n=c1 - c0 + 1))
    return msgs


class CudaPacker:
    """pack / unpack through libossg_b200.so (ossg_fold_pack / ossg_fold_unpack)."""

    def empty(self, f, n_elems: int) -> torch.Tensor:
        return torch.empty(n_elems, dtype=f.data.dtype, device=f.data.device)

    def buffer_elems(self, f, window: FoldWindow) -> int:
        import ctypes as C
        d = f.descriptor()
        return int(_lib.lib().ossg_fold_buffer_elems(C.byref(d), C.byref(window)))

    def pack(self, f, window: FoldWindow, buf: torch.Tensor, stream) -> None:
        import ctypes as C
        d = f.descriptor()
        _lib.check(_lib.lib().ossg_fold_pack(C.byref(d), C.byref(window), buf.data_ptr(), stream))

    def unpack(self, f, window: FoldWindow, buf: torch.Tensor, stream) -> None:
        import ctypes as C
        d = f.descriptor()
        _lib.check(_lib.lib().ossg_fold_unpack(C.byref(d), C.byref(window), buf.data_ptr(), stream))

    def stream(self, f):
        return torch.cuda.current_stream(f.data.device).cuda_stream

    def copy_rows(self, dst, src):
        dst.copy_(src)


def _window(f, gNx: int, src_i0: int, n: int) -> FoldWindow:
    return FoldWindow(gNx, f.grid.i0, f.grid.Nx, src_i0, n)


def fold_exchange(fields: Sequence, packer=None, group=None) -> None:
    """Fold-partner exchange for a batch of fields on an (x, y)-partitioned tripolar grid: every
    top-row rank packs the source rows each destination needs, all messages of the batch to one
    partner travel as ONE send (fields and k-levels aggregated), then each rank unpacks with the
    i-reversal and sign.  Ranks below the top row return immediately."""
    import torch.distributed as dist
    from .fields import is_zipper
    fields = [f for f in fields if f.boundary_conditions is not None and is_zipper(f.boundary_conditions.north)]
    if not fields:
        return
    arch = fields[0].grid.architecture
    if not arch.is_top_row():
        return
    packer = packer or CudaPacker()
    gNx = fields[0].grid.global_size[0]
    me = arch.local_rank
    # per-partner segment lists: (field, window, n_elems)
    send_segs, recv_segs, local_segs = {}, {}, []
    for f in fields:
        locx = _lib.OSSG_FACE if f.descriptor().locx == _lib.OSSG_FACE else _lib.OSSG_CENTER
        for msg in fold_plan(gNx, arch.Rx, arch.Ry, locx):
            if msg.src != me and msg.dst != me:
                continue
            w = _window(f, gNx, msg.src_i0, msg.n)
            ne = packer.buffer_elems(f, w)
            if msg.src == me and msg.dst == me:
                local_segs.append((f, w, ne))
            elif msg.src == me:
                send_segs.setdefault(msg.dst, []).append((f, w, ne))
            else:
                recv_segs.setdefault(msg.src, []).append((f, w, ne))
    stream = packer.stream(fields[0])
    # pack: one buffer per partner (all fields share a dtype per buffer; split by dtype if mixed)
    ops, send_bufs, recv_bufs = [], {}, {}
    for peer, segs in sorted(send_segs.items()):
        for dt in sorted({s[0].data.dtype for s in segs}, key=str):
            sel = [s for s in segs if s[0].data.dtype == dt]
            buf = packer.empty(sel[0][0], sum(s[2] for s in sel))
            off = 0
            for (f, w, ne) in sel:
                packer.pack(f, w, buf[off:off + ne], stream)
                off += ne
            send_bufs[(peer, dt)] = buf
            ops.append(dist.P2POp(dist.isend, buf, peer, group))
    for peer, segs in sorted(recv_segs.items()):
        for dt in sorted({s[0].data.dtype for s in segs}, key=str):
            sel = [s for s in segs if s[0].data.dtype == dt]
            buf = packer.empty(sel[0][0], sum(s[2] for s in sel))
            recv_bufs[(peer, dt)] = (buf, sel)
            ops.append(dist.P2POp(dist.irecv, buf, peer, group))
    # self part (Rx == 1, or the middle rank of an odd top row, or the wrapped x-Face column)
    for (f, w, ne) in local_segs:
        buf = packer.empty(f, ne)
        packer.pack(f, w, buf, stream)
        packer.unpack(f, w, buf, stream)
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for (peer, dt), (buf, sel) in recv_bufs.items():
        off = 0
        for (f, w, ne) in sel:
            packer.unpack(f, w, buf[off:off + ne], stream)
            off += ne


class NcclExchange:
    """fill_halo_regions! of fields on an (x, y)-partitioned tripolar grid through NCCL send/recv in
    libossg_b200.so (ossg_fold_exchange + ossg_halo_exchange): per phase ONE pack kernel, ONE
    ncclGroup with one message per partner, ONE unpack kernel, on the caller's stream, with a
    persistent buffer -- no per-call allocation and no host wait.  This is the north-star's NCCL
    arm; PeerFold / PeerHalo are the peer-memory forms measured against it.

    Collective constructor over `group` (all ranks, same field list).  torch.distributed does not
    expose its ncclComm_t, so the constructor creates one through ossg_nccl_comm_init_rank (the
    unique id travels over `group`); a Julia host passes NCCL.jl's communicator instead."""

    _comms: dict = {}

    def __init__(self, fields: Sequence, group=None):
        import ctypes as C
        import torch.distributed as dist
        self.fields = list(fields)
        arch = self.fields[0].grid.architecture
        self.arch = arch
        L = _lib.lib()
        dev = self.fields[0].data.device
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        key = (id(group), dev.index, world)
        if key not in NcclExchange._comms:
            uid = C.create_string_buffer(128)
            if rank == 0:
                _lib.check(L.ossg_nccl_get_unique_id(uid))
            box = [bytes(uid.raw)]
            dist.broadcast_object_list(box, src=0, group=group)
            comm = C.c_void_p()
            with torch.cuda.device(dev):
                _lib.check(L.ossg_nccl_comm_init_rank(C.byref(comm), world, box[0], rank))
            NcclExchange._comms[key] = comm
        self.comm = NcclExchange._comms[key]
        gN = self.fields[0].grid.global_size
        self.layout = _lib.ExchangeLayout(gN[0], gN[1], arch.Rx, arch.Ry, rank, 0)
        B = _lib.OSSG_MAX_PEER_BATCH
        self._batches = []
        need = 16
        for b0 in range(0, len(self.fields), B):
            chunk = self.fields[b0:b0 + B]
            arr = (_lib.FieldDesc * len(chunk))(*[f.descriptor() for f in chunk])
            self._batches.append((arr, len(chunk)))
            need = max(need, L.ossg_fold_exchange_workspace_bytes(arr, len(chunk), C.byref(self.layout)),
                       L.ossg_halo_exchange_workspace_bytes(arr, len(chunk), C.byref(self.layout)))
        self.workspace = torch.empty(need, dtype=torch.uint8, device=dev)

    def fold(self, stream=None) -> None:
        self._run(_lib.lib().ossg_fold_exchange, stream)

    def halos(self, stream=None) -> None:
        self._run(_lib.lib().ossg_halo_exchange, stream)

    def _run(self, fn, stream) -> None:
        import ctypes as C
        dev = self.fields[0].data.device
        stream = stream if stream is not None else torch.cuda.current_stream(dev).cuda_stream
        for (arr, n) in self._batches:
            rc = fn(self.comm, arr, n, C.byref(self.layout), self.workspace.data_ptr(), self.workspace.numel(), stream)
            if rc:
                _lib.check(rc)

    def __call__(self, stream=None) -> None:
        """fold on the top row, then south/north rows, then west/east columns (asynchronous)."""
        self.fold(stream)
        self.halos(stream)


class PeerFold:
    """The fold of an x-partitioned top row as ONE kernel per rank over NVLink peer memory
    (ossg_zipper_fill_peer): every top-row rank maps the other ranks' field arrays through CUDA IPC
    once, then each call launches a zipper kernel that loads its source rows from the owning rank
    and orders itself with the peers through epoch flags in device memory -- no pack/unpack, no
    staging buffer, no NCCL call and no host synchronisation per exchange.

    Collective constructor: every rank of `group` must build it for the same list of fields (ranks
    below the top row take part in the handle exchange only)."""

    def __init__(self, fields: Sequence, group=None, push=True):
        import ctypes as C
        import torch.distributed as dist
        from .fields import is_zipper
        # push: False / 0 = pull (peer loads), True / 1 = push (peer stores), 2 or "bulk" = push through
        # bulk copies (rows staged in shared memory, one cp.async.bulk transfer per row segment)
        self.push = 2 if push in (2, "bulk") else int(bool(push))
        self.fields = [f for f in fields if f.boundary_conditions is not None and is_zipper(f.boundary_conditions.north)]
        arch = fields[0].grid.architecture
        if arch.Rx > _lib.OSSG_MAX_PEERS:
            raise ValueError(f"PeerFold supports at most {_lib.OSSG_MAX_PEERS} ranks in x, got {arch.Rx}")
        self.arch, self.group, self.epoch = arch, group, 0
        self._launches = None
        self.active = arch.is_top_row() and bool(self.fields)
        L = _lib.lib()
        dev = fields[0].data.device
        # flag block: ready[8], done[8] (uint64) + block counter; zero once, epochs only grow
        self.flags = torch.zeros(2 * _lib.OSSG_MAX_PEERS + 1, dtype=torch.int64, device=dev)
        torch.cuda.synchronize(dev)

        def export(ptr):
            h, off = _lib.IpcHandle(), C.c_uint64()
            _lib.check(L.ossg_ipc_export(ptr, C.byref(h), C.byref(off)))
            return bytes(h.bytes), int(off.value)

        mine = None
        if self.active:
            mine = {"rx": arch.rx, "flags": export(self.flags.data_ptr()),
                    "fields": [export(f.data.data_ptr()) for f in self.fields]}
        world = dist.get_world_size(group)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine, group=group)
        if not self.active:
            return
        top = {g["rx"]: g for g in gathered if g is not None}
        assert sorted(top) == list(range(arch.Rx)), "every top-row rank must construct the same PeerFold"

        self._opened = []

        def open_(hb, off):
            h, out = _lib.IpcHandle(), C.c_void_p()
            C.memmove(h.bytes, hb, 64)
            _lib.check(L.ossg_ipc_open(C.byref(h), off, C.byref(out)))
            self._opened.append(hb)
            return out.value

        gNx = self.fields[0].grid.global_size[0]
        sx = local_sizes(gNx, arch.Rx)
        fd = _lib.PeerFoldDesc()
        fd.gNx, fd.me, fd.nranks, fd.push = gNx, arch.rx, arch.Rx, int(self.push)
        for r in range(arch.Rx):
            fd.i0[r], fd.nx[r] = 1 + sum(sx[:r]), sx[r]
            fd.flags_peer[r] = None if r == arch.rx else open_(*top[r]["flags"])
        fd.flags_local = self.flags.data_ptr()
        self.fold = fd
        self.peer_ptrs = []
        for k, f in enumerate(self.fields):
            self.peer_ptrs.append([f.data.data_ptr() if r == arch.rx else open_(*top[r]["fields"][k])
                                   for r in range(arch.Rx)])

    def __call__(self, stream=None) -> None:
        """One fold exchange of all fields (collective over the top row; asynchronous)."""
        import ctypes as C
        if not self.active:
            return
        L = _lib.lib()
        dev = self.fields[0].data.device
        stream = stream if stream is not None else torch.cuda.current_stream(dev).cuda_stream
        if self._launches is None:          # descriptors are built once: a call is one ctypes call per batch
            B = _lib.OSSG_MAX_PEER_BATCH
            self._launches = []
            for b0 in range(0, len(self.fields), B):
                chunk = list(range(b0, min(b0 + B, len(self.fields))))
                arr = (_lib.PeerField * len(chunk))()
                for n, k in enumerate(chunk):
                    arr[n].f = self.fields[k].descriptor()
                    for r in range(self.arch.Rx):
                        arr[n].peer[r] = self.peer_ptrs[k][r]
                self._launches.append((arr, len(chunk)))
        fold_ref = C.byref(self.fold)
        for (arr, n) in self._launches:
            self.epoch += 1
            self.fold.epoch = self.epoch
            rc = L.ossg_zipper_fill_peer(arr, n, fold_ref, stream)
            if rc:
                _lib.check(rc)

    def timed_out(self) -> bool:
        """True when a kernel of this rank gave up waiting for a peer (synchronises the device).  The
        kernel traps in that case, so normally the synchronisation itself raises first."""
        word = int(self.flags[2 * _lib.OSSG_MAX_PEERS].item())
        return (word >> 32) != 0

    def close(self) -> None:
        """Release the peer mappings (collective in spirit: call it on every rank once no exchange is
        in flight; the fields themselves stay alive as long as this object does)."""
        import ctypes as C
        if getattr(self, "_opened", None):
            torch.cuda.synchronize(self.fields[0].data.device)
            L = _lib.lib()
            for hb in self._opened:
                h = _lib.IpcHandle()
                C.memmove(h.bytes, hb, 64)
                L.ossg_ipc_close(C.byref(h))
            self._opened = []
        self.active = False


class PeerHalo:
    """fill_halo_regions! of fields on an (x, y)-partitioned tripolar grid entirely over NVLink peer
    memory: the fold (PeerFold, top row), then south/north rows, then west/east columns over the
    whole y extent (ossg_halo_exchange_peer, one kernel per phase) -- the order of
    distributed_fill_halo_regions(), without NCCL, staging buffers or host synchronisation.
    Collective constructor over `group` (all ranks, same field list, world size <= 8)."""

    def __init__(self, fields: Sequence, group=None):
        import ctypes as C
        import torch.distributed as dist
        arch = fields[0].grid.architecture
        world = dist.get_world_size(group)
        if world > _lib.OSSG_MAX_PEERS:
            raise ValueError(f"PeerHalo supports at most {_lib.OSSG_MAX_PEERS} ranks, got {world}")
        self.arch, self.fields, self.epoch = arch, list(fields), 0
        self.fold = PeerFold(fields, group=group)
        L = _lib.lib()
        dev = fields[0].data.device
        # one flag block and one epoch counter per phase: the epochs a flag block sees must be consecutive (1, 2, 3, ...),
        # because "rank r has finished exchange e" is an arrival COUNT (e x arrivals per kernel), not a stored epoch
        self.flags = [torch.zeros(2 * _lib.OSSG_MAX_PEERS + 1, dtype=torch.int64, device=dev) for _ in range(3)]
        self.epochs = [0, 0, 0]
        torch.cuda.synchronize(dev)

        def export(ptr):
            h, off = _lib.IpcHandle(), C.c_uint64()
            _lib.check(L.ossg_ipc_export(ptr, C.byref(h), C.byref(off)))
            return bytes(h.bytes), int(off.value)

        self._opened = []

        def open_(hb, off):
            h, out = _lib.IpcHandle(), C.c_void_p()
            C.memmove(h.bytes, hb, 64)
            _lib.check(L.ossg_ipc_open(C.byref(h), off, C.byref(out)))
            self._opened.append(hb)
            return out.value

        # staging buffers of the column phase (x split only): [batch][from_west | from_east], 8 bytes per element
        B = _lib.OSSG_MAX_PEER_BATCH
        g0 = self.fields[0].grid
        rows = (g0.Ny + 2 * g0.Hy) * (g0.Nz + 2 * g0.Hz)
        nbatches = (len(self.fields) + B - 1) // B
        self.staged = arch.Rx > 1
        self.stage = [[torch.empty(min(B, len(self.fields) - b * B) * rows * g0.Hx, dtype=torch.float64, device=dev) for _ in range(2)]
                      for b in range(nbatches)] if self.staged else []
        mine = {"flags": [export(t.data_ptr()) for t in self.flags], "fields": [export(f.data.data_ptr()) for f in self.fields],
                "stage": [[export(t.data_ptr()) for t in pair] for pair in self.stage]}
        gathered = [None] * world
        dist.all_gather_object(gathered, mine, group=group)
        me = arch.local_rank
        nb = {"north": arch.rank_of(arch.rx, arch.ry + 1) if arch.ry + 1 < arch.Ry else -1,
              "south": arch.rank_of(arch.rx, arch.ry - 1) if arch.ry > 0 else -1,
              "east": arch.rank_of((arch.rx + 1) % arch.Rx, arch.ry),
              "west": arch.rank_of((arch.rx - 1) % arch.Rx, arch.ry)}
        gN = self.fields[0].grid.global_size
        self.descs = []
        for phase in (0, 1, 2):
            hd = _lib.PeerHaloDesc()
            hd.me, hd.nranks, hd.phase = me, world, phase
            hd.north, hd.south, hd.east, hd.west = nb["north"], nb["south"], nb["east"], nb["west"]
            hd.flags_local = self.flags[phase].data_ptr()
            for r in set(v for v in nb.values() if v >= 0 and v != me):
                hd.flags_peer[r] = open_(*gathered[r]["flags"][phase])
            self.descs.append(hd)
        # per batch: where this rank pushes (the neighbours' buffers) and what it unpacks (its own)
        self.stage_ptrs = []
        for b in range(nbatches if self.staged else 0):
            self.stage_ptrs.append(dict(
                to_east=open_(*gathered[nb["east"]]["stage"][b][0]),      # east neighbour's "from west" buffer
                to_west=open_(*gathered[nb["west"]]["stage"][b][1]),      # west neighbour's "from east" buffer
                from_west=self.stage[b][0].data_ptr(), from_east=self.stage[b][1].data_ptr()))

        def size_of(rank):
            rx, ry = rank // arch.Ry, rank % arch.Ry
            _, _, nx, ny = tile_of(gN[0], gN[1], arch.Rx, arch.Ry, rx, ry)
            return nx, ny

        self._launches = []
        for b0 in range(0, len(self.fields), B):
            chunk = list(range(b0, min(b0 + B, len(self.fields))))
            arr = (_lib.PeerHaloField * len(chunk))()
            for n, k in enumerate(chunk):
                f = self.fields[k]
                arr[n].f = f.descriptor()
                for name in ("north", "south", "east", "west"):
                    r = nb[name]
                    ptr = None if r < 0 else (f.data.data_ptr() if r == me else open_(*gathered[r]["fields"][k]))
                    setattr(arr[n], name, ptr)
                arr[n].ny_north = size_of(nb["north"])[1] if nb["north"] >= 0 else 0
                arr[n].ny_south = size_of(nb["south"])[1] if nb["south"] >= 0 else 0
                arr[n].nx_east = size_of(nb["east"])[0]
                arr[n].nx_west = size_of(nb["west"])[0]
            self._launches.append((arr, len(chunk)))

    def __call__(self, stream=None) -> None:
        import ctypes as C
        L = _lib.lib()
        dev = self.fields[0].data.device
        stream = stream if stream is not None else torch.cuda.current_stream(dev).cuda_stream
        self.fold(stream)
        for phase in (0, 1, 2) if self.staged else (0, 1):
            for b, (arr, n) in enumerate(self._launches):
                hd = self.descs[phase]
                self.epochs[phase] += 1
                hd.epoch = self.epochs[phase]
                if self.staged and phase > 0:
                    sp = self.stage_ptrs[b]
                    hd.stage_to_east, hd.stage_to_west = sp["to_east"], sp["to_west"]
                    hd.stage_from_west, hd.stage_from_east = sp["from_west"], sp["from_east"]
                rc = L.ossg_halo_exchange_peer(arr, n, C.byref(hd), stream)
                if rc:
                    _lib.check(rc)

    def timed_out(self) -> bool:
        w = [int(t[2 * _lib.OSSG_MAX_PEERS].item()) >> 32 for t in self.flags]
        return any(w) or (self.fold.active and self.fold.timed_out())

    def close(self) -> None:
        import ctypes as C
        self.fold.close()
        if self._opened:
            torch.cuda.synchronize(self.fields[0].data.device)
            L = _lib.lib()
            for hb in self._opened:
                h = _lib.IpcHandle()
                C.memmove(h.bytes, hb, 64)
                L.ossg_ipc_close(C.byref(h))
            self._opened = []


def distributed_fill_halo_regions(fields: Sequence, stream=None, packer=None, group=None) -> None:
    """fill_halo_regions! for fields on a distributed tripolar grid: fold exchange on the top row
    (interior columns), then -- when x is not partitioned -- the local periodic west/east fill.
    Ordinary south/north and west/east neighbour exchanges are ordinary_halo_exchange()."""
    import contextlib
    ctx = contextlib.nullcontext()
    if stream is not None and fields and fields[0].data.is_cuda:
        # `stream`: a torch.cuda.Stream (or a raw cudaStream_t); every pack / send / recv / unpack of
        # the exchange is enqueued on it, ordered with the caller's producer kernels
        ts = stream if isinstance(stream, torch.cuda.Stream) else torch.cuda.ExternalStream(int(stream), device=fields[0].data.device)
        ctx = torch.cuda.stream(ts)
    with ctx:
        fold_exchange(fields, packer=packer, group=group)
        ordinary_halo_exchange(fields, packer=packer, group=group)


def ordinary_halo_exchange(fields: Sequence, packer=None, group=None) -> None:
    """South/north slabs between j-adjacent ranks, then west/east slabs (periodic in x) over the
    whole local y extent so the corners carry the already exchanged / folded rows.  One message
    per neighbour per direction per field batch."""
    import torch.distributed as dist
    if not fields:
        return
    arch = fields[0].grid.architecture
    packer = packer or CudaPacker()

    def exchange(pairs):
        """pairs: list of (send_view or None, send_peer, recv_view or None, recv_peer)."""
        ops, pending = [], []
        for (sv, sp, rv, rp) in pairs:
            if sv is not None:
                sb = sv.contiguous()
                ops.append(dist.P2POp(dist.isend, sb, sp, group))
                pending.append(sb)
            if rv is not None:
                rb = torch.empty(rv.shape, dtype=rv.dtype, device=rv.device)
                ops.append(dist.P2POp(dist.irecv, rb, rp, group))
                pending.append((rv, rb))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        for item in pending:
            if isinstance(item, tuple):
                packer.copy_rows(item[0], item[1])

    # ---- y direction (interior columns and x halos alike: whole rows)
    pairs = []
    for f in fields:
        g = f.grid
        Hy, Ny = g.Hy, g.Ny
        d = f.data
        if arch.ry + 1 < arch.Ry:       # send my top interior rows north, receive my north halo
            peer = arch.rank_of(arch.rx, arch.ry + 1)
            pairs.append((d[:, Ny:Ny + Hy, :], peer, d[:, Ny + Hy:Ny + 2 * Hy, :], peer))
        if arch.ry > 0:                 # send my bottom interior rows south, receive my south halo
            peer = arch.rank_of(arch.rx, arch.ry - 1)
            pairs.append((d[:, Hy:2 * Hy, :], peer, d[:, 0:Hy, :], peer))
    exchange(pairs)

    # ---- x direction, periodic
    pairs = []
    for f in fields:
        g = f.grid
        Hx, Nx = g.Hx, g.Nx
        d = f.data
        if arch.Rx == 1:
            d[:, :, 0:Hx] = d[:, :, Nx:Nx + Hx]
            d[:, :, Nx + Hx:Nx + 2 * Hx] = d[:, :, Hx:2 * Hx]
            continue
        east = arch.rank_of((arch.rx + 1) % arch.Rx, arch.ry)
        west = arch.rank_of((arch.rx - 1) % arch.Rx, arch.ry)
        pairs.append((d[:, :, Nx:Nx + Hx], east, d[:, :, 0:Hx], west))                     # my east interior -> east rank's west halo
        pairs.append((d[:, :, Hx:2 * Hx], west, d[:, :, Nx + Hx:Nx + 2 * Hx], east))       # my west interior -> west rank's east halo
    exchange(pairs)
