"""Host-side logic of the mirrored reference interface and the C-ABI surface, without a GPU:
symbols exported, argument errors, boundary-condition defaults, partition arithmetic, fold
communication plan, and loud failure when no CUDA device exists.  CPU only."""
import ctypes as C
import re

import numpy as np
import pytest
import torch

from oracle import oracle as orc

NO_GPU = not torch.cuda.is_available()


def test_library_loads_and_exports_every_declared_symbol(ossg):
    L = ossg._lib.lib()
    names = ossg._lib.declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/ossg_b200.h but not exported"
    assert L.ossg_version() == 1
    assert isinstance(L.ossg_last_error(), bytes)
    # the measurement / test utilities live in their own library, outside the product ABI
    U = ossg._lib.testutil()
    util_names = ossg._lib.declared_symbols(ossg._lib.UTIL_HEADER_PATH)
    assert set(util_names) >= {"ossg_synth_field", "ossg_fp64_peak_probe", "ossg_selftest_sqrt_div", "ossg_peer_copy_probe"}
    for n in util_names:
        assert hasattr(U, n), f"{n} declared in include/ossg_b200_testutil.h but not exported"
        assert n == "ossg_testutil_last_error" or not hasattr(L, n), f"{n} must not be exported by the product library"


def test_header_cites_the_reference_for_each_entry_point(ossg):
    text = open(ossg._lib.HEADER_PATH).read()
    for cite in ("src/tripolar_grid.jl:59-66", "src/zipper_boundary_condition.jl:70-138", "src/distributed_tripolar_grid.jl",
                 "src/tripolar_grid.jl:81-83", "src/tripolar_grid_extensions.jl:17"):
        assert cite in text


def test_struct_layouts_match_the_header(ossg):
    # sizes the C compiler gives the structs in include/ossg_b200.h (checked by compiling a probe)
    import subprocess, tempfile, os
    src = '#include "ossg_b200.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(ossg_tripolar_params), sizeof(ossg_field_desc), sizeof(ossg_fold_window), sizeof(ossg_ipc_handle), sizeof(ossg_peer_fold), sizeof(ossg_peer_field), sizeof(ossg_peer_halo), sizeof(ossg_peer_halo_field));return 0;}'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "p.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.dirname(ossg._lib.HEADER_PATH), os.path.join(d, "p.c"), "-o", os.path.join(d, "p")], check=True)
        out = subprocess.run([os.path.join(d, "p")], capture_output=True, text=True).stdout.split()
    L_ = ossg._lib
    assert [int(x) for x in out] == [C.sizeof(t) for t in (L_.TripolarParams, L_.FieldDesc, L_.FoldWindow, L_.IpcHandle, L_.PeerFoldDesc,
                                                           L_.PeerField, L_.PeerHaloDesc, L_.PeerHaloField)]


def test_constants_match_the_header(ossg):
    # the #define'd limits and codes of include/ossg_b200.h against their ctypes mirror
    import re
    text = open(ossg._lib.HEADER_PATH).read()
    defs = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+(OSSG_[A-Z0-9_]+)\s+\(?(-?\d+)\)?", text)}
    for name in ("OSSG_MAX_PEERS", "OSSG_MAX_PEER_BATCH", "OSSG_CENTER", "OSSG_FACE", "OSSG_F64", "OSSG_F32",
                 "OSSG_EINVAL", "OSSG_EODD_NX", "OSSG_EHALO", "OSSG_EWINDOW", "OSSG_EALIGN", "OSSG_ELOCATION", "OSSG_EWORKSPACE"):
        assert name in defs, name
        assert defs[name] == getattr(ossg._lib, name), name


def test_argument_errors_do_not_need_a_device(ossg):
    """Validation happens before any CUDA call: same messages the reference throws."""
    from ossg_b200.grids import make_params
    L = ossg._lib.lib()
    out = ossg._lib.ptr_array([8] * 20)
    p = make_params((61, 30, 1), (4, 4, 4), -80, 55, 70, orc.R_EARTH)
    assert L.ossg_tripolar_build(C.byref(p), 0, out, 16, 1 << 20, None) == ossg._lib.OSSG_EODD_NX
    assert "The number of cells in the longitude dimension should be even!" in ossg._lib.last_error()   # tripolar_grid.jl:82
    p = make_params((60, 30, 1), (4, 30, 4), -80, 55, 70, orc.R_EARTH)
    assert L.ossg_tripolar_build(C.byref(p), 0, out, 16, 1 << 20, None) == ossg._lib.OSSG_EHALO
    p = make_params((60, 30, 1), (4, 4, 4), -80, 55, 70, orc.R_EARTH, tile=(31, 1, 31, 30))
    assert L.ossg_tripolar_build(C.byref(p), 0, out, 16, 1 << 20, None) == ossg._lib.OSSG_EWINDOW
    p = make_params((60, 30, 1), (4, 4, 4), -80, 55, 70, orc.R_EARTH)
    assert L.ossg_tripolar_build(C.byref(p), 7, out, 16, 1 << 20, None) == ossg._lib.OSSG_EINVAL
    assert L.ossg_tripolar_build(C.byref(p), 0, out, 16, 8, None) == ossg._lib.OSSG_EWORKSPACE
    assert L.ossg_tripolar_build(C.byref(p), 0, ossg._lib.ptr_array([4] * 20), 16, 1 << 20, None) == ossg._lib.OSSG_EALIGN
    tb = ((16 + 4 * 60 + 4 * 30 + 9 * 5 + 1) // 2 * 2) * 8          # 128-byte stamp header + x/y tables + lat-lon rows
    assert L.ossg_tripolar_table_bytes(C.byref(p)) == tb
    rc = (38 * 4 + 15) // 16 * 16                                   # one 4-byte counter per window row (coordinate -> tile dependency)
    assert L.ossg_tripolar_workspace_bytes(C.byref(p), 0) == tb + rc
    assert L.ossg_tripolar_workspace_bytes(C.byref(p), 1) == tb + 8 * 68 * 38 * 8 + rc
    assert L.ossg_tripolar_workspace_init(None, 1 << 20, None) == ossg._lib.OSSG_EALIGN
    assert L.ossg_tripolar_workspace_init(16, 64, None) == ossg._lib.OSSG_EWORKSPACE
    f = ossg._lib.FieldDesc(8, 0, 10, 10, 1, 4, 4, 4, 2, 0, 0, 0, 1.0)
    assert L.ossg_zipper_fill(C.byref(f), None) == ossg._lib.OSSG_ELOCATION
    f = ossg._lib.FieldDesc(8, 0, 11, 10, 1, 4, 4, 4, 0, 0, 0, 0, 1.0)
    assert L.ossg_zipper_fill(C.byref(f), None) == ossg._lib.OSSG_EODD_NX
    f = ossg._lib.FieldDesc(0, 0, 10, 10, 1, 4, 4, 4, 0, 0, 0, 0, 1.0)
    assert L.ossg_zipper_fill(C.byref(f), None) == ossg._lib.OSSG_EINVAL


def test_odd_longitude_size_raises_like_the_reference(ossg):
    with pytest.raises(ValueError, match="should be even"):           # src/tripolar_grid.jl:81-83
        ossg.TripolarGrid(size=(61, 30, 1))


@pytest.mark.skipif(not NO_GPU, reason="only meaningful on a machine without a GPU")
def test_no_cpu_fallback(ossg):
    """The product path must fail loudly when there is no CUDA device (no oracle, no torch path)."""
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ossg.TripolarGrid(ossg.GPU(), size=(60, 30, 1))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ossg.TripolarGrid(ossg.CPU(), size=(60, 30, 1))
    assert ossg._lib.lib().ossg_device_ok() != 0


def test_product_never_imports_the_oracle():
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "orthogonalsphericalshellgrids.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")) or fn == "Makefile":
                text = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"(from|import)\s+oracle|ossg_oracle|oracle/", text), os.path.join(dirpath, fn)


class _StubGrid:
    """Enough of a grid for Field / boundary-condition logic on the CPU."""
    def __init__(self, ossg, size=(10, 10, 1), halo=(4, 4, 4), arch=None, i0=1, j0=1, global_size=None):
        self.Nx, self.Ny, self.Nz = size
        self.Hx, self.Hy, self.Hz = halo
        self.architecture = arch if arch is not None else ossg.GPU()
        self.dtype = torch.float64
        self.arrays = {}
        self.i0, self.j0 = i0, j0
        self.global_size = global_size or size
        self.size, self.halo = size, halo


def _field(ossg, grid, loc, **kw):
    shape = (grid.Nz + 2 * grid.Hz, grid.Ny + 2 * grid.Hy, grid.Nx + 2 * grid.Hx)
    return ossg.Field(loc, grid, data=torch.zeros(shape, dtype=torch.float64), **kw)


def test_default_zipper_signs(ossg):
    """test/test_zipper_boundary_conditions.jl:14-23; src/tripolar_grid_extensions.jl:49-53,57-80"""
    g = _StubGrid(ossg)
    C_, F_ = ossg.Center, ossg.Face
    for loc, s in (((C_, C_, C_), 1), ((F_, C_, C_), -1), ((C_, F_, C_), -1), ((F_, F_, C_), 1)):
        f = _field(ossg, g, loc)
        assert isinstance(f.boundary_conditions.north.classification, ossg.Zipper)
        assert f.boundary_conditions.north.condition == s
        assert f.descriptor().sign == s and ossg.sign(loc[0], loc[1]) == s
    # a user-supplied zipper is kept (:69), anything else on the north side is replaced
    keep = ossg.FieldBoundaryConditions(north=ossg.ZipperBoundaryCondition(1))
    assert _field(ossg, g, (F_, C_, C_), boundary_conditions=keep).boundary_conditions.north.condition == 1
    repl = ossg.FieldBoundaryConditions(north=ossg.PeriodicBoundaryCondition())
    assert _field(ossg, g, (F_, C_, C_), boundary_conditions=repl).boundary_conditions.north.condition == -1
    assert _field(ossg, g, (C_, C_, C_), boundary_conditions=None).boundary_conditions is None      # :62
    assert ossg.ZipperBoundaryCondition().condition == 1 and ossg.bc_str(ossg.ZipperBoundaryCondition()) == "Zipper"


def test_horizontal_field_and_immersed_wrapper_host_logic(ossg):
    """(Center, Center, Nothing) fields (the immersed-boundary bottom height) have one
    level and no z halo in their descriptor; an ImmersedBoundaryGrid forwards every grid attribute
    (the TRG union, src/tripolar_grid.jl:371) and cannot be built without a device."""
    g = _StubGrid(ossg, size=(10, 8, 5), halo=(3, 2, 4))
    f = ossg.Field((ossg.Center, ossg.Center, None), g, data=torch.zeros((1, 8 + 4, 10 + 6), dtype=torch.float64))
    d = f.descriptor()
    assert (d.Nx, d.Ny, d.Nz, d.Hx, d.Hy, d.Hz, d.locx, d.locy, d.sign) == (10, 8, 1, 3, 2, 0, 0, 0, 1.0)
    assert tuple(f.interior().shape) == (1, 8, 10)
    ossg.set_(f, np.arange(80.0).reshape(8, 10))
    assert float(f[10, 8, 1]) == 79.0 and float(f[1, 1, 1]) == 0.0
    with pytest.raises(ValueError):
        ossg.Field((ossg.Center, ossg.Center, None), g, data=torch.zeros((5 + 8, 12, 16), dtype=torch.float64))
    from ossg_b200.immersed import ImmersedBoundaryGrid
    w = ImmersedBoundaryGrid.__new__(ImmersedBoundaryGrid)
    w.underlying_grid = g
    assert (w.Nx, w.Ny, w.Nz, w.Hy) == (10, 8, 5, 2) and w.architecture is g.architecture
    assert _field(ossg, w, (ossg.Face, ossg.Center, ossg.Center)).boundary_conditions.north.condition == -1
    if NO_GPU:
        g.arrays = {"lam_cc": None}
        with pytest.raises(Exception):
            ossg.ImmersedBoundaryGrid(g, ossg.GridFittedBottom(0.0))


def test_zipper_is_north_only(ossg):
    """src/zipper_boundary_condition.jl:58-62"""
    g = _StubGrid(ossg)
    for side in ("south", "west", "east", "bottom", "top"):
        bcs = ossg.FieldBoundaryConditions(**{side: ossg.ZipperBoundaryCondition()})
        with pytest.raises(ValueError, match="north only"):
            _field(ossg, g, (ossg.Center, ossg.Center, ossg.Center), boundary_conditions=bcs)
    assert ossg.validate_boundary_condition_location(ossg.Zipper(), ossg.Center, "north") is None


def test_regularize_field_boundary_conditions(ossg):
    """src/tripolar_grid_extensions.jl:25-44: sign -1 for :u and :v, +1 otherwise."""
    g = _StubGrid(ossg)
    for name, s in (("u", -1), ("v", -1), ("c", 1), ("T", 1), ("w", 1)):
        r = ossg.regularize_field_boundary_conditions(ossg.FieldBoundaryConditions(), g, name)
        assert r.north.condition == s and isinstance(r.north.classification, ossg.Zipper)


def test_partition_arithmetic(ossg):
    """src/distributed_tripolar_grid.jl:41-49: equal split, the last rank takes the remainder."""
    from ossg_b200.distributed import local_sizes, tile_of
    assert local_sizes(240, 4) == [60, 60, 60, 60] and local_sizes(10, 3) == [3, 3, 4] and local_sizes(7, 1) == [7]
    seen = np.zeros((37, 50), dtype=int)
    for rx in range(3):
        for ry in range(4):
            i0, j0, nx, ny = tile_of(50, 37, 3, 4, rx, ry)
            seen[j0 - 1:j0 - 1 + ny, i0 - 1:i0 - 1 + nx] += 1
    assert (seen == 1).all()
    a = ossg.Distributed(partition=ossg.Partition(2, 3), rank=4, world_size=6)
    assert (a.rx, a.ry) == (1, 1) and not a.is_top_row() and a.rank_of(1, 2) == 5
    assert ossg.Distributed(partition=ossg.Partition(2, 3), rank=5, world_size=6).is_top_row()
    with pytest.raises(ValueError, match="Y-partitioning"):            # :30-31, opt-in strictness
        ossg.Distributed(partition=ossg.Partition(2, 1), rank=0, world_size=2, y_slabs_only=True)
    with pytest.raises(ValueError):
        ossg.Distributed(partition=ossg.Partition(2, 2), rank=0, world_size=2)


@pytest.mark.parametrize("Nx,Rx", [(24, 1), (24, 2), (24, 3), (24, 4), (26, 4), (10, 5), (4320, 8)])
@pytest.mark.parametrize("locx", [0, 1])
def test_fold_plan_covers_every_destination_column_exactly_once(ossg, Nx, Rx, locx):
    from ossg_b200.distributed import fold_plan, local_sizes
    msgs = fold_plan(Nx, Rx, 2, locx)
    sx = local_sizes(Nx, Rx)
    starts = [1 + sum(sx[:r]) for r in range(Rx)]
    need = {}
    for m in msgs:
        assert m.src % 2 == 1 and m.dst % 2 == 1                      # top row of Ry = 2
        srx, drx = m.src // 2, m.dst // 2
        assert starts[srx] <= m.src_i0 and m.src_i0 + m.n - 1 <= starts[srx] + sx[srx] - 1   # sender owns the columns
        for c in range(m.src_i0, m.src_i0 + m.n):
            need.setdefault(drx, []).append(c)
    for drx in range(Rx):
        want = []
        for i in range(starts[drx], starts[drx] + sx[drx]):
            ip = Nx - i + 1 + locx
            want.append(ip - Nx if ip > Nx else ip)
        assert sorted(need[drx]) == sorted(want)
    if Rx > 1 and locx == 0 and Nx % Rx == 0:
        # x-Center, equal split: exactly one partner per rank, the mirror rank
        assert sorted((m.src // 2, m.dst // 2) for m in msgs) == sorted((Rx - 1 - r, r) for r in range(Rx))


def test_with_halo_and_reconstruct_forward_the_tripolar_parameters(ossg, monkeypatch):
    """src/with_halo.jl:5-23 and src/distributed_tripolar_grid.jl:201-226 re-run the constructor
    with the conformal-mapping scalars; checked by intercepting the constructor."""
    import ossg_b200.grids as G
    calls = []

    def fake(arch=None, FT=None, **kw):
        calls.append((arch, FT, kw))
        return "grid"
    monkeypatch.setattr(G, "TripolarGrid", fake)
    old = _StubGrid(ossg, size=(20, 12, 3), halo=(4, 4, 4))
    old.conformal_mapping = ossg.Tripolar(45, 25, -75)
    old.radius = 1.0
    old.z = G._generate_z((-100, 0), 3)
    assert G.with_halo((4, 13, 4), old) == "grid"
    arch, FT, kw = calls[-1]
    assert kw["halo"] == (4, 13, 4) and kw["size"] == (20, 12, 3) and kw["radius"] == 1.0
    assert (kw["north_poles_latitude"], kw["first_pole_longitude"], kw["southernmost_latitude"]) == (45, 25, -75)
    assert np.allclose(kw["z"], [-100, -100 + 100 / 3, -100 + 200 / 3, 0])


def test_reference_arm_of_the_bench_runs_without_a_gpu():
    """`bench.py --impl reference` times the CPU restatement of the reference (the oracle built with glibc libm) on the
    host cores and prints ONE JSON line with the contract's keys; it needs no GPU and must not touch the CUDA library."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "grid cells/s for coords+metrics (FP64)" and d["unit"] == "cells/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "4320x2160" in d["config"]["workload"]
