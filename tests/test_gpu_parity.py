"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI of
libossg_b200.so (via the Python host mirror of the reference interface) and is compared with the
CPU oracle on identical inputs.

Tolerances (BASELINE.json north_star): lam/phi and metrics <= 1e-12 relative in Float64 -- the
implementation is designed to be bit-identical to the oracle's deterministic arithmetic, so the
tests also report the number of differing bit patterns -- and the fold is bit-exact.
Float32: FP64 compute rounded once, i.e. identical to rounding the oracle's FP64 result.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import oracle as orc

import os as _os
ROOT = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu

REL_TOL = 1e-12          # north_star tolerance, Float64

CONFIGS = [
    # size, halo, kwargs                                                   (source)
    ((60, 30, 1), (4, 4, 4), {}),                                         # README.md:54 / BASELINE config A
    ((4, 5, 1), (4, 4, 4), dict(first_pole_longitude=75, north_poles_latitude=35)),    # test/runtests.jl:10-14
    ((10, 10, 1), (4, 4, 4), {}),                                         # test_zipper_boundary_conditions.jl:6 (Nx % 4 != 0)
    ((360, 180, 1), (4, 4, 4), dict(first_pole_longitude=75, north_poles_latitude=35)),  # test_tripolar_grid.jl:59
    ((360, 180, 50), (4, 4, 4), {}),                                      # BASELINE config B
    ((120, 60, 1), (5, 5, 5), dict(north_poles_latitude=45, first_pole_longitude=25)),  # examples/bickley_jet.jl:20-23
    ((40, 34, 1), (5, 31, 5), dict(north_poles_latitude=45, first_pole_longitude=35)),  # large Hy (with_halo, runtests.jl:71)
    ((12, 8, 1), (2, 7, 1), {}),                                          # Hy = Ny - 1 (degenerate fold sources)
    ((66, 37, 1), (3, 2, 1), dict(southernmost_latitude=-75.5)),          # ragged sizes, non-integer south
    ((1440, 720, 1), (4, 4, 4), {}),                                      # BASELINE config C: the tile kernel's straight-line path
    ((2880, 1440, 1), (4, 4, 4), dict(north_poles_latitude=60, first_pole_longitude=-110)),   # 1/8 degree, other poles
]


def rel_err(got, want):
    denom = np.maximum(np.abs(want), np.finfo(np.float64).tiny)
    err = np.where(got == want, 0.0, np.abs(got - want) / denom)
    err = np.where(np.isnan(got) & np.isnan(want), 0.0, err)
    return err


def lam_err(got, want):
    """longitudes compare modulo 360"""
    d = np.abs(got - want)
    d = np.minimum(d, np.abs(360.0 - d))
    return d / np.maximum(np.abs(want), 1.0)


def assert_grid_matches(grid_arrays, ref, tol=REL_TOL):
    nbits = 0
    for name in orc.ALL_NAMES:
        got = grid_arrays[name]
        want = ref[name]
        assert got.shape == want.shape, name
        err = lam_err(got, want) if name.startswith("lam") else rel_err(got, want)
        worst = float(err.max())
        assert worst <= tol, f"{name}: max rel err {worst:.3e} at {np.unravel_index(err.argmax(), err.shape)}"
        # exact zeros must be exact (pole metrics), src/tripolar_grid.jl M2 consequence
        assert np.array_equal(got == 0.0, want == 0.0), f"{name}: zero pattern differs"
        nbits += int((got.view(np.uint64) != want.view(np.uint64)).sum())
    return nbits


def to_np(grid):
    return {n: grid.arrays[n].parent.cpu().numpy() for n in orc.ALL_NAMES}


@pytest.mark.parametrize("size,halo,kw", CONFIGS)
def test_construction_matches_oracle_f64(ossg, oracle, size, halo, kw):
    grid = ossg.TripolarGrid(ossg.GPU(), ossg.Float64, size=size, halo=halo, **kw)
    ref = oracle.build(orc.make_params(size, halo, **kw))
    nbits = assert_grid_matches(to_np(grid), ref)
    assert nbits == 0, f"{nbits} elements are not bit-identical to the oracle"


@pytest.mark.parametrize("size,halo,kw", CONFIGS[:5])
def test_construction_matches_oracle_f32(ossg, oracle, size, halo, kw):
    grid = ossg.TripolarGrid(ossg.GPU(), ossg.Float32, size=size, halo=halo, **kw)
    assert grid.eltype() == torch.float32
    ref = oracle.build(orc.make_params(size, halo, **kw), dtype=np.float32)
    for name in orc.ALL_NAMES:
        got = grid.arrays[name].parent.cpu().numpy()
        assert got.dtype == np.float32
        assert np.array_equal(got.view(np.uint32), ref[name].view(np.uint32)), name


def test_readme_known_answers_on_gpu(ossg):
    """README.md:54-59"""
    grid = ossg.TripolarGrid(size=(60, 30, 1))
    p = orc.make_params((60, 30, 1))
    s = orc.readme_summary(to_np(grid), p)
    assert round(s.centre[0], 4) == 70.0 and round(s.centre[1], 4) == 1.8005
    assert f"{s.lon_extent:.6g}" == "359.885" and f"{s.lat_extent:.6g}" == "175.667"
    assert f"{s.dlam_min:.6g}" == "0.279019" and f"{s.dlam_max:.6g}" == "6.32049"
    assert f"{s.dphi_min:.6g}" == "0.429975" and f"{s.dphi_max:.6g}" == "5.86207"


def _call_staged(ossg, params, coords_host, metrics_dtype=torch.float64):
    """ossg_tripolar_metrics on coordinates supplied by the caller (here: the oracle's)."""
    L = ossg._lib.lib()
    dev = torch.device("cuda", torch.cuda.current_device())
    cin = [torch.as_tensor(coords_host[n], device=dev).contiguous() for n in orc.COORD_NAMES]
    out = [torch.empty_like(cin[0], dtype=metrics_dtype) for _ in range(12)]
    nbytes = L.ossg_tripolar_workspace_bytes(C.byref(params), ossg._lib.OSSG_F64)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    et = ossg._lib.OSSG_F64 if metrics_dtype == torch.float64 else ossg._lib.OSSG_F32
    rc = L.ossg_tripolar_metrics(C.byref(params), ossg._lib.ptr_array([t.data_ptr() for t in cin]), et,
                                 ossg._lib.ptr_array([t.data_ptr() for t in out]), ws.data_ptr(), ws.numel(),
                                 torch.cuda.current_stream().cuda_stream)
    ossg._lib.check(rc)
    torch.cuda.synchronize()
    return dict(zip(orc.METRIC_NAMES, [t.cpu().numpy() for t in out]))


@pytest.mark.parametrize("size,halo,kw", [CONFIGS[0], CONFIGS[3], CONFIGS[6]])
def test_staged_metrics_from_oracle_coordinates(ossg, oracle, oracle_libm, size, halo, kw):
    """Metric kernels fed with coordinates that did NOT come from the GPU: the oracle's own, and
    the glibc-libm oracle's (a different last-ulp rounding of every coordinate)."""
    from ossg_b200.grids import make_params
    for orac in (oracle, oracle_libm):
        p = orc.make_params(size, halo, **kw)
        coords = orac.coords(p)
        want = oracle.metrics(p, coords)        # deterministic-math oracle on the same inputs
        params = make_params(size, halo, kw.get("southernmost_latitude", -80), kw.get("north_poles_latitude", 55),
                             kw.get("first_pole_longitude", 70), orc.R_EARTH)
        got = _call_staged(ossg, params, coords)
        for name in orc.METRIC_NAMES:
            err = rel_err(got[name], want[name])
            assert float(err.max()) <= REL_TOL, name
            assert np.array_equal(got[name].view(np.uint64), want[name].view(np.uint64)), name


def _call_coords(ossg, params):
    L = ossg._lib.lib()
    dev = torch.device("cuda", torch.cuda.current_device())
    nx, ny = params.nx or params.Nx, params.ny or params.Ny
    out = [torch.empty((ny + 2 * params.Hy, nx + 2 * params.Hx), dtype=torch.float64, device=dev) for _ in range(8)]
    nbytes = L.ossg_tripolar_workspace_bytes(C.byref(params), ossg._lib.OSSG_F64)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    ossg._lib.check(L.ossg_tripolar_coords(C.byref(params), ossg._lib.ptr_array([t.data_ptr() for t in out]),
                                           ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return dict(zip(orc.COORD_NAMES, [t.cpu().numpy() for t in out]))


def test_staged_coords(ossg, oracle):
    from ossg_b200.grids import make_params
    size, halo = (60, 30, 1), (4, 4, 4)
    got = _call_coords(ossg, make_params(size, halo, -80, 55, 70, orc.R_EARTH))
    want = oracle.coords(orc.make_params(size, halo))
    for n in orc.COORD_NAMES:
        assert np.array_equal(got[n].view(np.uint64), want[n].view(np.uint64)), n


@pytest.mark.parametrize("Rx,Ry", [(1, 2), (2, 1), (2, 2), (4, 2), (3, 3), (1, 4)])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_tiles_equal_slices_of_the_serial_grid(ossg, oracle, Rx, Ry, dtype):
    """src/distributed_tripolar_grid.jl:41-49,112-120: a rank's arrays are the jrange (and, new,
    irange) slice of the global padded arrays, offsets preserved."""
    from ossg_b200.distributed import tile_of
    from ossg_b200.grids import build_arrays, make_params
    size, halo = (48, 40, 1), (4, 4, 4)
    npdt = np.float64 if dtype == torch.float64 else np.float32
    ref = oracle.build(orc.make_params(size, halo), dtype=npdt)
    dev = torch.device("cuda", torch.cuda.current_device())
    for rx in range(Rx):
        for ry in range(Ry):
            i0, j0, nx, ny = tile_of(size[0], size[1], Rx, Ry, rx, ry)
            params = make_params(size, halo, -80, 55, 70, orc.R_EARTH, tile=(i0, j0, nx, ny))
            tiles = build_arrays(params, dtype, dev)
            torch.cuda.synchronize()
            for name, t in zip(orc.ALL_NAMES, tiles):
                want = ref[name][j0 - 1:j0 - 1 + ny + 2 * halo[1], i0 - 1:i0 - 1 + nx + 2 * halo[0]]
                got = t.cpu().numpy()
                assert np.array_equal(got.view(np.uint8), np.ascontiguousarray(want).view(np.uint8)), (name, rx, ry)


@pytest.mark.parametrize("size,halo", [((20, 12, 1), (4, 4, 4)), ((10, 10, 1), (4, 4, 4)), ((26, 14, 1), (3, 5, 1)), ((64, 20, 1), (8, 2, 1))])
def test_arbitrary_windows_equal_slices_of_the_serial_grid(ossg, oracle, size, halo):
    """Windows that are not tiles of a regular partition (any i0, nx, j0, ny the C ABI accepts): halves and other
    x ranges whose padded columns wrap around either end of the grid, with and without their own mirror columns --
    every launch form of the coordinate kernel (pair form for whole-x and mirror-closed windows, bounding thread
    range with idle threads, per-element form) against slices of the oracle's serial grid, Nx % 4 != 0 included."""
    from ossg_b200.grids import build_arrays, make_params
    Nx, Ny = size[0], size[1]
    Hx, Hy = halo[0], halo[1]
    ref = oracle.build(orc.make_params(size, halo))
    dev = torch.device("cuda", torch.cuda.current_device())
    xs = [(1, Nx), (1, Nx // 2), (Nx // 2 + 1, Nx - Nx // 2), (1, 1), (Nx, 1), (2, Nx - 2), (Nx // 4 + 1, Nx // 2), (3, max(1, Nx // 3)), (Nx - 2, 3)]
    ys = [(1, Ny), (1, max(1, Ny // 2)), (Ny // 2 + 1, Ny - Ny // 2), (Ny, 1), (Ny - 2, 3), (2, 1)]
    for (i0, nx) in xs:
        for (j0, ny) in ys:
            if (ny - 1 < Hy and j0 + ny - 1 == Ny and False):
                continue
            params = make_params(size, halo, -80, 55, 70, orc.R_EARTH, tile=(i0, j0, nx, ny))
            tiles = build_arrays(params, torch.float64, dev)
            torch.cuda.synchronize()
            for name, t in zip(orc.ALL_NAMES, tiles):
                want = ref[name][j0 - 1:j0 - 1 + ny + 2 * Hy, i0 - 1:i0 - 1 + nx + 2 * Hx]
                assert np.array_equal(t.cpu().numpy().view(np.uint64), np.ascontiguousarray(want).view(np.uint64)), (name, i0, nx, j0, ny)


@pytest.mark.parametrize("Rx,Ry,dtype", [(1, 8, torch.float64), (2, 2, torch.float64), (3, 2, torch.float32), (1, 4, torch.float32)])
def test_fine_grid_tiles_equal_slices_of_the_serial_grid(ossg, oracle, Rx, Ry, dtype):
    """The same on a 1/4 degree grid, where every tile runs the tile kernel's straight-line path,
    the mirror-pair coordinate kernel (Rx = 1) and the per-element one (Rx > 1): the layouts of
    bench.py's multi-GPU runs (Y slabs) and of the (x, y) partitions."""
    from ossg_b200.distributed import tile_of
    from ossg_b200.grids import build_arrays, make_params
    size, halo = (1440, 720, 1), (4, 4, 4)
    npdt = np.float64 if dtype == torch.float64 else np.float32
    oracle.set_threads(oracle.max_threads())
    ref = oracle.build(orc.make_params(size, halo), dtype=npdt)
    dev = torch.device("cuda", torch.cuda.current_device())
    for rx in range(Rx):
        for ry in range(Ry):
            i0, j0, nx, ny = tile_of(size[0], size[1], Rx, Ry, rx, ry)
            params = make_params(size, halo, -80, 55, 70, orc.R_EARTH, tile=(i0, j0, nx, ny))
            tiles = build_arrays(params, dtype, dev)
            torch.cuda.synchronize()
            for name, t in zip(orc.ALL_NAMES, tiles):
                want = ref[name][j0 - 1:j0 - 1 + ny + 2 * halo[1], i0 - 1:i0 - 1 + nx + 2 * halo[0]]
                got = t.cpu().numpy()
                assert np.array_equal(got.view(np.uint8), np.ascontiguousarray(want).view(np.uint8)), (name, rx, ry)


# ------------------------------------------------------------------ zipper
LOCS = {"cc": (0, 0), "fc": (1, 0), "cf": (0, 1), "ff": (1, 1)}


def _poisoned_field(shape, Nx, Ny, Nz, Hx, Hy, Hz, npdt, seed):
    rng = np.random.default_rng(seed)
    c = np.full(shape, np.nan, dtype=npdt)
    c[Hz:Hz + Nz, Hy:Hy + Ny, Hx:Hx + Nx] = rng.uniform(-1, 1, (Nz, Ny, Nx)).astype(npdt)
    return c


def _desc(ossg, t, dims, loc, sign, k0=0, nk=0):
    Nx, Ny, Nz, Hx, Hy, Hz = dims
    et = ossg._lib.OSSG_F64 if t.dtype == torch.float64 else ossg._lib.OSSG_F32
    return ossg._lib.FieldDesc(t.data_ptr(), et, Nx, Ny, Nz, Hx, Hy, Hz, loc[0], loc[1], k0, nk, float(sign))


def _same_bits(a, b):
    u = np.uint64 if a.dtype == np.float64 else np.uint32
    return np.array_equal(a.view(u), b.view(u))


@pytest.mark.parametrize("locname", list(LOCS))
@pytest.mark.parametrize("npdt", [np.float64, np.float32])
@pytest.mark.parametrize("sign", [1, -1])
@pytest.mark.parametrize("dims", [(10, 10, 1, 4, 4, 4), (60, 30, 3, 4, 4, 2), (360, 180, 5, 4, 4, 4), (34, 17, 2, 3, 9, 1), (4, 5, 1, 4, 4, 4)])
def test_zipper_fill_bit_exact(ossg, oracle, locname, npdt, sign, dims):
    """src/zipper_boundary_condition.jl:70-138 on NaN-poisoned halos: exactly the fold targets are
    written, everything else (other halos, interior) is untouched, bit for bit."""
    Nx, Ny, Nz, Hx, Hy, Hz = dims
    loc = LOCS[locname]
    shape = (Nz + 2 * Hz, Ny + 2 * Hy, Nx + 2 * Hx)
    host = _poisoned_field(shape, Nx, Ny, Nz, Hx, Hy, Hz, npdt, seed=hash((locname, sign)) % 1000)
    dev = torch.as_tensor(host).cuda()
    d = _desc(ossg, dev, dims, loc, sign)
    ossg._lib.check(ossg._lib.lib().ossg_zipper_fill(C.byref(d), None))
    torch.cuda.synchronize()
    want = oracle.zipper(host.copy(), Nx, Ny, Nz, Hx, Hy, Hz, loc[0], loc[1], sign)
    assert _same_bits(dev.cpu().numpy(), want)
    # zipper + periodic x (fill_halo_regions!)
    dev2 = torch.as_tensor(host).cuda()
    d2 = _desc(ossg, dev2, dims, loc, sign)
    ossg._lib.check(ossg._lib.lib().ossg_fill_halo_regions_batch(C.byref(d2), 1, None))
    torch.cuda.synchronize()
    want2 = oracle.fill_halo_regions(host.copy(), Nx, Ny, Nz, Hx, Hy, Hz, loc[0], loc[1], sign)
    assert _same_bits(dev2.cpu().numpy(), want2)


def test_zipper_batch_mixed_fields_one_launch(ossg, oracle):
    dims = (60, 30, 4, 4, 4, 3)
    Nx, Ny, Nz, Hx, Hy, Hz = dims
    shape = (Nz + 2 * Hz, Ny + 2 * Hy, Nx + 2 * Hx)
    specs = [("cc", 1, np.float64), ("fc", -1, np.float64), ("cf", -1, np.float32), ("ff", 1, np.float32),
             ("fc", -1, np.float32), ("cc", 1, np.float32)]
    hosts = [_poisoned_field(shape, *dims, npdt, seed=k) for k, (_, _, npdt) in enumerate(specs)]
    devs = [torch.as_tensor(h).cuda() for h in hosts]
    arr = (ossg._lib.FieldDesc * len(specs))(*[_desc(ossg, t, dims, LOCS[l], s) for t, (l, s, _) in zip(devs, specs)])
    ossg._lib.check(ossg._lib.lib().ossg_zipper_fill_batch(arr, len(specs), None))
    torch.cuda.synchronize()
    for t, h, (l, s, _) in zip(devs, hosts, specs):
        want = oracle.zipper(h.copy(), Nx, Ny, Nz, Hx, Hy, Hz, LOCS[l][0], LOCS[l][1], s)
        assert _same_bits(t.cpu().numpy(), want), l


def test_zipper_k_range(ossg, oracle):
    dims = (20, 12, 6, 2, 3, 2)
    Nx, Ny, Nz, Hx, Hy, Hz = dims
    shape = (Nz + 2 * Hz, Ny + 2 * Hy, Nx + 2 * Hx)
    host = _poisoned_field(shape, *dims, np.float64, seed=3)
    dev = torch.as_tensor(host).cuda()
    d = _desc(ossg, dev, dims, (1, 0), -1, k0=3, nk=4)
    ossg._lib.check(ossg._lib.lib().ossg_zipper_fill(C.byref(d), None))
    torch.cuda.synchronize()
    want = oracle.zipper(host.copy(), Nx, Ny, Nz, Hx, Hy, Hz, 1, 0, -1, k0=3, nk=4)
    assert _same_bits(dev.cpu().numpy(), want)


def test_reference_zipper_testset(ossg):
    """test/test_zipper_boundary_conditions.jl:5-45 through the mirrored API."""
    grid = ossg.TripolarGrid(size=(10, 10, 1))
    Nx, Ny, _ = grid.size
    Hx, Hy, _ = grid.halo
    c, u, v = ossg.CenterField(grid), ossg.XFaceField(grid), ossg.YFaceField(grid)
    for f, s in ((c, 1), (u, -1), (v, -1)):
        assert isinstance(f.boundary_conditions.north.classification, ossg.Zipper)      # :14-16
        assert f.boundary_conditions.north.condition == s                               # :21-23
    for f in (c, u, v):
        ossg.set_(f, 1)
    ossg.fill_halo_regions(c)
    ossg.fill_halo_regions(u)
    ossg.fill_halo_regions(v)
    torch.cuda.synchronize()
    assert bool((c[slice(None, None), slice(Ny + 1, Ny + Hy), 1] == 1).all())             # :35
    assert bool((v[slice(None, None), slice(Ny + 1, Ny + Hy), 1] == -1).all())            # :36
    assert bool((u[slice(2, Nx - 1), slice(Ny + 1, Ny + Hy), 1] == -1).all())             # :39-40
    assert bool((u[1, slice(Ny + 1, Ny + Hy), 1] == 1).all())                             # :42,44
    assert bool((u[Nx + 1, slice(Ny + 1, Ny + Hy), 1] == 1).all())                        # :43,45


def test_reference_zipper_testset_fold_row(ossg):
    """test/test_zipper_boundary_conditions.jl:56-72: c = u = x, row Ny after the fill."""
    grid = ossg.TripolarGrid(size=(10, 10, 1))
    c, u = ossg.CenterField(grid), ossg.XFaceField(grid)
    ossg.set_(c, lambda lam, phi, z: lam + 0 * phi + 0 * z)
    ossg.set_(u, lambda lam, phi, z: lam + 0 * phi + 0 * z)
    ossg.fill_halo_regions([c, u])
    torch.cuda.synchronize()
    row = c.interior()[0, 9, :].cpu().numpy()
    assert np.array_equal(row, row[::-1])                                                # :65
    urow = u.interior()[0, 9, :].cpu().numpy()
    assert np.array_equal(urow[1:5], -urow[6:10][::-1])                                  # :68-72


def test_immersed_bottom_height_is_symmetric_across_the_fold(ossg, oracle):
    """test/test_zipper_boundary_conditions.jl:46-54: a random bottom under
    ImmersedBoundaryGrid(grid, GridFittedBottom(bottom)) is mirror-symmetric in row Ny; the whole
    (Center, Center, Nothing) field equals the oracle's zipper + periodic fill of the same input.
    Then the pole mask of examples/bickley_jet.jl:25-29 and fields on the wrapped grid."""
    under = ossg.TripolarGrid(size=(10, 10, 1))
    Nx, Ny, Hx, Hy = under.Nx, under.Ny, under.Hx, under.Hy
    rng = np.random.default_rng(54)
    bottom = rng.random((Ny, Nx))
    grid = ossg.ImmersedBoundaryGrid(under, ossg.GridFittedBottom(bottom))
    h = grid.immersed_boundary.bottom_height
    assert tuple(h.data.shape) == (1, Ny + 2 * Hy, Nx + 2 * Hx) and h.loc[2] is None
    got = h.data.cpu().numpy()
    row = got[0, Hy + Ny - 1, Hx:Hx + Nx]
    assert np.array_equal(row, row[::-1])                                                # :54
    assert np.array_equal(row[:Nx // 2], bottom[Ny - 1, :Nx // 2])                       # left half kept
    want = np.zeros_like(got)
    want[0, Hy:Hy + Ny, Hx:Hx + Nx] = bottom
    want = oracle.fill_halo_regions(want, Nx, Ny, 1, Hx, Hy, 0, 0, 0, 1.0)
    assert _same_bits(got, want)
    # fields on the wrapped grid take the zipper like on the bare one (TRG union, src/tripolar_grid.jl:371)
    c, u = ossg.CenterField(grid), ossg.XFaceField(grid)
    assert c.boundary_conditions.north.condition == 1 and u.boundary_conditions.north.condition == -1
    ossg.set_(c, lambda lam, phi, z: lam + 0 * phi + 0 * z)
    ossg.fill_halo_regions(c)
    crow = c.interior()[0, Ny - 1, :].cpu().numpy()
    assert np.array_equal(crow, crow[::-1])                                              # :65
    # bickley-jet mask: both poles and the southern cap are immersed, and the mask is fold-symmetric
    lp, pp = 45.0, 25.0
    under = ossg.TripolarGrid(size=(120, 60, 1), halo=(5, 5, 5), first_pole_longitude=lp, north_poles_latitude=pp)
    def mask(lam, phi):
        near = lambda l0: ((lam - l0).abs() < 5) & ((pp - phi).abs() < 5)
        return (near(lp) | near(lp + 180) | (phi < -78)).to(lam.dtype)
    grid = ossg.ImmersedBoundaryGrid(under, ossg.GridFittedBottom(mask))
    hb = grid.immersed_boundary.bottom_height.interior()[0].cpu().numpy()
    assert np.array_equal(hb[-1], hb[-1][::-1])
    assert hb[-1, 0] == 1 and hb[-1, 60] == 1 and hb[0].all() and 0 < hb.mean() < 0.2
    wet = ~ossg.immersed_cell(grid)
    assert tuple(wet.shape) == (1, 60 + 10, 120 + 10) and bool(wet[0, 5 + 30, 5 + 30])


def test_fold_quirk_sign_flip_is_reproduced(ossg, oracle):
    """for (Face, Center), sign -1, c[Nx/2+1, Ny] flips on EVERY fill
    (src/zipper_boundary_condition.jl:102 with i' == i); a drop-in must do the same."""
    dims = (12, 9, 1, 2, 2, 1)
    Nx, Ny, Nz, Hx, Hy, Hz = dims
    host = _poisoned_field((Nz + 2 * Hz, Ny + 2 * Hy, Nx + 2 * Hx), *dims, np.float64, seed=11)
    dev = torch.as_tensor(host).cuda()
    d = _desc(ossg, dev, dims, (1, 0), -1)
    want = host.copy()
    for _ in range(3):
        ossg._lib.check(ossg._lib.lib().ossg_zipper_fill(C.byref(d), None))
        oracle.zipper(want, Nx, Ny, Nz, Hx, Hy, Hz, 1, 0, -1)
    torch.cuda.synchronize()
    assert _same_bits(dev.cpu().numpy(), want)
    v0 = host[Hz, Ny + Hy - 1, Nx // 2 + Hx]
    assert dev.cpu().numpy()[Hz, Ny + Hy - 1, Nx // 2 + Hx] == -v0


def test_fold_pack_unpack_equals_zipper(ossg, oracle):
    """An (x)-partitioned top row assembled from pack/unpack messages equals the serial zipper."""
    from ossg_b200.distributed import fold_plan, local_sizes
    gNx, Ny, Nz, Hx, Hy, Hz = 24, 10, 2, 3, 3, 1
    for Rx in (1, 2, 3, 4):
        for locname, loc in LOCS.items():
            for sign in (1, -1):
                host = _poisoned_field((Nz + 2 * Hz, Ny + 2 * Hy, gNx + 2 * Hx), gNx, Ny, Nz, Hx, Hy, Hz, np.float64, seed=5)
                want = oracle.zipper(host.copy(), gNx, Ny, Nz, Hx, Hy, Hz, loc[0], loc[1], sign)
                sx = local_sizes(gNx, Rx)
                starts = [1 + sum(sx[:r]) for r in range(Rx)]
                tiles = []
                for r in range(Rx):      # local tiles: interior columns + (stale) x halos
                    t = np.ascontiguousarray(host[:, :, starts[r] - 1:starts[r] - 1 + sx[r] + 2 * Hx])
                    tiles.append(torch.as_tensor(t).cuda())
                L = ossg._lib.lib()
                for msg in fold_plan(gNx, Rx, 1, loc[0]):
                    src, dst = msg.src, msg.dst
                    fs = _desc(ossg, tiles[src], (sx[src], Ny, Nz, Hx, Hy, Hz), loc, sign)
                    fd = _desc(ossg, tiles[dst], (sx[dst], Ny, Nz, Hx, Hy, Hz), loc, sign)
                    ws = ossg._lib.FoldWindow(gNx, starts[src], sx[src], msg.src_i0, msg.n)
                    wd = ossg._lib.FoldWindow(gNx, starts[dst], sx[dst], msg.src_i0, msg.n)
                    ne = L.ossg_fold_buffer_elems(C.byref(fs), C.byref(ws))
                    buf = torch.empty(ne, dtype=torch.float64, device="cuda")
                    ossg._lib.check(L.ossg_fold_pack(C.byref(fs), C.byref(ws), buf.data_ptr(), None))
                    ossg._lib.check(L.ossg_fold_unpack(C.byref(fd), C.byref(wd), buf.data_ptr(), None))
                torch.cuda.synchronize()
                for r in range(Rx):
                    got = tiles[r].cpu().numpy()[:, :, Hx:Hx + sx[r]]
                    exp = want[:, :, starts[r] - 1 + Hx:starts[r] - 1 + Hx + sx[r]]
                    assert _same_bits(np.ascontiguousarray(got), np.ascontiguousarray(exp)), (Rx, locname, sign, r)


def test_peer_fold_kernel_equals_zipper(ossg, oracle):
    """ossg_zipper_fill_peer (the one-kernel NVLink fold) with the top-row "ranks" emulated on one
    GPU: every rank's tile is its own array, peer pointers are plain device pointers and each
    rank's kernel runs on its own stream (they wait for each other through the epoch flags).
    Two epochs, all four staggerings in one batch, Rx = 1..4, both signs, FP64 and FP32, pull
    (peer loads), push (peer stores) and bulk push (rows staged in shared memory, cp.async.bulk row
    segments) forms.  The bulk form also runs on wider grids whose chunks of 256 columns are full,
    straddle ranks, have odd row pitches (the 16-byte alignment of a segment then changes from row to
    row) and need two staging passes (Hy = 6)."""
    from ossg_b200.distributed import local_sizes
    lib_ = ossg._lib
    L = lib_.lib()
    Ny, Hx, Hz = 10, 3, 1
    # (dtype, form, global Nx, Hy, widest); widest: bulk_waves < 0, i.e. chunks of up to 544
    # columns (one chunk per row of a 540-column tile: three columns per thread) instead of the
    # load-balanced narrower ones
    Nz = 2
    cases = [(np.float64, 0, 24, 3, 0), (np.float32, 0, 24, 3, 0), (np.float64, 1, 24, 3, 0), (np.float32, 1, 24, 3, 0),
             (np.float64, 2, 24, 3, 0), (np.float32, 2, 24, 3, 0), (np.float64, 2, 1100, 3, 0), (np.float32, 2, 1100, 6, 0),
             (np.float64, 2, 1030, 6, 0), (np.float32, 2, 778, 3, 0), (np.float64, 2, 1080, 4, 1), (np.float32, 2, 1086, 3, 1),
             (np.float64, 2, 1100, 6, 1)]
    for npdt, push, gNx, Hy, widest in cases:
        for Rx in (1, 2, 3, 4) + ((7, 8) if gNx > 24 else ()):
            sx = local_sizes(gNx, Rx)
            starts = [1 + sum(sx[:r]) for r in range(Rx)]
            specs = [(loc, sign) for loc in LOCS.values() for sign in (1, -1)]
            hosts = [_poisoned_field((Nz + 2 * Hz, Ny + 2 * Hy, gNx + 2 * Hx), gNx, Ny, Nz, Hx, Hy, Hz, npdt, seed=11 + k)
                     for k in range(len(specs))]
            tiles = [[torch.as_tensor(np.ascontiguousarray(h[:, :, starts[r] - 1:starts[r] - 1 + sx[r] + 2 * Hx])).cuda()
                      for r in range(Rx)] for h in hosts]
            flags = [torch.zeros(2 * lib_.OSSG_MAX_PEERS + 1, dtype=torch.int64, device="cuda") for _ in range(Rx)]
            streams = [torch.cuda.Stream() for _ in range(Rx)]
            torch.cuda.synchronize()
            wants = [h.copy() for h in hosts]
            for epoch in (1, 2):
                for r in range(Rx):
                    fd = lib_.PeerFoldDesc()
                    fd.gNx, fd.me, fd.nranks, fd.epoch, fd.push = gNx, r, Rx, epoch, push
                    fd.bulk_waves = -1 if widest else 0
                    for q in range(Rx):
                        fd.i0[q], fd.nx[q] = starts[q], sx[q]
                        fd.flags_peer[q] = None if q == r else flags[q].data_ptr()
                    fd.flags_local = flags[r].data_ptr()
                    arr = (lib_.PeerField * len(specs))()
                    for k, (loc, sign) in enumerate(specs):
                        arr[k].f = _desc(ossg, tiles[k][r], (sx[r], Ny, Nz, Hx, Hy, Hz), loc, sign)
                        for q in range(Rx):
                            arr[k].peer[q] = tiles[k][q].data_ptr()
                    lib_.check(L.ossg_zipper_fill_peer(arr, len(specs), C.byref(fd), streams[r].cuda_stream))
                torch.cuda.synchronize()
                for k, (loc, sign) in enumerate(specs):
                    oracle.zipper(wants[k], gNx, Ny, Nz, Hx, Hy, Hz, loc[0], loc[1], sign)
                    for r in range(Rx):
                        got = tiles[k][r].cpu().numpy()[:, :, Hx:Hx + sx[r]]
                        exp = wants[k][:, :, starts[r] - 1 + Hx:starts[r] - 1 + Hx + sx[r]]
                        assert _same_bits(np.ascontiguousarray(got), np.ascontiguousarray(exp)), (npdt, push, gNx, Rx, loc, sign, r, epoch)


# ------------------------------------------------------------------ argument errors through the ABI
def test_abi_argument_errors(ossg):
    from ossg_b200.grids import make_params
    L = ossg._lib.lib()
    p = make_params((61, 30, 1), (4, 4, 4), -80, 55, 70, orc.R_EARTH)
    out = ossg._lib.ptr_array([8] * 20)
    assert L.ossg_tripolar_build(C.byref(p), 0, out, 16, 1 << 20, None) == ossg._lib.OSSG_EODD_NX
    assert b"should be even" in L.ossg_last_error()
    p = make_params((60, 30, 1), (4, 30, 4), -80, 55, 70, orc.R_EARTH)
    assert L.ossg_tripolar_build(C.byref(p), 0, out, 16, 1 << 20, None) == ossg._lib.OSSG_EHALO
    with pytest.raises(ValueError, match="should be even"):
        ossg.TripolarGrid(size=(61, 30, 1))


def test_headline_grid_matches_oracle_bit_for_bit(ossg, oracle):
    """The benchmark configuration itself (1/12 degree, 4320 x 2160, halo 4, FP64): all 20 padded
    arrays against the multi-threaded oracle, array by array to bound host memory."""
    size, halo = (4320, 2160, 100), (4, 4, 4)
    oracle.set_threads(oracle.max_threads())
    ref = oracle.build(orc.make_params(size, halo))
    grid = ossg.TripolarGrid(ossg.GPU(), ossg.Float64, size=size, halo=halo)
    for name in orc.ALL_NAMES:
        got = grid.arrays[name].parent.cpu().numpy()
        assert np.array_equal(got.view(np.uint64), ref[name].view(np.uint64)), name
        ref[name] = None


def test_cartesian_nodes_and_orthogonality_on_gpu(ossg, oracle):
    """ossg_cartesian_nodes (Oceananigans lat_lon_to_cartesian / get_cartesian_nodes_and_vertices)
    bit for bit against the oracle, and the reference's orthogonality test
    (test/test_tripolar_grid.jl:36-76) evaluated on the GPU-produced FF nodes of the 1 degree grid
    with poles at (75, 35 N): centred-difference angle within 0.01 degree of 90 outside the masked
    pole windows and phi < -78 (bounds as in tests/test_oracle_golden.py)."""
    size = (360, 180, 1)
    grid = ossg.TripolarGrid(ossg.GPU(), size=size, first_pole_longitude=75, north_poles_latitude=35)
    for loc, suffix in (((ossg.Face, ossg.Face), "ff"), ((ossg.Center, ossg.Center), "cc")):
        x, y, z = ossg.cartesian_nodes(grid, loc, radius=1.0)
        lam, phi = grid.arrays["lam_" + suffix].parent.cpu().numpy(), grid.arrays["phi_" + suffix].parent.cpu().numpy()
        for got, want in zip((x, y, z), oracle.cartesian_nodes(lam, phi, 1.0)):
            assert np.array_equal(got.parent.cpu().numpy().view(np.uint64), want.view(np.uint64)), suffix
    xr = ossg.cartesian_nodes(grid, (ossg.Face, ossg.Face))[0]           # default radius = grid.radius
    want = oracle.cartesian_nodes(grid.arrays["lam_ff"].parent.cpu().numpy(), grid.arrays["phi_ff"].parent.cpu().numpy(), grid.radius)[0]
    assert np.array_equal(xr.parent.cpu().numpy().view(np.uint64), want.view(np.uint64))
    x32 = ossg.cartesian_nodes(grid, (ossg.Face, ossg.Face), dtype=torch.float32)[0]
    assert np.array_equal(x32.parent.cpu().numpy(), want.astype(np.float32))
    # orthogonality on the device
    Nx, Ny, H = 360, 180, 4
    x, y, z = (t.parent[H:H + Ny, H:H + Nx] for t in ossg.cartesian_nodes(grid, (ossg.Face, ossg.Face), radius=1.0))
    P = torch.stack([x, y, z], dim=-1)
    lamd = grid.arrays["lam_ff"].parent[H:H + Ny, H:H + Nx]
    phid = grid.arrays["phi_ff"].parent[H:H + Ny, H:H + Nx]
    masked = (((lamd - 75).abs() < 5) & ((35 - phid).abs() < 5)) | (((lamd - 255).abs() < 5) & ((35 - phid).abs() < 5)) | (phid < -78)
    v1, v2 = P[1:-1, 2:] - P[1:-1, :-2], P[2:, 1:-1] - P[:-2, 1:-1]
    cos = (v1 * v2).sum(-1) / (v1.norm(dim=-1) * v2.norm(dim=-1))
    ang = torch.rad2deg(torch.arccos(cos.clamp(-1, 1))) - 90
    assert float(ang[~masked[1:-1, 1:-1]].abs().max()) < 0.01


def test_mirror_pair_tile_kernel_matches_oracle():
    """The opt-in mirror-pair form of the tile kernel (OSSG_TILE_PAIRS=1: clusters of two CTAs that
    share the latitude part of the point stage through distributed shared memory) in a fresh process:
    FP64 and FP32 grids incl. one whose tiles are all partial (Nx/4 < 32), a Y-slab window, and the
    staged metrics entry point fed with glibc-libm coordinates, whose latitudes are NOT mirror
    symmetric bit for bit -- every CTA must then notice and fall back to the ordinary point stage."""
    import os
    import subprocess
    import sys
    code = r"""
import ctypes as C, sys, numpy as np, torch
sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests')
import ossg_b200 as ossg
from oracle import oracle as orc
from ossg_b200.grids import build_arrays, make_params
o = orc.Oracle(); o.set_threads(o.max_threads())
for size, dt, npdt in (((360, 180, 1), ossg.Float64, np.float64), ((1440, 720, 1), ossg.Float64, np.float64),
                       ((60, 30, 1), ossg.Float64, np.float64), ((1000, 300, 1), ossg.Float64, np.float64),
                       ((360, 180, 1), ossg.Float32, np.float32)):
    g = ossg.TripolarGrid(ossg.GPU(), dt, size=size)
    ref = o.build(orc.make_params(size, (4, 4, 4)), dtype=npdt)
    for n in orc.ALL_NAMES:
        assert np.array_equal(g.arrays[n].parent.cpu().numpy().view(np.uint8), ref[n].view(np.uint8)), (size, n)
ref = o.build(orc.make_params((1440, 720, 1), (4, 4, 4)))
p = make_params((1440, 720, 1), (4, 4, 4), -80, 55, 70, orc.R_EARTH, tile=(1, 361, 1440, 360))
for n, t in zip(orc.ALL_NAMES, build_arrays(p, torch.float64, torch.device('cuda', 0))):
    assert np.array_equal(t.cpu().numpy().view(np.uint8), np.ascontiguousarray(ref[n][360:360 + 368, :]).view(np.uint8)), n
# staged metrics from libm coordinates (a different last-ulp rounding of every coordinate)
ol = orc.Oracle(libm=True)
size = (360, 180, 1)
po = orc.make_params(size, (4, 4, 4))
coords = ol.coords(po)
want = o.metrics(po, coords)
L = ossg._lib.lib(); dev = torch.device('cuda', 0)
params = make_params(size, (4, 4, 4), -80, 55, 70, orc.R_EARTH)
cin = [torch.as_tensor(coords[n], device=dev).contiguous() for n in orc.COORD_NAMES]
out = [torch.empty_like(cin[0]) for _ in range(12)]
ws = torch.empty(L.ossg_tripolar_workspace_bytes(C.byref(params), 0), dtype=torch.uint8, device=dev)
ossg._lib.check(L.ossg_tripolar_metrics(C.byref(params), ossg._lib.ptr_array([t.data_ptr() for t in cin]), 0,
                                        ossg._lib.ptr_array([t.data_ptr() for t in out]), ws.data_ptr(), ws.numel(),
                                        torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
for n, t in zip(orc.METRIC_NAMES, out):
    assert np.array_equal(t.cpu().numpy().view(np.uint64), want[n].view(np.uint64)), n
print('pairs ok')
""" % (ROOT, ROOT)
    env = dict(os.environ, OSSG_TILE_PAIRS="1")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0 and "pairs ok" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]


def test_tma_staging_of_the_tile_kernel_matches_oracle():
    """The opt-in TMA staging of metrics_tile_kernel (OSSG_TILE_STAGING=tma, UTMALDG box loads on an
    mbarrier) in a fresh process: FP64 and FP32 grids and a tile window, bit for bit."""
    import os
    import subprocess
    import sys
    code = r"""
import sys, numpy as np, torch
sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests')
import ossg_b200 as ossg
from oracle import oracle as orc
from ossg_b200.grids import build_arrays, make_params
o = orc.Oracle(); o.set_threads(o.max_threads())
for size, dt, npdt in (((360, 180, 1), ossg.Float64, np.float64), ((1440, 720, 1), ossg.Float64, np.float64), ((360, 180, 1), ossg.Float32, np.float32)):
    g = ossg.TripolarGrid(ossg.GPU(), dt, size=size)
    ref = o.build(orc.make_params(size, (4, 4, 4)), dtype=npdt)
    for n in orc.ALL_NAMES:
        assert np.array_equal(g.arrays[n].parent.cpu().numpy().view(np.uint8), ref[n].view(np.uint8)), (size, n)
ref = o.build(orc.make_params((1440, 720, 1), (4, 4, 4)))
p = make_params((1440, 720, 1), (4, 4, 4), -80, 55, 70, orc.R_EARTH, tile=(1, 361, 1440, 360))
for n, t in zip(orc.ALL_NAMES, build_arrays(p, torch.float64, torch.device('cuda', 0))):
    assert np.array_equal(t.cpu().numpy().view(np.uint8), np.ascontiguousarray(ref[n][360:360 + 368, :]).view(np.uint8)), n
print('tma ok')
""" % (ROOT, ROOT)
    env = dict(os.environ, OSSG_TILE_STAGING="tma")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0 and "tma ok" in res.stdout, res.stdout[-1500:] + res.stderr[-1500:]


def test_branch_free_sqrt_div_match_ieee(ossg):
    """sqrt_nb / div_nb (the tile kernel's straight-line path) against the device's IEEE sqrt and
    division on 2^28 pseudo-random operands from the ranges the kernel admits."""
    from ossg_b200 import _lib
    bad = (C.c_uint64 * 2)()
    _lib.check_util(_lib.testutil().ossg_selftest_sqrt_div(1 << 28, 0x5EED, bad, None))
    assert (bad[0], bad[1]) == (0, 0)


# ------------------------------------------------------------------ full-size properties (no oracle)
@pytest.mark.parametrize("size", [(1440, 720, 75), (4320, 2160, 100)])
def test_full_size_structural_invariants(ossg, size):
    """Size-independent properties of the constructor at BASELINE sizes C and D
    (SURVEY section 8c 'structural invariants')."""
    grid = ossg.TripolarGrid(ossg.GPU(), size=size)
    Nx, Ny, _ = size
    H = 4
    a = {n: grid.arrays[n] for n in orc.ALL_NAMES}
    for n in orc.ALL_NAMES:
        p = a[n].parent
        assert bool(torch.isfinite(p).all()), n
        # periodic x halos over all rows
        assert torch.equal(p[:, :H], p[:, Nx:Nx + H]) and torch.equal(p[:, Nx + H:], p[:, H:2 * H]), n
    for n in ("lam_cc", "phi_cc", "dx_cc", "dy_cc", "az_cc"):        # x-Center, y-Center: row Ny mirror symmetric
        row = a[n][slice(1, Nx), Ny]
        assert torch.equal(row, row.flip(0)), n
    for n in ("lam_fc", "phi_fc", "dx_fc", "dy_fc", "az_fc"):        # x-Face: i <-> Nx - i + 2
        row = a[n][slice(2, Nx), Ny]
        assert torch.equal(row, row.flip(0)), n
    for n in ("dx_fc", "dy_fc", "az_fc"):                            # exact zeros at the two pole points
        assert float(a[n][1, Ny]) == 0.0 and float(a[n][Nx // 2 + 1, Ny]) == 0.0, n
    for n in orc.COORD_NAMES:                                        # south halo of coordinates stays 0 (south = nothing)
        assert float(a[n].parent[:H, :].abs().max()) == 0.0, n
    lam = a["lam_cc"].interior(Nx, Ny)
    phi = a["phi_cc"].interior(Nx, Ny)
    assert float(lam.min()) >= 0 and float(lam.max()) <= 360 and float(phi.max()) <= 90      # runtests.jl:32-34
    # north halo = fold image: lam_cc[i, Ny + j] == lam_cc[Nx - i + 1, Ny - j]
    for j in range(1, H + 1):
        assert torch.equal(a["lam_cc"][slice(1, Nx), Ny + j], a["lam_cc"][slice(1, Nx), Ny - j].flip(0))
        assert torch.equal(a["az_cf"][slice(1, Nx), Ny + j], a["az_cf"][slice(1, Nx), Ny - j + 1].flip(0))
    # metrics positive in the interior above the south continuation row
    for n in ("dx_cc", "dy_cc", "az_cc", "dx_ff", "dy_ff", "az_ff"):
        assert float(a[n][slice(1, Nx), slice(2, Ny - 1)].min()) > 0, n


def test_full_size_fill_round_trip(ossg):
    """Fold applied twice is the identity on the halo (fold o fold = id for sign^2 = 1), except
    the reference's (Face, Center) pole quirk; checked on a D-sized slab of 4 levels."""
    Nx, Ny, Nz, H = 4320, 2160, 4, 4
    L = ossg._lib.lib()
    for loc, sign in (((0, 0), 1), ((1, 0), -1), ((0, 1), -1), ((1, 1), 1)):
        t = torch.empty((Nz + 2 * H, Ny + 2 * H, Nx + 2 * H), dtype=torch.float64, device="cuda")
        d = _desc(ossg, t, (Nx, Ny, Nz, H, H, H), loc, sign)
        ossg._lib.check_util(ossg._lib.testutil().ossg_synth_field(C.byref(d), 0x5EED, 1, 1, None))
        before = t.clone()
        ossg._lib.check(L.ossg_fill_halo_regions_batch(C.byref(d), 1, None))
        once = t.clone()
        ossg._lib.check(L.ossg_fill_halo_regions_batch(C.byref(d), 1, None))
        torch.cuda.synchronize()
        k = slice(H, H + Nz)
        # interior below row Ny untouched
        assert torch.equal(before[k, H:H + Ny - 1, H:H + Nx], t[k, H:H + Ny - 1, H:H + Nx])
        # halos of interior levels are fully written (no NaN left) and stable under a second fill
        assert bool(torch.isfinite(once[k, H + Ny:, :]).all())
        assert torch.equal(once[k, H + Ny:, :], t[k, H + Ny:, :])
        # fold image: c[i, Ny+1] = s c[i', Ny - 1 (+1)]
        js = Ny - 1 if loc[1] == 0 else Ny
        src = before[k, H + js - 1, H:H + Nx]
        if loc[0] == 0:
            img = src.flip(-1)
            assert torch.equal(once[k, H + Ny, H:H + Nx], sign * img)
        else:
            img = src[:, 1:].flip(-1)                  # i = 2..Nx  <-  i' = Nx..2
            assert torch.equal(once[k, H + Ny, H + 1:H + Nx], sign * img)
            assert torch.equal(once[k, H + Ny, H], abs(sign) * src[:, 0])   # i = 1 wraps: |sign|


# ---------------------------------------------------------------- round 2: NaN payloads, Hy = 0, fill forms, table cache, with_halo
@pytest.mark.parametrize("npdt", [np.float64, np.float32])
@pytest.mark.parametrize("sign", [1, -1])
def test_fold_of_nan_sources_matches_the_cpu_multiplication(ossg, oracle, npdt, sign):
    """`sign * c` (src/zipper_boundary_condition.jl:81) is an IEEE multiplication by +-1: a NaN source
    comes back quieted with ITS OWN sign and payload on the CPUs the reference runs on.  Sources
    here contain quiet and signalling NaNs of both signs; the device result equals the oracle's
    (x86 mulsd / mulss) bit for bit."""
    dims = (20, 12, 2, 4, 4, 1)
    Nx, Ny, Nz, Hx, Hy, Hz = dims
    shape = (Nz + 2 * Hz, Ny + 2 * Hy, Nx + 2 * Hx)
    rng = np.random.default_rng(11)
    host = rng.standard_normal(shape).astype(npdt)
    if npdt == np.float64:
        pats = np.array([0x7ff8000000000001, 0xfff8dead0000beef, 0x7ff0000000000123, 0xfff00000000beef1], dtype=np.uint64)
        bits = host.view(np.uint64)
    else:
        pats = np.array([0x7fc00001, 0xffc0dead, 0x7f800123, 0xff80beef], dtype=np.uint32)
        bits = host.view(np.uint32)
    mask = rng.random(shape) < 0.2
    bits[mask] = pats[rng.integers(0, 4, size=int(mask.sum()))]
    for locname, loc in LOCS.items():
        dev = torch.as_tensor(host.copy()).cuda()
        d = _desc(ossg, dev, dims, loc, sign)
        ossg._lib.check(ossg._lib.lib().ossg_fill_halo_regions_batch(C.byref(d), 1, None))
        torch.cuda.synchronize()
        want = oracle.fill_halo_regions(host.copy(), Nx, Ny, Nz, Hx, Hy, Hz, loc[0], loc[1], sign)
        assert _same_bits(dev.cpu().numpy(), want), locname


def test_zero_y_halo_still_overwrites_the_fold_row(ossg, oracle):
    """Hy = 0: the j-loop of fold_north_* is empty, but the row-Ny overwrite of y-Center fields
    (:102, :135) does not depend on Hy; alone and inside a batch with periodic fill."""
    Nx, Ny, Nz, Hx, Hz = 16, 9, 2, 3, 1
    rng = np.random.default_rng(5)
    hosts, devs, descs = [], [], []
    for (loc, Hy, sign) in (((0, 0), 0, 1), ((1, 0), 0, -1), ((0, 1), 0, -1), ((0, 0), 2, 1)):
        dims = (Nx, Ny, Nz, Hx, Hy, Hz)
        h = rng.standard_normal((Nz + 2 * Hz, Ny + 2 * Hy, Nx + 2 * Hx))
        hosts.append((h, dims, loc, sign))
        devs.append(torch.as_tensor(h.copy()).cuda())
        descs.append(_desc(ossg, devs[-1], dims, loc, sign))
    arr = (ossg._lib.FieldDesc * len(descs))(*descs)
    ossg._lib.check(ossg._lib.lib().ossg_fill_halo_regions_batch(arr, len(descs), None))
    torch.cuda.synchronize()
    for (h, dims, loc, sign), dv in zip(hosts, devs):
        want = oracle.fill_halo_regions(h.copy(), *dims, loc[0], loc[1], sign)
        assert _same_bits(dv.cpu().numpy(), want), (dims, loc)


def test_fill_halo_regions_piece_paths_are_bit_identical(ossg, oracle):
    """The fused fill's periodic part: rows that are 32-byte aligned take the 256-bit piece path (one
    thread per row, L2::64B loads), other layouts the general per-(row, side) path; both against the oracle,
    four fields per launch."""
    L = ossg._lib.lib()
    for dims in ((60, 30, 3, 4, 4, 2), (62, 21, 5, 4, 3, 1), (34, 17, 2, 3, 9, 1), (64, 12, 2, 4, 2, 0)):
        Nx, Ny, Nz, Hx, Hy, Hz = dims
        shape = (Nz + 2 * Hz, Ny + 2 * Hy, Nx + 2 * Hx)
        fields = []
        for q, (locname, loc) in enumerate(LOCS.items()):
            sign = -1 if loc[0] != loc[1] else 1
            host = _poisoned_field(shape, Nx, Ny, Nz, Hx, Hy, Hz, np.float64, seed=31 + q)
            dev = torch.as_tensor(host).cuda()
            fields.append((host, dev, loc, sign, _desc(ossg, dev, dims, loc, sign)))
        arr = (ossg._lib.FieldDesc * 4)(*[f[4] for f in fields])
        ossg._lib.check(L.ossg_fill_halo_regions_batch(arr, 4, None))
        torch.cuda.synchronize()
        for host, dev, loc, sign, _ in fields:
            want = oracle.fill_halo_regions(host.copy(), Nx, Ny, Nz, Hx, Hy, Hz, loc[0], loc[1], sign)
            assert _same_bits(dev.cpu().numpy(), want), (dims, loc)


def test_table_cache_in_the_workspace(ossg, oracle):
    """The workspace header stamps the 1-D tables: a second build on the same workspace skips the
    table computation, other latitudes / sizes on the same workspace recompute; results equal the
    oracle every time."""
    from ossg_b200.grids import make_params
    L = ossg._lib.lib()
    dev = torch.device("cuda", 0)
    ws = torch.empty(1 << 20, dtype=torch.uint8, device=dev)
    ws.fill_(0xAB)                                                  # garbage, then the init the contract asks for
    ossg._lib.check(L.ossg_tripolar_workspace_init(ws.data_ptr(), ws.numel(), None))
    cases = [((60, 30, 1), (4, 4, 4), {}), ((60, 30, 1), (4, 4, 4), {}), ((60, 30, 1), (4, 7, 4), {}),
             ((60, 30, 1), (4, 4, 4), dict(north_poles_latitude=45)), ((40, 34, 1), (5, 31, 5), dict(southernmost_latitude=-70)),
             ((60, 30, 1), (4, 4, 4), dict(radius=1.0)), ((60, 30, 1), (4, 4, 4), {})]
    for size, halo, kw in cases:
        p = make_params(size, halo, kw.get("southernmost_latitude", -80), kw.get("north_poles_latitude", 55),
                        kw.get("first_pole_longitude", 70), kw.get("radius", orc.R_EARTH))
        shape = (size[1] + 2 * halo[1], size[0] + 2 * halo[0])
        out = [torch.full(shape, float("nan"), dtype=torch.float64, device=dev) for _ in range(20)]
        ossg._lib.check(L.ossg_tripolar_build(C.byref(p), 0, ossg._lib.ptr_array([t.data_ptr() for t in out]), ws.data_ptr(), ws.numel(), None))
        torch.cuda.synchronize()
        ref = oracle.build(orc.make_params(size, halo, **kw))
        got = {n: t.cpu().numpy() for n, t in zip(orc.ALL_NAMES, out)}
        assert assert_grid_matches(got, ref) == 0, (size, halo, kw)
    # the stamp is in place: words 0 and 5 of the header carry the magic numbers
    hdr = ws[:128].cpu().numpy().view(np.uint64)
    assert hdr[0] == 0x4f5353475f58595f and hdr[5] == 0x4f5353475f4c4c5f and hdr[1] == 60 and hdr[2] == 30


def test_with_halo_on_device_reuses_the_tables(ossg, oracle):
    """src/with_halo.jl:5-23, test/runtests.jl:52-71: the grid rebuilt with a larger halo equals the
    oracle's grid with that halo (the 1-D tables travel with the grid), Float64 and Float32."""
    for FT in (ossg.Float64, ossg.Float32):
        grid = ossg.TripolarGrid(ossg.GPU(), FT, size=(40, 34, 2), halo=(4, 4, 4), north_poles_latitude=45, first_pole_longitude=35)
        assert grid.tables is not None
        big = ossg.with_halo((5, 31, 5), grid)
        assert big.halo == (5, 31, 5) and big.size == grid.size and big.tables is not None
        kw = dict(north_poles_latitude=45, first_pole_longitude=35)
        ref = oracle.build(orc.make_params((40, 34, 2), (5, 31, 5), **kw))
        got = to_np(big)
        if FT is ossg.Float32:
            for n in orc.ALL_NAMES:
                assert np.array_equal(got[n].view(np.uint32), ref[n].astype(np.float32).view(np.uint32)), n
        else:
            assert assert_grid_matches(got, ref) == 0
        small = ossg.with_halo((2, 2, 2), big)
        ref2 = oracle.build(orc.make_params((40, 34, 2), (2, 2, 2), **kw))
        got2 = to_np(small)
        for n in orc.ALL_NAMES:
            want = ref2[n] if FT is ossg.Float64 else ref2[n].astype(np.float32)
            assert np.array_equal(got2[n], want, equal_nan=True), n


def test_two_streams_build_concurrently(ossg, oracle):
    """Re-entrancy: builds on two streams, each with its own workspace, share nothing (every launch of a
    build goes to the caller's stream); both results are the oracle's."""
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    from ossg_b200.grids import build_arrays, make_params
    dev = torch.device("cuda", 0)
    pa = make_params((360, 180, 1), (4, 4, 4), -80, 55, 70, orc.R_EARTH)
    pb = make_params((120, 60, 1), (5, 5, 5), -80, 45, 25, orc.R_EARTH)
    outs = []
    for it in range(3):
        a = build_arrays(pa, ossg.Float64, dev, stream=s1)
        b = build_arrays(pb, ossg.Float32, dev, stream=s2)
        outs = [a, b]
    torch.cuda.synchronize()
    ra = oracle.build(orc.make_params((360, 180, 1), (4, 4, 4)))
    rb = oracle.build(orc.make_params((120, 60, 1), (5, 5, 5), north_poles_latitude=45, first_pole_longitude=25))
    assert assert_grid_matches({n: t.cpu().numpy() for n, t in zip(orc.ALL_NAMES, outs[0])}, ra) == 0
    for n, t in zip(orc.ALL_NAMES, outs[1]):
        assert np.array_equal(t.cpu().numpy().view(np.uint32), rb[n].astype(np.float32).view(np.uint32)), n


def test_back_to_back_builds_over_poisoned_outputs(ossg, oracle):
    """Launch chain of the constructor (programmatic dependent launch; the tile blocks wait for the ten coordinate
    rows of their patches, csrc/ossg_tripolar.cu row counters): grids with different poles are built back to back
    into the SAME arrays, which are poisoned and read in between, with no host synchronisation inside the loop.
    A tile that started before its rows were complete, or that was served a stale cache line of the previous
    grid, shows as a NaN or a wrong bit.  Whole grid and an x-split tile (the per-row weights differ)."""
    from ossg_b200.grids import build_arrays, make_params
    dev = torch.device("cuda", 0)
    size, halo = (1440, 720, 1), (4, 4, 4)
    kws = [dict(north_poles_latitude=55, first_pole_longitude=70), dict(north_poles_latitude=62, first_pole_longitude=15)]
    refs = [oracle.build(orc.make_params(size, halo, **kw)) for kw in kws]
    for tile in (None, (721, 1, 720, 720)):
        ny, nx = (size[1], size[0]) if tile is None else (tile[3], tile[2])
        out = [torch.empty((ny + 2 * halo[1], nx + 2 * halo[0]), dtype=torch.float64, device=dev) for _ in range(20)]
        got = []
        for it in range(6):
            kw = kws[it % 2]
            for t in out:
                t.fill_(float("nan"))
            acc = sum(t[::7, ::5].sum() for t in out)          # a read pass over the poisoned arrays
            p = make_params(size, halo, -80, kw["north_poles_latitude"], kw["first_pole_longitude"], orc.R_EARTH, tile=tile)
            build_arrays(p, ossg.Float64, dev, out=out)
            if it >= 4:
                got.append([t.clone() for t in out])
        torch.cuda.synchronize()
        assert bool(torch.isnan(acc))
        for k, arrays in enumerate(got):
            ref = refs[k % 2]
            for n, t in zip(orc.ALL_NAMES, arrays):
                want = ref[n] if tile is None else ref[n][tile[1] - 1:tile[1] - 1 + ny + 2 * halo[1], tile[0] - 1:tile[0] - 1 + nx + 2 * halo[0]]
                assert np.array_equal(t.cpu().numpy().view(np.uint64), np.ascontiguousarray(want).view(np.uint64)), (tile, k, n)


def test_build_is_capturable_in_a_cuda_graph(ossg, oracle):
    """The constructor enqueues three kernels on the caller's stream and nothing else (no allocation, no
    synchronisation, no host wait), chained with programmatic dependent launch: it can be captured into a CUDA graph
    and replayed.  Every replay (outputs poisoned in between) gives the oracle's bits."""
    from ossg_b200.grids import make_params
    L = ossg._lib.lib()
    dev = torch.device("cuda", torch.cuda.current_device())
    size, halo = (360, 180, 1), (4, 4, 4)
    p = make_params(size, halo, -80, 55, 70, orc.R_EARTH)
    out = [torch.empty((size[1] + 2 * halo[1], size[0] + 2 * halo[0]), dtype=torch.float64, device=dev) for _ in range(20)]
    optr = ossg._lib.ptr_array([t.data_ptr() for t in out])
    ws = torch.empty(L.ossg_tripolar_workspace_bytes(C.byref(p), ossg._lib.OSSG_F64), dtype=torch.uint8, device=dev)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        ossg._lib.check(L.ossg_tripolar_workspace_init(ws.data_ptr(), ws.numel(), side.cuda_stream))
        ossg._lib.check(L.ossg_tripolar_build(C.byref(p), ossg._lib.OSSG_F64, optr, ws.data_ptr(), ws.numel(), side.cuda_stream))   # warm-up (module load, attributes)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        ossg._lib.check(L.ossg_tripolar_build(C.byref(p), ossg._lib.OSSG_F64, optr, ws.data_ptr(), ws.numel(),
                                              torch.cuda.current_stream().cuda_stream))
    ref = oracle.build(orc.make_params(size, halo))
    for rep in range(3):
        for t in out:
            t.fill_(float("nan"))
        graph.replay()
        torch.cuda.synchronize()
        for n, t in zip(orc.ALL_NAMES, out):
            assert np.array_equal(t.cpu().numpy().view(np.uint64), ref[n].view(np.uint64)), (rep, n)


@pytest.mark.parametrize("knob", ["OSSG_ROW_DEPS=1", "OSSG_FRAME=fork", "OSSG_FRAME=tail"])
def test_launch_options_in_a_child_process(ossg, knob):
    """The launch-form knobs are read once per process, so each runs in a child: OSSG_ROW_DEPS=1 (tile blocks wait on
    per-row counters instead of the whole coordinate grid), OSSG_FRAME=fork (halo frame on a side stream with an event
    pair) and OSSG_FRAME=tail (its own grid behind the tile kernel, programmatic launch).  Same parity tests, same bits."""
    import os, subprocess, sys
    if os.environ.get("OSSG_TEST_CHILD"):
        pytest.skip("child process")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    name, value = knob.split("=")
    env = dict(os.environ, OSSG_TEST_CHILD="1")
    env[name] = value
    res = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", os.path.join(root, "tests", "test_gpu_parity.py"),
                          "-k", "back_to_back_builds or construction_matches_oracle or tiles_equal_slices"], env=env, capture_output=True, text=True, timeout=600, cwd=root)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]


def test_missing_partner_traps_instead_of_passing_silently(ossg):
    """A rank whose partner never makes the matching call must fail loudly: the
    waiting kernel traps after OSSG_PEER_TIMEOUT_MS and the next CUDA call of that process raises.
    Runs in a child process (a trap takes the CUDA context with it)."""
    import subprocess, sys
    code = r'''
import ctypes as C, sys
sys.path.insert(0, %r)
import torch, ossg_b200
from ossg_b200 import _lib
L = _lib.lib()
Nx, Ny, Nz, H = 24, 10, 2, 3
tiles = [torch.zeros((Nz + 2, Ny + 2 * H, Nx // 2 + 2 * H), dtype=torch.float64, device="cuda") for _ in range(2)]
flags = [torch.zeros(17, dtype=torch.int64, device="cuda") for _ in range(2)]
fd = _lib.PeerFoldDesc()
fd.gNx, fd.me, fd.nranks, fd.epoch, fd.push = Nx, 0, 2, 1, 1
for q in range(2):
    fd.i0[q], fd.nx[q] = 1 + q * (Nx // 2), Nx // 2
fd.flags_peer[1] = flags[1].data_ptr()
fd.flags_local = flags[0].data_ptr()
arr = (_lib.PeerField * 1)()
arr[0].f = _lib.FieldDesc(tiles[0].data_ptr(), 0, Nx // 2, Ny, Nz, H, H, 1, 0, 0, 0, 0, 1.0)
arr[0].peer[0], arr[0].peer[1] = tiles[0].data_ptr(), tiles[1].data_ptr()
_lib.check(L.ossg_zipper_fill_peer(arr, 1, C.byref(fd), None))      # rank 1 never calls
try:
    torch.cuda.synchronize()
except Exception as e:
    print("RAISED", type(e).__name__)
    sys.exit(7)
print("PASSED SILENTLY")
''' % ROOT
    env = dict(_os.environ, OSSG_PEER_TIMEOUT_MS="300")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120, env=env)
    assert "PASSED SILENTLY" not in r.stdout and r.returncode != 0, (r.returncode, r.stdout[-300:], r.stderr[-300:])


def test_float32_longitude_reading_matches_the_oracle_option(ossg, oracle):
    """OSSG_LAMBDA1D_F32 / float32_longitudes=True: the 1-D longitudes pre-rounded to Float32 (one reading of
    generate_coordinate(FT, ...), src/tripolar_grid.jl:89); device and oracle agree bit for bit under either
    reading, and the two readings differ (the option is not a no-op)."""
    size, halo = (100, 60, 1), (4, 4, 4)           # 3.6 degree steps: not representable in Float32
    try:
        oracle.set_lambda_f32(True)
        ref = oracle.build(orc.make_params(size, halo))
    finally:
        oracle.set_lambda_f32(False)
    ref0 = oracle.build(orc.make_params(size, halo))
    for FT, cast in ((ossg.Float64, np.float64), (ossg.Float32, np.float32)):
        g = ossg.TripolarGrid(ossg.GPU(), FT, size=size, halo=halo, float32_longitudes=True)
        got = to_np(g)
        for n in orc.ALL_NAMES:
            assert np.array_equal(got[n], ref[n].astype(cast), equal_nan=True), (FT, n)
    assert not np.array_equal(ref["lam_cc"], ref0["lam_cc"])
    inner = (slice(5, -5), slice(5, -5))
    assert 0 < np.abs(ref["dx_cc"][inner] / ref0["dx_cc"][inner] - 1).max() < 1e-4
