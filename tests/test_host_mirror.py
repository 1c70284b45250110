"""The product's per-point headers (index maps + formulas executed by the CUDA kernels), built
for the host by tests/host_mirror, against the oracle's multi-pass restatement -- bit for bit,
including tile windows (== slices of the serial grid).  CPU only; the GPU tests repeat the
comparison on the real kernels."""
import ctypes as C

import numpy as np
import pytest

import host_mirror as hm
from oracle import oracle as orc

CONFIGS = [
    ((60, 30, 1), (4, 4, 4), {}),
    ((4, 5, 1), (4, 4, 4), dict(first_pole_longitude=75, north_poles_latitude=35)),
    ((10, 10, 1), (4, 4, 4), {}),
    ((120, 60, 1), (5, 5, 5), dict(north_poles_latitude=45, first_pole_longitude=25)),
    ((40, 34, 1), (5, 31, 5), dict(north_poles_latitude=45, first_pole_longitude=35)),
    ((12, 8, 1), (2, 7, 1), {}),
    ((66, 37, 1), (3, 2, 1), dict(south=-75.5)),
    ((8, 3, 1), (1, 1, 1), {}),
]


def bits(a):
    return np.ascontiguousarray(a).view(np.uint64)


@pytest.mark.parametrize("size,halo,kw", CONFIGS)
def test_one_pass_source_maps_equal_the_multi_pass_reference_pipeline(oracle, mirror, size, halo, kw):
    okw = {("southernmost_latitude" if k == "south" else k): v for k, v in kw.items()}
    g = oracle.build(orc.make_params(size, halo, **okw))
    q = hm.make(size, halo, **kw)
    c = hm.coords(mirror, q)
    for n, a in zip(orc.COORD_NAMES, c):
        assert np.array_equal(bits(a), bits(g[n])), n
    for closed in (False, True):
        m = hm.metrics(mirror, q, c, closed)
        for n, a in zip(orc.METRIC_NAMES, m):
            assert np.array_equal(bits(a), bits(g[n])), (n, closed)


@pytest.mark.parametrize("Rx,Ry", [(1, 2), (2, 1), (2, 2), (3, 2), (4, 1)])
def test_tile_windows_equal_slices(oracle, mirror, Rx, Ry):
    """src/distributed_tripolar_grid.jl:41-49,112-120 (and the new x split)."""
    size, halo = (24, 20, 1), (3, 4, 1)
    g = oracle.build(orc.make_params(size, halo))
    sx = [size[0] // Rx] * Rx; sx[-1] += size[0] - sum(sx)
    sy = [size[1] // Ry] * Ry; sy[-1] += size[1] - sum(sy)
    for rx in range(Rx):
        for ry in range(Ry):
            i0, j0 = 1 + sum(sx[:rx]), 1 + sum(sy[:ry])
            q = hm.make(size, halo, tile=(i0, j0, sx[rx], sy[ry]))
            c = hm.coords(mirror, q)
            m = hm.metrics(mirror, q, c, True)
            for n, a in zip(orc.ALL_NAMES, c + m):
                want = g[n][j0 - 1:j0 - 1 + sy[ry] + 2 * halo[1], i0 - 1:i0 - 1 + sx[rx] + 2 * halo[0]]
                assert np.array_equal(bits(a), bits(want)), (n, rx, ry)


def test_halo_source_matches_zipper_then_periodic(oracle, mirror):
    """halo_source (csrc/ossg_tripolar_core.cuh) is the closed form of: zipper over i in 1..Nx
    (src/zipper_boundary_condition.jl:70-138), then periodic x over every row."""
    for (Nx, Ny, Hx, Hy) in [(10, 7, 4, 4), (4, 5, 4, 4), (12, 9, 2, 8), (6, 4, 3, 3)]:
        for locx in (0, 1):
            for locy in (0, 1):
                # label every interior cell with a unique id, run the oracle's fills, compare
                ids = np.full((1, Ny + 2 * Hy, Nx + 2 * Hx), -1.0)
                ids[0, Hy:Hy + Ny, Hx:Hx + Nx] = np.arange(1, Nx * Ny + 1).reshape(Ny, Nx)
                oracle.fill_halo_regions(ids, Nx, Ny, 1, Hx, Hy, 0, locx, locy, 1)
                i, j, w = C.c_int(), C.c_int(), C.c_int()
                for J in range(1 - Hy, Ny + Hy + 1):
                    for I in range(1 - Hx, Nx + Hx + 1):
                        ok = mirror.hm_halo_source(locx, locy, I, J, Nx, Ny, C.byref(i), C.byref(j), C.byref(w))
                        got = ids[0, J + Hy - 1, I + Hx - 1]
                        if not ok:
                            assert got == -1.0 and J < 1
                        else:
                            assert got == (j.value - 1) * Nx + i.value, (Nx, Ny, locx, locy, I, J)
