"""Geometry-level known answers for the oracle's cell areas, none of which uses the reference's area
formula (Eriksson's spherical-triangle formula, tripolar_grid_utils.jl:23-43): the four Az arrays have
no vector in the reference's own tests, so they are pinned here by identities
that any correct composition must obey.  40-digit mpmath, CPU only.

1. Tiling identity.  The cells of Az_cc are the great-circle quadrilaterals through FF[i,j], FF[i+1,j],
   FF[i+1,j+1], FF[i,j+1]: rows 2..Ny-1 tile the cap north of the closed polygon through FF[.,2] up to
   row Ny, and the cells of row Ny reach across the fold to the mirror image of FF[.,Ny], so each of
   them is the same physical cell as its mirror and is counted twice.  Hence
       sum(Az_cc[:, 2..Ny-1]) + sum(Az_cc[:, Ny]) / 2  ==  R^2 * area north of the polygon FF[., 2]
   and likewise, with the CC points of row Ny lying ON the fold line,
       sum(Az_ff[:, 2..Ny])  ==  R^2 * area north of the polygon CC[., 1].
   The right-hand sides come from the Gauss-Bonnet theorem (area = 2 pi - sum of the turning angles of the
   geodesic polygon), evaluated in mpmath from the oracle's own FF / CC coordinates.  A wrong point
   order, staggering offset, fold pairing or row range in the oracle's (or the reference's) metric
   composition changes the left-hand side by O(1 cell); FP64 rounding changes it by ~1e-16 (measured: 6e-17 on the README grid, 3e-16 at 1 degree; asserted at 1e-13).
2. The same areas on a sample of 1/4 degree cells against an mpmath evaluation of rows C1 and M1 from
   scratch: the FP64 formula itself carries ~3e-16 / (cell area in steradians) of cancellation error,
   about 1e-11 relative at 1/4 degree -- the number behind DESIGN.md section 4's statement that a
   1e-12 tolerance on Az is only meaningful between implementations that execute the same operations.
"""
import math

import numpy as np
import pytest

mp = pytest.importorskip("mpmath")
from mpmath import mpf  # noqa: E402

from oracle import oracle as orc  # noqa: E402

mp.mp.dps = 40


def _rad(x):
    return x * mp.pi / 180


def _unit(lam, phi):
    l, p = _rad(mpf(float(lam))), _rad(mpf(float(phi)))
    return (mp.cos(l) * mp.cos(p), mp.sin(l) * mp.cos(p), mp.sin(p))


def _dot(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def _cross(a, b):
    return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def _tangent(at, to):
    """unit tangent at `at` of the great circle towards `to`"""
    d = _dot(at, to)
    t = tuple(to[k] - d * at[k] for k in range(3))
    n = mp.sqrt(_dot(t, t))
    return tuple(x / n for x in t)


def _area_left_of(vertices):
    """area enclosed on the LEFT of the closed geodesic polygon v0 -> v1 -> ... -> v0 on the unit sphere:
    2 pi - sum of the signed turning angles (Gauss-Bonnet; geodesic edges carry no curvature)"""
    n = len(vertices)
    total = mpf(0)
    for k in range(n):
        prev, cur, nxt = vertices[k - 1], vertices[k], vertices[(k + 1) % n]
        t_in = tuple(-x for x in _tangent(cur, prev))          # direction of travel on arrival
        t_out = _tangent(cur, nxt)
        total += mp.atan2(_dot(cur, _cross(t_in, t_out)), _dot(t_in, t_out))
    return 2 * mp.pi - total


@pytest.mark.parametrize("size,kw", [
    ((60, 30, 1), {}),                                            # BASELINE config A (README.md:54)
    ((360, 180, 1), {}),                                          # BASELINE config B
    ((36, 22, 1), dict(north_poles_latitude=35, first_pole_longitude=75, southernmost_latitude=-75)),
])
def test_cell_areas_tile_the_cap_north_of_their_lowest_row(oracle, size, kw):
    Nx, Ny, _ = size
    p = orc.make_params(size, **kw)
    g = oracle.build(p)
    R2 = mpf(p.radius) ** 2
    I = lambda name: orc.interior(g[name], p)                     # noqa: E731  (Ny, Nx), index [j-1, i-1]
    # --- Az_cc against the polygon through FF[., 2] (increasing i runs eastward: the cap is on the left)
    poly = [_unit(I("lam_ff")[1, i], I("phi_ff")[1, i]) for i in range(Nx)]
    cap = _area_left_of(poly) * R2
    az = I("az_cc")
    lhs = mpf(math.fsum(az[1:Ny - 1].ravel().tolist())) + mpf(math.fsum(az[Ny - 1].tolist())) / 2
    assert abs(lhs - cap) <= mpf("1e-13") * cap, (float(lhs), float(cap), float((lhs - cap) / cap))
    # the row-Ny cells really come in mirror pairs (M2: the y-Center overwrite, tripolar_grid.jl:230-273)
    assert np.array_equal(az[Ny - 1], az[Ny - 1][::-1])
    # --- Az_ff against the polygon through CC[., 1]
    poly = [_unit(I("lam_cc")[0, i], I("phi_cc")[0, i]) for i in range(Nx)]
    cap = _area_left_of(poly) * R2
    azf = I("az_ff")
    lhs = mpf(math.fsum(azf[1:Ny].ravel().tolist()))
    assert abs(lhs - cap) <= mpf("1e-13") * cap, (float(lhs), float(cap), float((lhs - cap) / cap))
    # --- sanity of the arbiter itself: the cap north of a latitude circle sampled by many points
    ring = [_unit(mpf(360) * k / 2000, mpf(-60)) for k in range(2000)]
    exact = 2 * mp.pi * (1 - mp.sin(_rad(mpf(-60))))
    assert abs(_area_left_of(ring) - exact) < mpf("2e-5")      # (a 2000-gon is not the circle; the formula is exact for the polygon)


def _point_mp(Nx, Ny, south, poles_lat, fpl, face_x, face_y, i, j):
    """rows C0, C1, C2 for ONE post-circshift point (i, j) in mpmath (src/generate_tripolar_coordinates.jl:60-87);
    i may leave 1..Nx (periodic)"""
    i = (i - 1) % Nx + 1
    ip = i - Nx // 4
    if ip < 1:
        ip += Nx
    a = mp.tan(_rad((90 - mpf(poles_lat)) / 2))
    lam1 = -180 + mpf(ip - 1) * 360 / Nx + (0 if face_x else mpf(180) / Nx)
    dphi = (90 - mpf(south)) / (Ny - 1)
    phi1 = mpf(south) + mpf(j - 1) * dphi - (dphi / 2 if face_y else 0)
    psi = mp.asinh(mp.tan(_rad((90 - phi1) / 2)) / a)
    x = a * mp.sin(_rad(lam1)) * mp.cosh(psi)
    y = a * mp.cos(_rad(lam1)) * mp.sinh(psi)
    lam = -(180 / mp.pi) * mp.atan(y / x) + (-90 if ip <= Nx // 2 else 90) + mpf(fpl) + 90
    phi = 90 - (360 / mp.pi) * mp.atan(mp.sqrt(y * y + x * x))
    return lam, phi


def _unit_mp(P):
    l, p = _rad(P[0]), _rad(P[1])
    return (mp.cos(l) * mp.cos(p), mp.sin(l) * mp.cos(p), mp.sin(p))


def _girard_quadrilateral(a, b, c, d):
    """area of the spherical quadrilateral a-b-c-d by Girard's theorem (sum of interior angles - 2 pi)"""
    v = [a, b, c, d]
    s = mpf(0)
    for k in range(4):
        t1, t2 = _tangent(v[k], v[k - 1]), _tangent(v[k], v[(k + 1) % 4])
        s += mp.acos(_dot(t1, t2))
    return s - 2 * mp.pi


def test_quarter_degree_areas_against_mpmath_sample(oracle):
    """Az_cc and Az_ff of 160 cells of the 1/4 degree grid (BASELINE config C) from scratch in mpmath, with
    Girard's theorem instead of the reference's triangle formula.  Asserted at 1e-9; the measured
    worst case (printed with -s) is the FP64 cancellation floor of the reference's own formula."""
    Nx, Ny = 1440, 720
    p = orc.make_params((Nx, Ny, 1))
    g = oracle.build(p)
    rng = np.random.default_rng(2)
    worst = {"az_cc": 0.0, "az_ff": 0.0}
    args = (Nx, Ny, p.southernmost_latitude, p.north_poles_latitude, p.first_pole_longitude)
    R2 = mpf(p.radius) ** 2
    for _ in range(80):
        i, j = int(rng.integers(2, Nx)), int(rng.integers(3, Ny - 1))
        if (i - Nx // 4 - 1) % (Nx // 4) in (0, 1, Nx // 4 - 1):                 # keep off the pole meridians (signed zeros)
            continue
        ff = [_unit_mp(_point_mp(*args, True, True, i + di, j + dj)) for di, dj in ((0, 0), (1, 0), (1, 1), (0, 1))]
        cc = [_unit_mp(_point_mp(*args, False, False, i + di, j + dj)) for di, dj in ((-1, -1), (0, -1), (0, 0), (-1, 0))]
        for name, quad in (("az_cc", ff), ("az_ff", cc)):
            want = _girard_quadrilateral(*quad) * R2
            have = mpf(float(orc.at(g[name], p, i, j)))
            rel = abs(have - want) / want
            worst[name] = max(worst[name], float(rel))
            assert rel < mpf("1e-9"), (name, i, j, float(have), float(want))
    print("worst relative deviation of the FP64 areas from the exact ones at 1/4 degree:", worst)
    assert max(worst.values()) > 1e-14      # (the floor is visible: these are not the same operations in higher precision)
