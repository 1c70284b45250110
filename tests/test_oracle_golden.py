"""The oracle against every known-answer the reference holds for the path:
README.md:54-59, test/test_zipper_boundary_conditions.jl, test/runtests.jl:8-41,
test/test_tripolar_grid.jl:36-76 -- plus the structural invariants that follow from the
reference's fill order.  CPU only."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_readme_known_answers(oracle):
    """README.md:54-59: `grid = TripolarGrid(size=(60, 30, 1))` prints
         centered at (λ, φ) = (70.0, 1.8005)
         longitude: extent 359.885 degrees, min(Δλ)=0.279019, max(Δλ)=6.32049
         latitude:  extent 175.667 degrees, min(Δφ)=0.429975, max(Δφ)=5.86207"""
    p = orc.make_params((60, 30, 1))
    s = orc.readme_summary(oracle.build(p), p)
    assert f"{s.centre[0]:.6g}" == "70" and f"{s.centre[1]:.5g}" == "1.8005"
    assert f"{s.lon_extent:.6g}" == "359.885"
    assert f"{s.dlam_min:.6g}" == "0.279019" and f"{s.dlam_max:.6g}" == "6.32049"
    assert f"{s.lat_extent:.6g}" == "175.667"
    assert f"{s.dphi_min:.6g}" == "0.429975" and f"{s.dphi_max:.6g}" == "5.86207"


def test_readme_known_answers_libm_variant(oracle_libm):
    p = orc.make_params((60, 30, 1))
    s = orc.readme_summary(oracle_libm.build(p), p)
    assert f"{s.lon_extent:.6g}" == "359.885" and f"{s.dlam_max:.6g}" == "6.32049" and f"{s.dphi_min:.6g}" == "0.429975"


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_reference_unit_tests(oracle, dtype):
    """test/runtests.jl:8-41"""
    p = orc.make_params((4, 5, 1), first_pole_longitude=75, north_poles_latitude=35, southernmost_latitude=-80)
    g = oracle.build(p, dtype=dtype)
    lam, phi = orc.interior(g["lam_cc"], p), orc.interior(g["phi_cc"], p)
    assert lam.dtype == dtype
    min_dphi = (phi[1, :] - phi[0, :]).min()
    assert lam.min() >= 0 and lam.max() <= 360 and phi.max() <= 90
    assert (phi + min_dphi / 10).min() >= -80


def _zipper_fill(oracle, c, Nx, Ny, loc, sign, H=4):
    return oracle.fill_halo_regions(c, Nx, Ny, 1, H, H, H, loc[0], loc[1], sign)


def test_reference_zipper_known_answers(oracle):
    """test/test_zipper_boundary_conditions.jl:27-45: constant fields on a 10x10x1 grid."""
    Nx = Ny = 10
    H = 4
    shape = (1 + 2 * H, Ny + 2 * H, Nx + 2 * H)

    def ones():
        c = np.zeros(shape)
        c[H, H:H + Ny, H:H + Nx] = 1
        return c
    c = _zipper_fill(oracle, ones(), Nx, Ny, (0, 0), 1)
    u = _zipper_fill(oracle, ones(), Nx, Ny, (1, 0), -1)
    v = _zipper_fill(oracle, ones(), Nx, Ny, (0, 1), -1)
    north = slice(H + Ny, H + Ny + H)
    assert (c[H, north, :] == 1).all()                               # :35
    assert (v[H, north, :] == -1).all()                              # :36
    assert (u[H, north, H + 1:H + Nx - 1] == -1).all()               # :39-40  i = 2..Nx-1
    assert (u[H, north, H] == 1).all()                               # :42,44  i = 1
    assert (u[H, north, H + Nx] == 1).all()                          # :43,45  i = Nx+1


def test_reference_zipper_fold_row(oracle):
    """test/test_zipper_boundary_conditions.jl:54,65-72: after the fill row Ny of a Center field
    equals its reverse, and a (Face, Center) field with sign -1 is antisymmetric away from the poles."""
    Nx = Ny = 10
    H = 4
    rng = np.random.default_rng(0)
    shape = (1 + 2 * H, Ny + 2 * H, Nx + 2 * H)
    c = np.zeros(shape)
    c[H, H:H + Ny, H:H + Nx] = rng.uniform(0, 360, (Ny, Nx))
    u = c.copy()
    _zipper_fill(oracle, c, Nx, Ny, (0, 0), 1)
    _zipper_fill(oracle, u, Nx, Ny, (1, 0), -1)
    row = c[H, H + Ny - 1, H:H + Nx]
    assert np.array_equal(row, row[::-1])                            # :54, :65
    urow = u[H, H + Ny - 1, H:H + Nx]
    assert np.array_equal(urow[1:5], -urow[6:10][::-1])              # :68-72


def test_orthogonality_of_the_one_degree_grid(oracle):
    """test/test_tripolar_grid.jl:36-76: angle between grid lines at FF nodes of the 1 degree grid
    with poles at (75, 35 N), excluding +-5 degrees around the singularities and phi < -78.
    The reference bounds the forward-difference angle (its kernel, :8-34) by that of a 90x90
    conformal cubed-sphere panel (external, not reproducible here); a forward difference on a
    1 degree mesh carries an O(dx) error of up to ~1.6 degrees next to the masked pole windows,
    so the restated bound is 2 degrees, and the second-order (centred) angle -- the actual
    orthogonality of the ellipse/hyperbola families -- must be below 0.01 degree."""
    p = orc.make_params((360, 180, 1), first_pole_longitude=75, north_poles_latitude=35)
    g = oracle.build(p)
    lamd, phid = orc.interior(g["lam_ff"], p), orc.interior(g["phi_ff"], p)
    lam, phi = np.deg2rad(lamd), np.deg2rad(phid)
    P = np.stack([np.cos(lam) * np.cos(phi), np.sin(lam) * np.cos(phi), np.sin(phi)], axis=-1)
    masked = ((np.abs(lamd - 75) < 5) & (np.abs(35 - phid) < 5)) | ((np.abs(lamd - 255) < 5) & (np.abs(35 - phid) < 5)) | (phid < -78)

    def angle(v1, v2):
        cos = (v1 * v2).sum(-1) / (np.linalg.norm(v1, axis=-1) * np.linalg.norm(v2, axis=-1))
        return np.rad2deg(np.arccos(np.clip(cos, -1, 1))) - 90
    fwd = angle(P[:-1, 1:] - P[:-1, :-1], P[1:, :-1] - P[:-1, :-1])
    assert np.abs(fwd[~masked[:-1, :-1]]).max() < 2.0
    ctr = angle(P[1:-1, 2:] - P[1:-1, :-2], P[2:, 1:-1] - P[:-2, 1:-1])
    assert np.abs(ctr[~masked[1:-1, 1:-1]]).max() < 0.01


@pytest.mark.parametrize("size,halo", [((60, 30, 1), (4, 4, 4)), ((10, 10, 1), (4, 4, 4)), ((24, 13, 1), (3, 6, 1))])
def test_structural_invariants(oracle, size, halo):
    """Consequences of the reference's fill order."""
    p = orc.make_params(size, halo)
    g = oracle.build(p)
    Nx, Ny, Hx, Hy = p.Nx, p.Ny, p.Hx, p.Hy
    at = lambda n, i, j: orc.at(g[n], p, i, j)
    for n in orc.ALL_NAMES:
        a = g[n]
        assert np.isfinite(a).all(), n
        assert np.array_equal(a[:, :Hx], a[:, Nx:Nx + Hx]) and np.array_equal(a[:, Nx + Hx:], a[:, Hx:2 * Hx]), n
    for n in orc.COORD_NAMES:
        assert (g[n][:Hy, :] == 0).all(), n                              # south = nothing
    for q in ("lam", "phi", "dx", "dy", "az"):
        cc = np.array([at(f"{q}_cc", i, Ny) for i in range(1, Nx + 1)])
        assert np.array_equal(cc, cc[::-1]), q                           # row-Ny overwrite
        fc = np.array([at(f"{q}_fc", i, Ny) for i in range(2, Nx + 1)])
        assert np.array_equal(fc, fc[::-1]), q
    for n in ("dx_fc", "dy_fc", "az_fc"):                                # hav of identical points
        assert at(n, 1, Ny) == 0.0 and at(n, Nx // 2 + 1, Ny) == 0.0, n
    # the two grid singularities (docstring src/tripolar_grid.jl:56-57) sit at north_poles_latitude
    # (only when Nx is a multiple of 4, so that the circshift by Nx/4 is exact)
    if Nx % 4 == 0:
        assert abs(at("phi_fc", 1, Ny) - 55) < 1e-9 and abs(at("phi_fc", Nx // 2 + 1, Ny) - 55) < 1e-9
    # first_pole_longitude (70) lands at i = Nx/2+1 (README: "centered at (70.0, ...)"), its antipode at i = 1
        assert abs(at("lam_fc", Nx // 2 + 1, Ny) - 70) < 1e-9 and abs(at("lam_fc", 1, Ny) - 250) < 1e-9
    # south continuation: rows j <= 1 are constant in i and equal the lat-lon formula
    R, d2r = p.radius, np.pi / 180
    for n in orc.METRIC_NAMES:
        rows = g[n][:Hy + 1, :]
        assert (rows == rows[:, :1]).all(), n
    assert at("dy_cc", 1, 1) == R * ((90 - (-80)) / Ny * (np.pi / 180))
    assert at("dx_cc", 1, 1 - Hy) == 0.0 and at("az_ff", 3, 1 - Hy) == 0.0       # never written by the lat-lon precompute
    phiC1 = -80 + 0.5 * 170 / Ny
    assert abs(at("dx_cc", 1, 1) / (R * (360 / Nx) * d2r * np.cos(phiC1 * d2r)) - 1) < 1e-14


def test_libm_sensitivity_motivates_shared_arithmetic(oracle, oracle_libm):
    """Two valid libms (this oracle's deterministic functions vs glibc) give coordinates that agree
    to ~1e-13 but cell areas that differ by far more than 1e-12 at 1 degree: the area formula
    cancels.  Hence the GPU path states the same arithmetic as the oracle
    and is compared bit-for-bit."""
    p = orc.make_params((360, 180, 1))
    a, b = oracle.build(p), oracle_libm.build(p)
    rel = lambda n: float(np.max(np.abs(a[n] - b[n]) / np.maximum(np.abs(a[n]), 1e-300)))
    assert rel("phi_ff") < 1e-10 and rel("dx_cc") < 1e-9
    assert rel("az_cc") > 1e-12


def _checksums(oracle):
    out = {}
    for key, (size, halo, kw) in {
        "A_60x30": ((60, 30, 1), (4, 4, 4), {}),
        "unit_4x5": ((4, 5, 1), (4, 4, 4), dict(first_pole_longitude=75, north_poles_latitude=35)),
        "zipper_10x10": ((10, 10, 1), (4, 4, 4), {}),
        "bickley_120x60_h5": ((120, 60, 1), (5, 5, 5), dict(north_poles_latitude=45, first_pole_longitude=25)),
        "B_360x180": ((360, 180, 1), (4, 4, 4), {}),
    }.items():
        g = oracle.build(orc.make_params(size, halo, **kw))
        out[key] = {n: hashlib.sha256(np.ascontiguousarray(g[n]).tobytes()).hexdigest()[:16] for n in orc.ALL_NAMES}
    return out


def test_oracle_output_is_pinned_by_committed_checksums(oracle):
    """tests/golden/oracle_checksums.json (generated by tools/make_golden.py from this oracle
    after it reproduced the README numbers) pins every bit of the oracle's output for five grids,
    so that a later edit of the oracle cannot silently move the parity target."""
    want = json.load(open(os.path.join(GOLDEN, "oracle_checksums.json")))
    assert _checksums(oracle) == want
