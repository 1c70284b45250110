"""An independent arbiter for the oracle: the formulas of the path (SURVEY.md section 8a rows C0, C1,
C2, M1) restated in 40-digit mpmath arithmetic -- a different language, a different number system
and none of the oracle's elementary-function code -- and compared with the oracle's Float64 output
on small grids.  This checks the oracle's *composition* (index conventions, staggering pairs,
circshift, argument order of the Delta-y block, the triangle decomposition of the cell areas), which
the reference's own tests pin only through the six README numbers.  CPU only.

Reference lines followed: src/tripolar_grid.jl:76,90,95-97,121-130; src/generate_tripolar_coordinates.jl:
60-87; src/OrthogonalSphericalShellGrids.jl:24; src/tripolar_grid_utils.jl:13-43; (external) Distances.jl
haversine, Oceananigans lat_lon_to_cartesian / spherical_area_quadrilateral.

Points on the pole meridians and the pole row (x == 0 or y == 0 in the ellipse/hyperbola map) are
left to the known-answer tests: there the result hangs on Julia's signed zeros, which a
multiprecision evaluation of sin(pi) does not reproduce.
"""
import pytest

mp = pytest.importorskip("mpmath")
from mpmath import mpf  # noqa: E402

from oracle import oracle as orc  # noqa: E402

mp.mp.dps = 40


def _rad(x):
    return x * mp.pi / 180


def _coordinates_mp(Nx, Ny, south, poles_lat, first_pole_lon):
    """(lam, phi, generic) per staggering, each a dict (i, j) -> mpf over the interior, post-circshift."""
    south, poles_lat, first_pole_lon = mpf(south), mpf(poles_lat), mpf(first_pole_lon)
    a = mp.tan(_rad((90 - poles_lat) / 2))                                     # tripolar_grid.jl:76
    lamF = [-180 + mpf(i - 1) * 360 / Nx for i in range(1, Nx + 1)]            # :90 (regular, Periodic)
    lamC = [lf + mpf(180) / Nx for lf in lamF]
    phiC = [south + mpf(j - 1) * (90 - south) / (Ny - 1) for j in range(1, Ny + 1)]   # :95  range(south, 90, length = Ny)
    dphi = phiC[1] - phiC[0]                                                   # :96
    phiF = [pc - dphi / 2 for pc in phiC]                                      # :97
    out = {}
    for name, lam1, phi1 in (("ff", lamF, phiF), ("fc", lamF, phiC), ("cf", lamC, phiF), ("cc", lamC, phiC)):   # coordinates.jl:60-63
        lam, phi, generic = {}, {}, {}
        for j in range(1, Ny + 1):
            psi = mp.asinh(mp.tan(_rad((90 - phi1[j - 1]) / 2)) / a)          # :66
            for ip in range(1, Nx + 1):                                        # pre-shift index
                x = a * mp.sin(_rad(lam1[ip - 1])) * mp.cosh(psi)             # :67
                y = a * mp.cos(_rad(lam1[ip - 1])) * mp.sinh(psi)             # :68
                exact_axis = (lam1[ip - 1] % 90 == 0) or phi1[j - 1] == 90
                if exact_axis:
                    l, p = mpf(0), mpf(0)
                else:
                    l = -(180 / mp.pi) * mp.atan(y / x)                       # :77
                    p = 90 - (360 / mp.pi) * mp.atan(mp.sqrt(y * y + x * x))  # :78
                    l += -90 if ip <= Nx // 2 else 90                          # :82
                    l += first_pole_lon + 90                                   # :86
                    l = (mp.fmod(l, 360) + 360) % 360                          # :87, convert_to_0_360
                i = ip + Nx // 4                                               # tripolar_grid.jl:121-130: new[i] = old[mod1(i - Nx/4, Nx)]
                if i > Nx:
                    i -= Nx
                lam[(i, j)], phi[(i, j)], generic[(i, j)] = l, p, not exact_axis
        out[name] = (lam, phi, generic)
    return out


def _haversine_mp(P, Q, R):
    (l1, p1), (l2, p2) = P, Q
    dl, dp = _rad(l2 - l1), _rad(p2 - p1)
    h = mp.sin(dp / 2) ** 2 + mp.cos(_rad(p1)) * mp.cos(_rad(p2)) * mp.sin(dl / 2) ** 2
    return 2 * R * mp.asin(min(mp.sqrt(h), mpf(1)))


def _unit(P):
    l, p = P
    return (mp.cos(_rad(l)) * mp.cos(_rad(p)), mp.sin(_rad(l)) * mp.cos(_rad(p)), mp.sin(_rad(p)))


def _dot(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def _cross(a, b):
    return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def _triangle(a, b, c):
    return 2 * mp.atan(abs(_dot(a, _cross(b, c))) / (1 + _dot(a, b) + _dot(b, c) + _dot(a, c)))


def _quadrilateral(a, b, c, d):
    return (_triangle(a, b, c) + _triangle(a, b, d) + _triangle(a, c, d) + _triangle(b, c, d)) / 2


@pytest.mark.parametrize("size,kw", [
    ((60, 30, 1), {}),                                                                   # README.md:54
    ((36, 22, 1), dict(north_poles_latitude=35, first_pole_longitude=75, southernmost_latitude=-75)),
])
def test_oracle_equals_the_multiprecision_restatement(oracle, size, kw):
    Nx, Ny, _ = size
    p = orc.make_params(size, **kw)
    south, poles, fpl, R = p.southernmost_latitude, p.north_poles_latitude, p.first_pole_longitude, mpf(p.radius)
    got = oracle.build(p)
    ref = _coordinates_mp(Nx, Ny, south, poles, fpl)

    # ---- coordinates, interior (C1 + C2)
    checked = 0
    for name in ("cc", "fc", "cf", "ff"):
        lam, phi, generic = ref[name]
        glam, gphi = orc.interior(got["lam_" + name], p), orc.interior(got["phi_" + name], p)
        for (i, j), ok in generic.items():
            if not ok:
                continue
            d = abs(mpf(float(glam[j - 1, i - 1])) - lam[(i, j)])
            d = min(d, abs(360 - d))
            assert d < mpf("1e-12"), (name, i, j, float(d))
            assert abs(mpf(float(gphi[j - 1, i - 1])) - phi[(i, j)]) < mpf("1e-12"), (name, i, j)
            checked += 1
    assert checked > 3.5 * Nx * Ny                          # only the pole meridians / pole row were skipped

    # ---- metrics, rows 2..Ny-1 (row 1 is the lat-lon continuation M3, row Ny reads the fold image),
    #      periodic in i, skipping stencils that touch a non-generic point (tripolar_grid_utils.jl:13-43)
    def pt(name, i, j):
        i = i + Nx if i < 1 else (i - Nx if i > Nx else i)
        lam, phi, generic = ref[name]
        return (lam[(i, j)], phi[(i, j)]) if generic[(i, j)] else None

    formulas = {
        "dx_cc": lambda i, j: ("h", pt("fc", i + 1, j), pt("fc", i, j)),            # :13
        "dx_fc": lambda i, j: ("h", pt("cc", i, j), pt("cc", i - 1, j)),            # :14
        "dx_cf": lambda i, j: ("h", pt("ff", i + 1, j), pt("ff", i, j)),            # :15
        "dx_ff": lambda i, j: ("h", pt("cf", i, j), pt("cf", i - 1, j)),            # :16
        "dy_cc": lambda i, j: ("h", pt("cf", i, j + 1), pt("cf", i, j)),            # :18
        "dy_fc": lambda i, j: ("h", pt("ff", i, j + 1), pt("ff", i, j)),            # :19
        "dy_cf": lambda i, j: ("h", pt("cc", i, j), pt("cc", i, j - 1)),            # :20
        "dy_ff": lambda i, j: ("h", pt("fc", i, j), pt("fc", i, j - 1)),            # :21
        "az_cc": lambda i, j: ("q", pt("ff", i, j), pt("ff", i + 1, j), pt("ff", i + 1, j + 1), pt("ff", i, j + 1)),   # :23-28
        "az_ff": lambda i, j: ("q", pt("cc", i - 1, j - 1), pt("cc", i, j - 1), pt("cc", i, j), pt("cc", i - 1, j)),   # :38-43
    }
    values = {}
    nchecked = 0
    for name, f in formulas.items():
        g = orc.interior(got[name], p)
        for j in range(2, Ny):
            for i in range(1, Nx + 1):
                kind, *pts = f(i, j)
                if any(q is None for q in pts):
                    continue
                want = _haversine_mp(pts[0], pts[1], R) if kind == "h" else _quadrilateral(*[_unit(q) for q in pts]) * R * R
                values[(name, i, j)] = want
                have = mpf(float(g[j - 1, i - 1]))
                tol = mpf("1e-12")                                   # the north-star tolerance; measured: 5e-14 (spacings), 3e-14 (areas)
                assert abs(have - want) <= tol * abs(want), (name, i, j, float(have), float(want))
                nchecked += 1
    assert nchecked > 9 * Nx * (Ny - 2)
    # product areas (:34-35)
    for name, a, b in (("az_fc", "dy_fc", "dx_fc"), ("az_cf", "dy_cf", "dx_cf")):
        g = orc.interior(got[name], p)
        for j in range(2, Ny):
            for i in range(1, Nx + 1):
                if (a, i, j) in values and (b, i, j) in values:
                    want = values[(a, i, j)] * values[(b, i, j)]
                    assert abs(mpf(float(g[j - 1, i - 1])) - want) <= mpf("1e-12") * abs(want), (name, i, j)


@pytest.mark.parametrize("size,kw", [
    ((60, 30, 1), {}),
    ((36, 22, 1), dict(north_poles_latitude=35, first_pole_longitude=75, southernmost_latitude=-75)),
    ((10, 10, 1), {}),                                                       # Nx % 4 != 0: no Center column sits on an axis
])
def test_pole_meridians_and_pole_row_against_their_limits(oracle, size, kw):
    """The points the test above leaves out: lambda_1D a multiple of 90 degrees (x == 0 or y == 0 in the
    ellipse/hyperbola map, generate_tripolar_coordinates.jl:67-68) and the pole row phi_1D == 90.  There
    the latitude is still unambiguous -- phi = 90 - (360/pi) atan(|y|) resp. atan(|x|), :78 -- and is
    checked exactly; the longitude hangs on the sign of a zero (atan(y / +-0), :77), which Julia's sind /
    cosd decide: whatever they decide, the result must be one of the two one-sided limits of the map, and
    exactly on the two poles (x == 0 and y == 0) the reference prescribes it itself (:73-75).  Together
    with the generic test, every interior coordinate of the grid is now accounted for."""
    Nx, Ny, _ = size
    p = orc.make_params(size, **kw)
    south, poles, fpl = mpf(p.southernmost_latitude), mpf(p.north_poles_latitude), mpf(p.first_pole_longitude)
    got = oracle.build(p)
    a = mp.tan(_rad((90 - poles) / 2))
    lamF = [-180 + mpf(i - 1) * 360 / Nx for i in range(1, Nx + 1)]
    lamC = [lf + mpf(180) / Nx for lf in lamF]
    phiC = [south + mpf(j - 1) * (90 - south) / (Ny - 1) for j in range(1, Ny + 1)]
    dphi = phiC[1] - phiC[0]
    phiF = [pc - dphi / 2 for pc in phiC]
    eps = mpf("1e-25")

    def general(l1, ph1, ip):
        psi = mp.asinh(mp.tan(_rad((90 - ph1) / 2)) / a)
        x, y = a * mp.sin(_rad(l1)) * mp.cosh(psi), a * mp.cos(_rad(l1)) * mp.sinh(psi)
        lam = -(180 / mp.pi) * mp.atan(y / x) + (-90 if ip <= Nx // 2 else 90) + fpl + 90
        return (mp.fmod(lam, 360) + 360) % 360, 90 - (360 / mp.pi) * mp.atan(mp.sqrt(x * x + y * y))

    checked = poles_seen = 0
    for name, lam1, phi1 in (("ff", lamF, phiF), ("fc", lamF, phiC), ("cf", lamC, phiF), ("cc", lamC, phiC)):
        glam, gphi = orc.interior(got["lam_" + name], p), orc.interior(got["phi_" + name], p)
        for j in range(1, Ny + 1):
            for i in range(1, Nx + 1):
                # C3 (tripolar_grid.jl:137-199): interior row Ny of the y-Center coordinates is overwritten by its
                # fold mirror for i > Nx/2 (zipper_boundary_condition.jl:102, :135) -- the expected value is the source's
                isrc = i
                if name in ("fc", "cc") and j == Ny and i > Nx // 2:
                    isrc = Nx - i + (2 if name == "fc" else 1)
                    if isrc > Nx:
                        isrc -= Nx
                ip = isrc - Nx // 4                                             # tripolar_grid.jl:121-130
                if ip < 1:
                    ip += Nx
                on_axis = lam1[ip - 1] % 90 == 0
                if not (on_axis or phi1[j - 1] == 90):
                    continue
                have_l, have_p = mpf(float(glam[j - 1, i - 1])), mpf(float(gphi[j - 1, i - 1]))
                at_pole = phi1[j - 1] == 90 and lam1[ip - 1] % 180 == 0        # x == 0 and y == 0
                if at_pole:
                    want = (mpf(-90 if ip == 1 else 90) + (-90 if ip <= Nx // 2 else 90) + fpl + 90) % 360     # :73-75, :82, :86-87
                    d = abs(have_l - want)
                    assert min(d, abs(360 - d)) < mpf("1e-12") and abs(have_p - 90) < mpf("1e-12"), (name, ip, j)
                    poles_seen += 1
                    continue
                # one-sided limits in whichever variable crosses an axis
                dl = eps if on_axis else 0
                dp = eps if phi1[j - 1] == 90 else 0
                cands = [general(lam1[ip - 1] + s * dl, phi1[j - 1] - dp, ip) for s in ((-1, 1) if on_axis else (1,))]
                if phi1[j - 1] == 90 and not on_axis:
                    cands.append(((cands[0][0] + 180) % 360, cands[0][1]))      # sinh(psi) = +-0: atan(+-0 / x)
                assert abs(have_p - cands[0][1]) < mpf("1e-12"), (name, ip, j, float(have_p), float(cands[0][1]))
                ds = []
                for (cl, _) in cands:
                    d = abs(have_l - cl)
                    ds.append(min(d, abs(360 - d)))
                assert min(ds) < mpf("1e-12"), (name, ip, j, float(have_l), [float(c[0]) for c in cands])
                checked += 1
    assert poles_seen == 2                                        # two pole values in the FC row Ny (for Nx % 4 != 0 the overwrite repeats the first)
    assert checked >= 2 * (2 * Ny - 1)
