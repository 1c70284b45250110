"""Elementary functions of the oracle (oracle/ossg_oracle_math.h) and of the product
(csrc/ossg_math.cuh, through the test-only host build) against 60-digit mpmath and glibc, and
against each other bit for bit.  CPU only."""
import mpmath as mp
import numpy as np
import pytest

import host_mirror as hm
from oracle import oracle as orc

mp.mp.prec = 200
RNG = np.random.default_rng(20261018)


def ulp_errors(ys, xs, f):
    worst = 0.0
    for x, y in zip(xs, ys):
        t = f(mp.mpf(float(x)))
        if t == 0:
            assert y == 0
            continue
        u = abs(np.spacing(float(t)))
        worst = max(worst, float(abs((mp.mpf(float(y)) - t) / mp.mpf(float(u)))))
    return worst


deg = lambda f: (lambda x: f(x * mp.pi / 180))
CASES = [
    # name, inputs, exact function, ulp bound
    ("sind", RNG.uniform(-720, 720, 1500), deg(mp.sin), 0.8),
    ("cosd", RNG.uniform(-720, 720, 1500), deg(mp.cos), 0.8),
    ("sind", RNG.uniform(-1e-3, 1e-3, 500), deg(mp.sin), 0.8),
    ("tand", RNG.uniform(0, 89.9, 1000), deg(mp.tan), 2.0),
    ("sin", RNG.uniform(-7, 7, 1500), mp.sin, 0.8),
    ("cos", RNG.uniform(-7, 7, 1500), mp.cos, 0.8),
    ("sin", np.array([np.pi, np.pi / 2, -np.pi, 2 * np.pi, 1e-9, 3.0e-3]), mp.sin, 0.8),
    ("cos", np.array([np.pi / 2, 3 * np.pi / 2, 1.5707963267948966]), mp.cos, 0.8),
    ("atan", np.concatenate([RNG.uniform(-5, 5, 1500), 10.0 ** RNG.uniform(-10, 10, 1000)]), mp.atan, 0.8),
    ("asin", np.concatenate([RNG.uniform(-1, 1, 1500), 10.0 ** RNG.uniform(-9, 0, 1000), [1.0, -1.0, 0.5, 0.975]]), mp.asin, 0.8),
    ("asinh", 10.0 ** RNG.uniform(-8, 3, 300), mp.asinh, 0.501),
    ("sinh", np.concatenate([10.0 ** RNG.uniform(-8, 1, 200), RNG.uniform(0, 8, 200)]), mp.sinh, 0.501),
    ("cosh", np.concatenate([10.0 ** RNG.uniform(-8, 1, 200), RNG.uniform(0, 8, 200)]), mp.cosh, 0.501),
]


@pytest.mark.parametrize("name,xs,f,bound", CASES, ids=[f"{c[0]}-{k}" for k, c in enumerate(CASES)])
def test_oracle_functions_against_mpmath(oracle, name, xs, f, bound):
    assert ulp_errors(oracle.math(name, xs), xs, f) <= bound


@pytest.mark.parametrize("name,xs,f,bound", CASES, ids=[f"{c[0]}-{k}" for k, c in enumerate(CASES)])
def test_product_functions_are_bit_identical_to_the_oracle(oracle, mirror, name, xs, f, bound):
    a = oracle.math(name, xs)
    b = hm.math(mirror, orc.MATH_FN[name], xs)
    assert np.array_equal(a.view(np.uint64), b.view(np.uint64))


def test_degree_trig_special_values(oracle, mirror):
    """Julia Base semantics the pole meridians depend on."""
    for fn in (lambda n, x: oracle.math(n, x), lambda n, x: hm.math(mirror, orc.MATH_FN[n], x)):
        s = fn("sind", [-180.0, 0.0, 180.0, -0.0, 360.0, -360.0, 90.0, 270.0, 30.0])
        assert np.array_equal(np.signbit(s[:6]), [True, False, False, True, False, True])
        assert (s[:6] == 0).all() and s[6] == 1.0 and s[7] == -1.0 and abs(s[8] - 0.5) < 1e-16
        c = fn("cosd", [90.0, -90.0, 270.0, 180.0, 0.0, 60.0])
        assert (c[:3] == 0).all() and not np.signbit(c[:3]).any() and c[3] == -1.0 and c[4] == 1.0 and abs(c[5] - 0.5) < 1.2e-16
        assert np.isnan(fn("sind", [np.inf, np.nan])).all() and np.isnan(fn("cosd", [np.inf])).all()
        a = fn("atan", [np.inf, -np.inf, 0.0, -0.0])
        assert a[0] == np.pi / 2 and a[1] == -np.pi / 2 and a[2] == 0 and np.signbit(a[3])
        assert fn("asin", [1.0])[0] == np.pi / 2 and np.isnan(fn("asin", [1.0000001])[0])
        assert fn("sinh", [0.0])[0] == 0.0 and fn("cosh", [0.0])[0] == 1.0 and fn("asinh", [0.0])[0] == 0.0


def test_functions_agree_with_glibc_within_one_ulp(oracle, oracle_libm):
    for name, xs in (("sin", RNG.uniform(-7, 7, 20000)), ("cos", RNG.uniform(-7, 7, 20000)),
                     ("atan", RNG.uniform(-50, 50, 20000)), ("asin", RNG.uniform(-1, 1, 20000)),
                     ("sinh", RNG.uniform(0, 8, 2000)), ("asinh", RNG.uniform(0, 100, 2000))):
        a, b = oracle.math(name, xs), oracle_libm.math(name, xs)
        ulps = np.abs(a - b) / np.spacing(np.maximum(np.abs(b), 1e-300))
        assert ulps.max() <= 1.0, name


# ------------------------------------------------------------------ small-argument / hot paths
def test_small_argument_paths_against_mpmath(oracle):
    """sin and asin below 2^-5 use short Taylor forms (every grid finer than ~3.5 degrees lands
    there); they must be as accurate as the fdlibm kernels they shortcut, on both sides of the
    switch."""
    xs = np.concatenate([RNG.uniform(-2.0 ** -5, 2.0 ** -5, 3000), 10.0 ** RNG.uniform(-12, -1.5, 2000),
                         np.nextafter(2.0 ** -5, [0.0, 1.0]), [2.0 ** -5, -2.0 ** -5, 7.0e-4, 1.45e-3]])
    assert ulp_errors(oracle.math("sin", xs), xs, mp.sin) <= 0.56
    assert ulp_errors(oracle.math("asin", xs), xs, mp.asin) <= 0.56


def test_sincosd_is_bit_identical_to_sind_and_cosd(mirror):
    xs = np.concatenate([RNG.uniform(-720, 720, 200000), np.arange(-720.0, 721.0, 15.0),
                         np.nextafter(np.arange(-360.0, 361.0, 45.0), np.inf), np.nextafter(np.arange(-360.0, 361.0, 45.0), -np.inf),
                         [0.0, -0.0, 1e-300, -1e-300, np.inf, np.nan]])
    for both, one in ((10, 0), (11, 1)):
        a, b = hm.math(mirror, both, xs), hm.math(mirror, one, xs)
        assert np.array_equal(a.view(np.uint64), b.view(np.uint64))


def test_atan_short_is_bit_identical_to_atan_below_its_threshold(mirror, oracle):
    xs = np.concatenate([RNG.uniform(-1.0e-4, 1.0e-4, 300000), 10.0 ** RNG.uniform(-300, -4, 100000),
                         -(10.0 ** RNG.uniform(-30, -4, 1000)), np.nextafter(1.0e-4, 0.0) * np.array([1.0, -1.0]),
                         [0.0, -0.0, 2.0 ** -27, 2.0 ** -28, 5e-324]])
    xs = xs[np.abs(xs) < 1.0e-4]
    a = hm.math(mirror, 12, xs)
    assert np.array_equal(a.view(np.uint64), hm.math(mirror, orc.MATH_FN["atan"], xs).view(np.uint64))
    assert np.array_equal(a.view(np.uint64), oracle.math("atan", xs).view(np.uint64))


def _haversine_rows(n, spread_deg):
    lam_x = RNG.uniform(0, 360, n)
    phi_x = RNG.uniform(-80, 90, n)
    lam_y = lam_x + RNG.uniform(-spread_deg, spread_deg, n)
    phi_y = np.clip(phi_x + RNG.uniform(-spread_deg, spread_deg, n), -90, 90)
    d2r = np.pi / 180
    r1, r2 = phi_x * d2r, phi_y * d2r
    return np.stack([lam_x, 0.5 * r1, np.cos(r1), lam_y, 0.5 * r2, np.cos(r2), np.full(n, 6371.0e3)], axis=1)


def test_tile_kernel_haversine_is_bit_identical_to_the_general_form(mirror):
    """haversine_hot (one predicate, straight-line small paths) against haversine_pre across the
    switch at 2^-6 rad half-difference (1.79 degrees), identical points and pole-like jumps."""
    rows = np.concatenate([_haversine_rows(100000, s) for s in (1e-6, 0.08, 1.0, 3.4, 3.7, 8.0, 200.0)])
    rows[:50, 3] = rows[:50, 0]; rows[:50, 4] = rows[:50, 1]; rows[:50, 5] = rows[:50, 2]   # identical points
    a, b = hm.haversine(mirror, 1, rows), hm.haversine(mirror, 0, rows)
    assert np.array_equal(a.view(np.uint64), b.view(np.uint64))
    assert (a[:50] == 0).all()


def test_tile_kernel_quadrilateral_is_bit_identical_to_the_general_form(mirror):
    """spherical_area_quadrilateral_hot (short atan when all four triangle tangents are below
    1e-4) against the general form for cells from 1e-4 to 5 degrees, incl. degenerate ones."""
    rows = []
    for size in (1e-4, 1.0 / 12, 0.25, 0.9, 1.1, 1.3, 5.0):
        n = 30000
        lat = RNG.uniform(-80, 89, n); lon = RNG.uniform(0, 360, n)
        j = lambda: RNG.uniform(-0.1, 0.1, n) * size
        rows.append(np.stack([lat + j(), lon + j(), lat + j(), lon + size + j(), lat + size + j(), lon + size + j(),
                              lat + size + j(), lon + j()], axis=1))
    rows = np.concatenate(rows)
    rows[:20, 2:4] = rows[:20, 0:2]                                      # two coincident corners
    a, b = hm.quadrilateral(mirror, 1, rows), hm.quadrilateral(mirror, 0, rows)
    assert np.array_equal(a.view(np.uint64), b.view(np.uint64))


def test_fused_asinh_sinh_cosh_equals_the_separate_functions(mirror, oracle):
    """The table kernel derives psi = asinh(t), sinh(psi), cosh(psi) from one double-double
    exponential; the doubles must be those of the oracle's three separate functions."""
    ts = np.concatenate([10.0 ** RNG.uniform(-9, 2.5, 60000), -(10.0 ** RNG.uniform(-3, 2, 2000)), [0.0, -0.0, 1.0, 35.7]])
    psi = hm.math(mirror, 13, ts)
    assert np.array_equal(psi.view(np.uint64), oracle.math("asinh", ts).view(np.uint64))
    assert np.array_equal(hm.math(mirror, 14, ts).view(np.uint64), oracle.math("sinh", psi).view(np.uint64))
    assert np.array_equal(hm.math(mirror, 15, ts).view(np.uint64), oracle.math("cosh", psi).view(np.uint64))


def test_specialised_unit_vector_trig_is_bit_identical(mirror):
    """sincosd_lat_spec (|x| <= 90) and sincosd_lon_spec (0 <= x < 360) -- the tile kernel's
    straight-line forms -- against sind / cosd, inside their ranges, at every octant boundary and
    (through the `bad` fallback) outside."""
    edges = np.arange(-360.0, 361.0, 45.0)
    lat = np.concatenate([RNG.uniform(-90, 90, 200000), edges, np.nextafter(edges, np.inf), np.nextafter(edges, -np.inf),
                          [0.0, -0.0, 1e-300, -1e-300, 90.0, -90.0, 91.0, np.nan, np.inf]])
    lon = np.concatenate([RNG.uniform(0, 360, 200000), edges, np.nextafter(edges, np.inf), np.nextafter(edges, -np.inf),
                          [0.0, -0.0, 1e-300, 359.99999999999994, 360.0, -1.0, np.nan]])
    for xs, (fs, fc) in ((lat, (16, 17)), (lon, (18, 19))):
        assert np.array_equal(hm.math(mirror, fs, xs).view(np.uint64), hm.math(mirror, 0, xs).view(np.uint64))
        assert np.array_equal(hm.math(mirror, fc, xs).view(np.uint64), hm.math(mirror, 1, xs).view(np.uint64))


def test_inline_cosine_of_the_point_stage_is_bit_identical(mirror, oracle):
    """cos_rad_inl (the tile kernel's cos(deg2rad(phi))) takes the direct cosine kernel below the HIGH WORD of pi/4 and
    the n = 0 reduction between that bound and pi/4: the same bits as cos_rad / the oracle over the latitude range,
    in particular on the interval between the two bounds and at its ends."""
    pio4 = np.float64(float.fromhex("0x1.921fb54442d18p-1"))
    lo = np.array([0x3FE921FB00000000], dtype=np.uint64).view(np.float64)[0]
    band = np.linspace(lo, pio4, 20001)
    edge = np.array([lo, pio4, np.nextafter(lo, 0), np.nextafter(lo, 1), np.nextafter(pio4, 0), np.nextafter(pio4, 1)])
    xs = np.concatenate([RNG.uniform(-np.pi / 2, np.pi / 2, 300000), band, -band, edge, -edge, [0.0, -0.0, np.pi / 2, -np.pi / 2, 1e-300]])
    got = hm.math(mirror, 22, xs)
    assert np.array_equal(got.view(np.uint64), hm.math(mirror, orc.MATH_FN["cos"], xs).view(np.uint64))
    assert np.array_equal(got.view(np.uint64), oracle.math("cos", xs).view(np.uint64))


def test_table_driven_atan_is_bit_identical(mirror, oracle):
    """atan_lut (coordinate kernel) against atan_f / the oracle over all reduction ranges, at the
    range bounds, at the zero-numerator arguments 0.5, 1, 1.5 and at 0, inf, nan."""
    b = np.array([0.4375, 0.6875, 1.1875, 2.4375, 0.5, 1.0, 1.5, 2.0 ** 66])
    xs = np.concatenate([RNG.uniform(-6, 6, 300000), 10.0 ** RNG.uniform(-12, 25, 100000), -(10.0 ** RNG.uniform(-3, 3, 20000)),
                         b, -b, np.nextafter(b, np.inf), np.nextafter(b, -np.inf), [0.0, -0.0, np.inf, -np.inf, np.nan, 1e300]])
    a = hm.math(mirror, 20, xs)
    assert np.array_equal(a.view(np.uint64), hm.math(mirror, orc.MATH_FN["atan"], xs).view(np.uint64))
    assert np.array_equal(hm.math(mirror, 21, xs).view(np.uint64), a.view(np.uint64))      # straight-line form + its fallback
    fin = ~np.isnan(xs)
    assert np.array_equal(a[fin].view(np.uint64), oracle.math("atan", xs[fin]).view(np.uint64))
