// csrc/ossg_tripolar_core.cuh -- per-point / per-cell formulas and index maps of the tripolar
// constructor, shared by all kernels in ossg_tripolar.cu.
//
// Reference rows implemented here:
//   C0  1-D coordinates                      src/tripolar_grid.jl:76,90,95-97
//   C1  cofocal ellipse/hyperbola map        src/generate_tripolar_coordinates.jl:53-89
//   C2  circshift by Nx/4                    src/tripolar_grid.jl:121-130
//   C3  coordinate halos (zipper +1, periodic x, south halo stays 0)   src/tripolar_grid.jl:137-199
//   M1  haversine spacings / quadrilateral areas                       src/tripolar_grid_utils.jl:4-45
//   M2  metric halos (zipper +1 incl. row-Ny overwrite, periodic x)    src/tripolar_grid.jl:230-273
//   M3  south continuation with lat-lon metrics                        src/tripolar_grid.jl:287-300,336-357
//
// The reference materialises C2/C3/M2 as array passes (circshift, 20 fill_halo_regions!, 12
// deepcopy).  Here they are index maps: every element of every padded output array is a pure
// function of a *source* interior index, so a halo element is evaluated at its source -- which
// yields the same bits as copying -- and a rank can build any window of the global padded
// arrays without communication.
#pragma once

#include "ossg_math.cuh"

namespace ossg {

// staggering ids, in the array order of the C ABI (cc, fc, cf, ff)
enum Loc : int { CC = 0, FC = 1, CF = 2, FF = 3 };
OSSG_HD int loc_x(int loc) { return loc & 1; }        // 1 = Face in x
OSSG_HD int loc_y(int loc) { return (loc >> 1) & 1; } // 1 = Face in y

struct GridDims {
    int Nx, Ny, Hx, Hy;
};

// 1-D tables produced by the table kernel (device memory, doubles):
//   xs[0..4)  : a*sind(lamF), a*cosd(lamF), a*sind(lamC), a*cosd(lamC), each Nx long, indexed by
//               the POST-circshift i-1 (value taken at the pre-shift index mod1(i - Nx/4, Nx))
//   ys[0..4)  : cosh(psiF), sinh(psiF), cosh(psiC), sinh(psiC), each Ny long, indexed by j-1
//   ll        : (Hy+1) rows x 9 lat-lon metrics for j = 1-Hy .. 1 (row r = j + Hy - 1)
struct Tables {
    const double* asF; const double* acF; const double* asC; const double* acC;
    const double* chF; const double* shF; const double* chC; const double* shC;
    const double* ll;
};
enum LLCol : int { LL_DX_FC = 0, LL_DX_CF, LL_DX_FF, LL_DX_CC, LL_DY, LL_AZ_FC, LL_AZ_CF, LL_AZ_FF, LL_AZ_CC, LL_NCOL };

OSSG_HOSTDEV size_t tables_doubles(int Nx, int Ny, int Hy) { return 4 * (size_t)Nx + 4 * (size_t)Ny + (size_t)LL_NCOL * (Hy + 1); }

// ---------------------------------------------------------------- C0: 1-D coordinates
// longitude = (-180, 180), Periodic, regular: faces -180 + (i-1) 360/Nx, centres half a cell on
OSSG_HD double lam_face(int i, int Nx)
{
    return m::rational_point(-180.0, (double)(Nx - (i - 1)), 180.0, (double)(i - 1), (double)Nx);
}
OSSG_HD double lam_center(int i, int Nx)
{
    return m::rational_point(-180.0, (double)(2 * Nx - (2 * i - 1)), 180.0, (double)(2 * i - 1), (double)(2 * Nx));
}
// range(south, 90, length = Ny)[j]
OSSG_HD double phi_center(int j, int Ny, double south)
{
    return m::rational_point(south, (double)(Ny - j), 90.0, (double)(j - 1), (double)(Ny - 1));
}
OSSG_HD double phi_face(int j, int Ny, double south)
{
    const double dphi = phi_center(2, Ny, south) - phi_center(1, Ny, south);   // tripolar_grid.jl:96
    return phi_center(j, Ny, south) - dphi / 2;                                 // :97
}
OSSG_HD double focal_distance(double north_poles_latitude)
{
    return m::tand((90.0 - north_poles_latitude) / 2);                          // :76
}
// pre-circshift index of post-shift i: new[i] = old[mod1(i - Nx/4, Nx)]        (:121-130)
OSSG_HD int preshift(int i, int Nx)
{
    const int is = i - Nx / 4;
    return is < 1 ? is + Nx : is;
}

// (external) LatitudeLongitudeGrid(size, latitude=(south,90), longitude=(-180,180), halo, radius)
// precomputed metrics of row j in [1-Hy, 1]; Oceananigans' precompute kernel starts at row
// 2-Hy, so row 1-Hy of the x-spacings and areas keeps its initial 0.
OSSG_HD void latlon_row(int j, int Nx, int Ny, int Hy, double south, double R, double* row /*[LL_NCOL]*/)
{
    const double pi = m::PI_F64;
    const double d2r = pi / 180.0;
    const double dlam = 360.0 / (double)Nx;
    const double dphi = (90.0 - south) / (double)Ny;
    row[LL_DY] = R * (dphi * d2r);
    if (j == 1 - Hy) {
        row[LL_DX_FC] = row[LL_DX_CF] = row[LL_DX_FF] = row[LL_DX_CC] = 0.0;
        row[LL_AZ_FC] = row[LL_AZ_CF] = row[LL_AZ_FF] = row[LL_AZ_CC] = 0.0;
        return;
    }
    const double phiF_j  = m::rational_point(south, (double)(Ny - (j - 1)), 90.0, (double)(j - 1), (double)Ny);
    const double phiF_j1 = m::rational_point(south, (double)(Ny - j), 90.0, (double)j, (double)Ny);
    const double phiC_j  = m::rational_point(south, (double)(2 * Ny - (2 * j - 1)), 90.0, (double)(2 * j - 1), (double)(2 * Ny));
    const double phiC_jm = m::rational_point(south, (double)(2 * Ny - (2 * j - 3)), 90.0, (double)(2 * j - 3), (double)(2 * Ny));
    const double rdl = R * (dlam * d2r);
    row[LL_DX_FC] = rdl * m::cos_rad(pi * phiC_j / 180);
    row[LL_DX_CC] = row[LL_DX_FC];
    row[LL_DX_CF] = rdl * m::cos_rad(pi * phiF_j / 180);
    row[LL_DX_FF] = row[LL_DX_CF];
    const double r2dl = R * R * (dlam * d2r);
    row[LL_AZ_FC] = r2dl * (m::sin_rad(pi * phiF_j1 / 180) - m::sin_rad(pi * phiF_j / 180));
    row[LL_AZ_CC] = row[LL_AZ_FC];
    row[LL_AZ_CF] = r2dl * (m::sin_rad(pi * phiC_j / 180) - m::sin_rad(pi * phiC_jm / 180));
    row[LL_AZ_FF] = row[LL_AZ_CF];
}

// ---------------------------------------------------------------- C1: one staggering point
// as_i = a*sind(lam1D[ip]), ac_i = a*cosd(lam1D[ip]) (ip = pre-shift index), ch_j/sh_j = cosh/sinh psi_j
//
// Split in two so that the coordinate kernel can exploit the mirror symmetry of the map:
// lam1D[ip'] = -lam1D[ip] (ip' = Nx+2-ip for faces, Nx+1-ip for centres) gives x' = -x, y' = y
// EXACTLY (sind is odd, cosd even, to the last bit), hence atan(y'/x') = -atan(y/x) and the same
// phi; only the cheap tail (hemisphere shift, pole shift, wrap) differs between the two points.
struct PointCore {
    double lam_a;     // -(180/pi) atan(y/x)             (:77, before the pole special case)
    double phi;       // 90 - (360/pi) atan(sqrt(y^2+x^2)) (:78)
    bool pole;        // on_the_north_pole                (:74)
};

OSSG_HD PointCore tripolar_core(double as_i, double ac_i, double ch_j, double sh_j)
{
    const double x = as_i * ch_j;                                             // :67
    const double y = ac_i * sh_j;                                             // :68
    constexpr double m180_pi = -180.0 / m::PI_F64;
    constexpr double p360_pi = 360.0 / m::PI_F64;
    PointCore c;
    c.pole = (x == 0.0) & (y == 0.0);                                         // :74
    c.lam_a = m180_pi * m::atan_f(y / x);                                     // :77
    c.phi = 90.0 - p360_pi * m::atan_f(sqrt(y * y + x * x));                  // :78
    return c;
}

// the same point for the coordinate kernel's hot loop: table-driven atan, branch-free sqrt unless
// x = y = 0 (the poles) or tiny
OSSG_HD PointCore tripolar_core_lut(double as_i, double ac_i, double ch_j, double sh_j, const double* atan_table)
{
    const double x = as_i * ch_j;                                             // :67
    const double y = ac_i * sh_j;                                             // :68
    constexpr double m180_pi = -180.0 / m::PI_F64;
    constexpr double p360_pi = 360.0 / m::PI_F64;
    PointCore c;
    c.pole = (x == 0.0) & (y == 0.0);                                         // :74
    c.lam_a = m180_pi * m::atan_lut(y / x, atan_table);                       // :77
    const double rho2 = y * y + x * x;
    const double rho = (m::hi_abs(rho2) >= m::SQRT_NB_MIN_HI && rho2 < 0x1p1000) ? m::sqrt_nb(rho2) : sqrt(rho2);
    c.phi = 90.0 - p360_pi * m::atan_lut(rho, atan_table);                    // :78
    return c;
}

// Speculative straight-line forms (no branch): the same bits as tripolar_core / tripolar_finish
// whenever `bad` stays false; a bad point (x or y zero or tiny: pole row, pole meridians;
// reduction-boundary arguments; |lambda| >= 360 before the wrap) is recomputed by the caller
// through the general functions.
OSSG_HD PointCore tripolar_core_sl(double as_i, double ac_i, double ch_j, double sh_j, const double* atan_table, bool& bad)
{
    const double x = as_i * ch_j;                                             // :67
    const double y = ac_i * sh_j;                                             // :68
    constexpr double m180_pi = -180.0 / m::PI_F64;
    constexpr double p360_pi = 360.0 / m::PI_F64;
    PointCore c;
    c.pole = false;                                                           // x = y = 0 is a bad point
    bad |= (m::hi_abs(y) < m::DIV_NB_MIN_HI) | (m::hi_abs(x) - 0x39B00000u >= 0x0C800000u);   // |y| >= 2^-969, 2^-100 <= |x| < 2^100
    c.lam_a = m180_pi * m::atan_sl(m::div_nb(y, x), atan_table, bad);         // :77
    const double rho2 = y * y + x * x;                                        // (>= 2^-200 here)
    c.phi = 90.0 - p360_pi * m::atan_sl(m::sqrt_nb(rho2), atan_table, bad);   // :78
    return c;
}
OSSG_HD double tripolar_finish_sl(double lam_a, bool ip_le_half, double fpl90, bool& bad)
{
    double l = lam_a + (ip_le_half ? -90.0 : 90.0);                           // :82 (not a pole: :77 took lam_a)
    l += fpl90;                                                               // :86, fpl90 = first_pole_longitude + 90
    bad |= !(fabs(l) < 360.0);                                                // then fmod(l, 360) == l and l + 360 < 720
    const double r = l + 360.0;
    return (r < 360.0) ? r : (r - 360.0);
}

// tail of the point: pole value (:75-77), hemisphere shift (:82), pole-longitude shift (:86), wrap (:87)
OSSG_HD double tripolar_finish(double lam_a, bool pole, bool ip_is_one, bool ip_le_half, double first_pole_longitude)
{
    const double north_pole_value = ip_is_one ? -90.0 : 90.0;                 // :75
    double l = pole ? north_pole_value : lam_a;                               // :77
    l += ip_le_half ? -90.0 : 90.0;                                           // :82
    l += first_pole_longitude + 90.0;                                         // :86
    // convert_to_0_360, src/OrthogonalSphericalShellGrids.jl:24 (fmod(l,360) == l when |l| < 360;
    // r in [360, 720): r - 360 is exact by Sterbenz and equals fmod(r, 360))
    double r = (fabs(l) < 360.0) ? l : fmod(l, 360.0);
    r += 360.0;
    return (r < 360.0) ? r : (r < 720.0) ? (r - 360.0) : fmod(r, 360.0);
}

OSSG_HD void tripolar_point(double as_i, double ac_i, double ch_j, double sh_j, bool ip_is_one,
                            bool ip_le_half, double first_pole_longitude, double& lam, double& phi)
{
    const PointCore c = tripolar_core(as_i, ac_i, ch_j, sh_j);
    phi = c.phi;
    lam = tripolar_finish(c.lam_a, c.pole, ip_is_one, ip_le_half, first_pole_longitude);
}

OSSG_HD_NOINLINE void tripolar_point_ool(double as_i, double ac_i, double ch_j, double sh_j, bool ip_is_one,
                                         bool ip_le_half, double first_pole_longitude, double& lam, double& phi)
{
    tripolar_point(as_i, ac_i, ch_j, sh_j, ip_is_one, ip_le_half, first_pole_longitude, lam, phi);
}

// ---------------------------------------------------------------- C3 / M2 index maps
// Source of the padded element (I, J) of a field at (locx, locy) after
//   zipper north fill over i in 1..Nx (src/zipper_boundary_condition.jl:70-138, sign handled by
//   the caller), then periodic x fill over every row (external Oceananigans ordering).
// Returns false for the south halo (J < 1), which neither fill touches.
// `wrapped` is set when the x-Face fold source wrapped around (i' > Nx): the reference takes
// |sign| there (zipper_boundary_condition.jl:74-75, 91-92).
OSSG_HD bool halo_source(int locx, int locy, int I, int J, int Nx, int Ny, int& i, int& j, bool& wrapped)
{
    wrapped = false;
    if (J < 1) return false;
    i = I < 1 ? I + Nx : (I > Nx ? I - Nx : I);
    j = J;
    const bool north = J > Ny;
    const bool redundant = (J == Ny) && (locy == 0) && (i > Nx / 2);   // row-Ny overwrite, :102,135
    if (north || redundant) {
        int ip = locx ? Nx - i + 2 : Nx - i + 1;
        if (ip > Nx) { ip -= Nx; wrapped = true; }
        i = ip;
        if (north) {
            const int d = J - Ny;
            j = locy ? Ny - d + 1 : Ny - d;
        }
    }
    return true;
}

// coordinate value (lam, phi) of staggering `loc` at padded global index (I, J), straight from
// the tables (C1 + C2 + C3).  South halo: 0.
OSSG_HD void coord_at(const Tables& tb, const GridDims& g, double first_pole_longitude, int loc, int I, int J,
                      double& lam, double& phi)
{
    int i, j; bool wrapped;
    if (!halo_source(loc_x(loc), loc_y(loc), I, J, g.Nx, g.Ny, i, j, wrapped)) { lam = 0.0; phi = 0.0; return; }
    const int ip = preshift(i, g.Nx);
    const bool fx = loc_x(loc) != 0, fy = loc_y(loc) != 0;
    const double as_i = fx ? tb.asF[i - 1] : tb.asC[i - 1];
    const double ac_i = fx ? tb.acF[i - 1] : tb.acC[i - 1];
    const double ch_j = fy ? tb.chF[j - 1] : tb.chC[j - 1];
    const double sh_j = fy ? tb.shF[j - 1] : tb.shC[j - 1];
    tripolar_point_ool(as_i, ac_i, ch_j, sh_j, ip == 1, ip <= g.Nx / 2, first_pole_longitude, lam, phi);
}

// ---------------------------------------------------------------- M1 formulas
// (external) Distances.jl 0.10 Haversine()(x, y): x = (lam, phi) of the first point.
//     dlam = deg2rad(lam_y - lam_x); dphi = phi2 - phi1 (radians)
//     a = sin(dphi/2)^2 + cos(phi1) cos(phi2) sin(dlam/2)^2;  2 (R asin(min(sqrt(a), 1)))
// Callers pass hphi = deg2rad(phi)/2 and c = cos(deg2rad(phi)) (hoisted per point by the tile
// kernel).  Halving is exact, so hphi2 - hphi1 == (phi2 - phi1)/2 and
// (lam_y - lam_x) * (D2R/2) == ((lam_y - lam_x) * D2R)/2 bit for bit.
constexpr double D2R = m::PI_F64 / 180.0;
constexpr double HALF_D2R = D2R * 0.5;

OSSG_HD double haversine_tail(double hdphi, double hdlam, double c1, double c2, double radius)
{
    const double s1 = m::sin_rad(hdphi);
    const double s2 = m::sin_rad(hdlam);
    const double a = s1 * s1 + c1 * c2 * (s2 * s2);
    double r = sqrt(a);
    if (!(r < 1.0)) r = (r != r) ? r : 1.0;                 // min(sqrt(a), 1)
    return 2 * (radius * m::asin_f(r));
}
OSSG_HD_NOINLINE double haversine_tail_ool(double hdphi, double hdlam, double c1, double c2, double radius)
{
    return haversine_tail(hdphi, hdlam, c1, c2, radius);
}

OSSG_HD double haversine_pre(double lam_x, double hphi1, double c1, double lam_y, double hphi2, double c2, double radius)
{
    return haversine_tail(hphi2 - hphi1, (lam_y - lam_x) * HALF_D2R, c1, c2, radius);
}

// The same value for the tile kernel.  Both half-differences below 2^-6 (spacings under ~1.8
// degrees): both sines take their small path, a <= 2^-12 (1 + c1 c2) keeps sqrt(a) below 2^-5, so
// min(.,1) is the identity and asin takes its small path too -- one predicate, straight-line code.
// Anything else (pole neighbourhoods, coarse grids, NaN) goes through the general function.
constexpr double HAV_HOT = 0x1p-6;
OSSG_HD double haversine_hot(double lam_x, double hphi1, double c1, double lam_y, double hphi2, double c2, double radius)
{
    const double hdphi = hphi2 - hphi1;
    const double hdlam = (lam_y - lam_x) * HALF_D2R;
    if ((fabs(hdphi) < HAV_HOT) & (fabs(hdlam) < HAV_HOT)) {
        const double s1 = m::sin_small(hdphi);
        const double s2 = m::sin_small(hdlam);
        const double a = s1 * s1 + c1 * c2 * (s2 * s2);
        return (2 * radius) * m::asin_small(sqrt(a));     // == 2 (R asin): doubling is exact
    }
    return haversine_tail_ool(hdphi, hdlam, c1, c2, radius);
}

// Preconditions of the speculative forms below, accumulated over a whole cell.  An FP64 comparison costs the issue
// slot twice (half-rate pipe), so the range tests run on the HIGH WORDS with integer max / min (one VIMNMX3 per two
// values) and are judged once per cell:
//   squares z of the haversine half-differences: unsigned max; z < 2^-12  <=>  hi(z) < 0x3F300000 for z >= 0, which
//     implies |x| < 2^-6 = HAV_HOT (|x| >= 2^-6 squares to at least 2^-12 exactly); a NaN or infinite z has a high
//     word of at least 0x7FF00000 under either sign;
//   haversine radicands a: signed min; a >= 2^-970 (sqrt_nb), a negative radicand has a negative high word;
//   triangle numerators n = |.| : unsigned min (div_nb); tangents t: unsigned max, hi(t) < hi(ATAN_SHORT) implies
//     0 <= t < ATAN_SHORT, a negative or NaN tangent (never a real cell) compares high and takes the general path.
// -DOSSG_RANGE_FP_TESTS keeps the round-1 form (FP64 comparisons ORed into a flag).
struct SpecRange {
#if defined(OSSG_RANGE_FP_TESTS)
    bool bad = false;
    OSSG_HD void haversine(double hdphi, double hdlam, double, double, double a)
    {
        bad |= !((fabs(hdphi) < 0x1p-6) & (fabs(hdlam) < 0x1p-6)) | (m::hi_abs(a) < m::SQRT_NB_MIN_HI);
    }
    OSSG_HD void triangles(double n0, double n1, double n2, double n3, double t0, double t1, double t2, double t3)
    {
        bad |= (m::hi_abs(n0) < m::DIV_NB_MIN_HI) | (m::hi_abs(n1) < m::DIV_NB_MIN_HI) |
               (m::hi_abs(n2) < m::DIV_NB_MIN_HI) | (m::hi_abs(n3) < m::DIV_NB_MIN_HI);
        // (a tiny or non-finite denominator makes |t| huge or NaN and fails this test as well)
        bad |= !((fabs(t0) < m::ATAN_SHORT) & (fabs(t1) < m::ATAN_SHORT) & (fabs(t2) < m::ATAN_SHORT) & (fabs(t3) < m::ATAN_SHORT));
    }
    OSSG_HD void flag(bool b) { bad |= b; }
    OSSG_HD bool violated() const { return bad; }
#else
    unsigned zmax = 0u, tmax = 0u, nmin = 0xffffffffu;
    int amin = 0x7fffffff;
    OSSG_HD static unsigned umax(unsigned a, unsigned b) { return a > b ? a : b; }
    OSSG_HD static unsigned umin(unsigned a, unsigned b) { return a < b ? a : b; }
    OSSG_HD void haversine(double, double, double z1, double z2, double a)
    {
        zmax = umax(zmax, umax(m::hi_bits(z1), m::hi_bits(z2)));
        const int ah = (int)m::hi_bits(a);
        amin = ah < amin ? ah : amin;
    }
    OSSG_HD void triangles(double n0, double n1, double n2, double n3, double t0, double t1, double t2, double t3)
    {
        nmin = umin(umin(nmin, m::hi_bits(n0)), umin(m::hi_bits(n1), umin(m::hi_bits(n2), m::hi_bits(n3))));   // n = |.|: no sign bit
        tmax = umax(umax(tmax, m::hi_bits(t0)), umax(m::hi_bits(t1), umax(m::hi_bits(t2), m::hi_bits(t3))));
    }
    OSSG_HD void flag(bool b) { zmax = b ? 0xffffffffu : zmax; }
    OSSG_HD bool violated() const
    {
        return (zmax >= 0x3F300000u) | (amin < (int)m::SQRT_NB_MIN_HI) | (nmin < m::DIV_NB_MIN_HI) | (tmax >= ATAN_SHORT_HI);
    }
    static constexpr unsigned ATAN_SHORT_HI = 0x3F1A36E2u;     // high word of ATAN_SHORT = 1e-4 (0x3F1A36E2EB1C432D)
#endif
};

// Speculative form for straight-line code: always evaluates the hot sequence (with the branch-free
// sqrt) and records in `rg` whether its preconditions failed; the caller recomputes the cell through
// the general functions when anything was violated.  Same bits as haversine_pre otherwise.
OSSG_HD double haversine_spec(double lam_x, double hphi1, double c1, double lam_y, double hphi2, double c2,
                              double radius2, SpecRange& rg)
{
    const double hdphi = hphi2 - hphi1;
    const double hdlam = (lam_y - lam_x) * HALF_D2R;
    const double s1 = m::sin_small(hdphi);
    const double s2 = m::sin_small(hdlam);
    const double a = s1 * s1 + c1 * c2 * (s2 * s2);
    rg.haversine(hdphi, hdlam, hdphi * hdphi, hdlam * hdlam, a);      // (the squares are sin_small's own z)
    return radius2 * m::asin_small(m::sqrt_nb(a));        // radius2 = 2 R: == 2 (R asin), doubling is exact
}
template <class P>   // P: { lam, hphi, cosphi } of a staged patch point (csrc/ossg_tripolar.cu)
OSSG_HD double haversine_spec(const P& x, const P& y, double radius2, SpecRange& rg)
{
    return haversine_spec(x.lam, x.hphi, x.cosphi, y.lam, y.hphi, y.cosphi, radius2, rg);
}

struct Vec3 { double x, y, z; };

// (external) Oceananigans.Grids.lat_lon_to_cartesian(lat, lon, 1) =
//     (cosd(lon) cosd(lat), sind(lon) cosd(lat), sind(lat))
OSSG_HD Vec3 lat_lon_to_cartesian(double lat, double lon)
{
    double sin_lat, cos_lat, sin_lon, cos_lon;
    m::sincosd(lat, sin_lat, cos_lat);
    m::sincosd(lon, sin_lon, cos_lon);
    Vec3 v;
    v.x = cos_lon * cos_lat;
    v.y = sin_lon * cos_lat;
    v.z = sin_lat;
    return v;
}
// straight-line form for the tile kernel: valid for |lat| <= 90 and 0 <= lon < 360, otherwise sets `bad`
OSSG_HD Vec3 lat_lon_to_cartesian_spec(double lat, double lon, bool& bad)
{
    double sin_lat, cos_lat, sin_lon, cos_lon;
    m::sincosd_lat_spec(lat, sin_lat, cos_lat, bad);
    m::sincosd_lon_spec(lon, sin_lon, cos_lon, bad);
    Vec3 v;
    v.x = cos_lon * cos_lat;
    v.y = sin_lon * cos_lat;
    v.z = sin_lat;
    return v;
}
OSSG_HD double dot3(const Vec3& a, const Vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
OSSG_HD Vec3 cross3(const Vec3& a, const Vec3& b)
{
    Vec3 c;
    c.x = a.y * b.z - a.z * b.y;
    c.y = a.z * b.x - a.x * b.z;
    c.z = a.x * b.y - a.y * b.x;
    return c;
}
// (external) Oceananigans.Grids.spherical_area_quadrilateral: half the sum of the four triangles
// (a,b,c), (a,b,d), (a,c,d), (b,c,d); triangle area = 2 atan(|a1.(a2 x a3)| / (1 + a1.a2 + a2.a3 + a1.a3)).
// The six pair dot products and three cross products are shared between the triangles; each
// value is produced by the same expression as in the unshared form.
struct QuadTans { double abc, abd, acd, bcd; };
OSSG_HD QuadTans quadrilateral_tangents(const Vec3& a, const Vec3& b, const Vec3& c, const Vec3& d)
{
    const double ab = dot3(a, b), ac = dot3(a, c), ad = dot3(a, d);
    const double bc = dot3(b, c), bd = dot3(b, d), cd = dot3(c, d);
    const Vec3 bxc = cross3(b, c), bxd = cross3(b, d), cxd = cross3(c, d);
    QuadTans t;
    t.abc = fabs(dot3(a, bxc)) / (1 + ab + bc + ac);
    t.abd = fabs(dot3(a, bxd)) / (1 + ab + bd + ad);
    t.acd = fabs(dot3(a, cxd)) / (1 + ac + cd + ad);
    t.bcd = fabs(dot3(b, cxd)) / (1 + bc + cd + bd);
    return t;
}
OSSG_HD double spherical_area_quadrilateral(const Vec3& a, const Vec3& b, const Vec3& c, const Vec3& d)
{
    const QuadTans t = quadrilateral_tangents(a, b, c, d);
    return 0.5 * (2 * m::atan_f(t.abc) + 2 * m::atan_f(t.abd) + 2 * m::atan_f(t.acd) + 2 * m::atan_f(t.bcd));
}
OSSG_HD_NOINLINE double quadrilateral_from_tangents_ool(double t_abc, double t_abd, double t_acd, double t_bcd)
{
    return 0.5 * (2 * m::atan_f(t_abc) + 2 * m::atan_f(t_abd) + 2 * m::atan_f(t_acd) + 2 * m::atan_f(t_bcd));
}
// tile-kernel form: all four tangents below ATAN_SHORT (cells under ~1.1 degrees) -> short atan
OSSG_HD double spherical_area_quadrilateral_hot(const Vec3& a, const Vec3& b, const Vec3& c, const Vec3& d)
{
    const QuadTans t = quadrilateral_tangents(a, b, c, d);
    if ((fabs(t.abc) < m::ATAN_SHORT) & (fabs(t.abd) < m::ATAN_SHORT) & (fabs(t.acd) < m::ATAN_SHORT) & (fabs(t.bcd) < m::ATAN_SHORT))
        // == 0.5 (2 atan + 2 atan + 2 atan + 2 atan): doubling and halving are exact
        return m::atan_short(t.abc) + m::atan_short(t.abd) + m::atan_short(t.acd) + m::atan_short(t.bcd);
    return quadrilateral_from_tangents_ool(t.abc, t.abd, t.acd, t.bcd);
}

// Speculative form (see haversine_spec): branch-free divisions and short atans.
OSSG_HD double spherical_area_quadrilateral_spec(const Vec3& a, const Vec3& b, const Vec3& c, const Vec3& d, SpecRange& rg)
{
    const double ab = dot3(a, b), ac = dot3(a, c), ad = dot3(a, d);
    const double bc = dot3(b, c), bd = dot3(b, d), cd = dot3(c, d);
    const Vec3 bxc = cross3(b, c), bxd = cross3(b, d), cxd = cross3(c, d);
    const double n_abc = fabs(dot3(a, bxc)), n_abd = fabs(dot3(a, bxd)), n_acd = fabs(dot3(a, cxd)), n_bcd = fabs(dot3(b, cxd));
    const double t_abc = m::div_nb(n_abc, 1 + ab + bc + ac);
    const double t_abd = m::div_nb(n_abd, 1 + ab + bd + ad);
    const double t_acd = m::div_nb(n_acd, 1 + ac + cd + ad);
    const double t_bcd = m::div_nb(n_bcd, 1 + bc + cd + bd);
    rg.triangles(n_abc, n_abd, n_acd, n_bcd, t_abc, t_abd, t_acd, t_bcd);
    return m::atan_short(t_abc) + m::atan_short(t_abd) + m::atan_short(t_acd) + m::atan_short(t_bcd);
}

// Out-of-line copies for the generic (cold) per-element path: the halo-frame kernel evaluates
// metrics_cell up to four times per element, fully inlined that is ~17k SASS instructions which
// run once each -- an instruction-cache miss per line.  Same arithmetic, one copy of the code.
OSSG_HD_NOINLINE double haversine_pre_ool(double lam_x, double hphi1, double c1, double lam_y, double hphi2, double c2, double radius)
{
    return haversine_pre(lam_x, hphi1, c1, lam_y, hphi2, c2, radius);
}
template <class P>
OSSG_HD double haversine_pre_ool(const P& x, const P& y, double radius)
{
    return haversine_pre_ool(x.lam, x.hphi, x.cosphi, y.lam, y.hphi, y.cosphi, radius);
}
OSSG_HD_NOINLINE double cos_rad_ool(double x) { return m::cos_rad(x); }
OSSG_HD_NOINLINE Vec3 lat_lon_to_cartesian_ool(double lat, double lon) { return lat_lon_to_cartesian(lat, lon); }
OSSG_HD_NOINLINE double spherical_area_quadrilateral_ool(const Vec3& a, const Vec3& b, const Vec3& c, const Vec3& d)
{
    return spherical_area_quadrilateral(a, b, c, d);
}

// metric ids in C-ABI order
enum Metric : int { DX_CC = 0, DX_FC, DX_CF, DX_FF, DY_CC, DY_CF, DY_FC, DY_FF, AZ_CC, AZ_FC, AZ_CF, AZ_FF, N_METRIC };
// staggering class of each metric (which halo map applies to it in M2)
OSSG_HD int metric_loc(int mtr)
{
    switch (mtr) {
    case DX_CC: case DY_CC: case AZ_CC: return CC;
    case DX_FC: case DY_FC: case AZ_FC: return FC;
    case DX_CF: case DY_CF: case AZ_CF: return CF;
    default: return FF;
    }
}

// M1 at interior cell (i, j) through a coordinate accessor
//   acc(loc, I, J, lam, phi)  -> padded coordinate values.
// `mask` selects which staggering classes are wanted (bit loc); unwanted outputs are left alone.
template <class Acc>
OSSG_HD void metrics_cell(const Acc& acc, int i, int j, double R, unsigned mask, double* out /*[N_METRIC]*/)
{
    double l0, p0, l1, p1, l2, p2, l3, p3;
    const bool want_cc = mask & (1u << CC), want_fc = mask & (1u << FC), want_cf = mask & (1u << CF), want_ff = mask & (1u << FF);
    // FF points: (i,j) (i+1,j) (i+1,j+1) (i,j+1)     -> dx_cf, dy_fc, az_cc
    if (want_cf || want_fc || want_cc) {
        acc(FF, i, j, l0, p0); acc(FF, i + 1, j, l1, p1); acc(FF, i + 1, j + 1, l2, p2); acc(FF, i, j + 1, l3, p3);
        const double r0 = p0 * D2R, r1 = p1 * D2R, r3 = p3 * D2R;
        const double c0 = cos_rad_ool(r0);
        if (want_cf) out[DX_CF] = haversine_pre_ool(l1, 0.5 * r1, cos_rad_ool(r1), l0, 0.5 * r0, c0, R);      // utils.jl:15
        if (want_fc) out[DY_FC] = haversine_pre_ool(l3, 0.5 * r3, cos_rad_ool(r3), l0, 0.5 * r0, c0, R);      // :19
        if (want_cc) {
            const Vec3 a = lat_lon_to_cartesian_ool(p0, l0), b = lat_lon_to_cartesian_ool(p1, l1);
            const Vec3 c = lat_lon_to_cartesian_ool(p2, l2), d = lat_lon_to_cartesian_ool(p3, l3);
            out[AZ_CC] = spherical_area_quadrilateral_ool(a, b, c, d) * (R * R);               // :23-28
        }
    }
    // CC points: (i,j) (i-1,j) (i,j-1) (i-1,j-1)     -> dx_fc, dy_cf, az_ff
    if (want_fc || want_cf || want_ff) {
        acc(CC, i, j, l0, p0); acc(CC, i - 1, j, l1, p1); acc(CC, i, j - 1, l2, p2); acc(CC, i - 1, j - 1, l3, p3);
        const double r0 = p0 * D2R, r1 = p1 * D2R, r2 = p2 * D2R;
        const double c0 = cos_rad_ool(r0);
        if (want_fc) out[DX_FC] = haversine_pre_ool(l0, 0.5 * r0, c0, l1, 0.5 * r1, cos_rad_ool(r1), R);      // :14
        if (want_cf) out[DY_CF] = haversine_pre_ool(l0, 0.5 * r0, c0, l2, 0.5 * r2, cos_rad_ool(r2), R);      // :20
        if (want_ff) {
            const Vec3 a = lat_lon_to_cartesian_ool(p3, l3), b = lat_lon_to_cartesian_ool(p2, l2);
            const Vec3 c = lat_lon_to_cartesian_ool(p0, l0), d = lat_lon_to_cartesian_ool(p1, l1);
            out[AZ_FF] = spherical_area_quadrilateral_ool(a, b, c, d) * (R * R);               // :38-43
        }
    }
    // FC points: (i,j) (i+1,j) (i,j-1)               -> dx_cc, dy_ff
    if (want_cc || want_ff) {
        acc(FC, i, j, l0, p0);
        const double r0 = p0 * D2R;
        const double c0 = cos_rad_ool(r0);
        if (want_cc) {
            acc(FC, i + 1, j, l1, p1);
            const double r1 = p1 * D2R;
            out[DX_CC] = haversine_pre_ool(l1, 0.5 * r1, cos_rad_ool(r1), l0, 0.5 * r0, c0, R);              // :13
        }
        if (want_ff) {
            acc(FC, i, j - 1, l2, p2);
            const double r2 = p2 * D2R;
            out[DY_FF] = haversine_pre_ool(l0, 0.5 * r0, c0, l2, 0.5 * r2, cos_rad_ool(r2), R);              // :21
        }
    }
    // CF points: (i,j) (i-1,j) (i,j+1)               -> dx_ff, dy_cc
    if (want_ff || want_cc) {
        acc(CF, i, j, l0, p0);
        const double r0 = p0 * D2R;
        const double c0 = cos_rad_ool(r0);
        if (want_ff) {
            acc(CF, i - 1, j, l1, p1);
            const double r1 = p1 * D2R;
            out[DX_FF] = haversine_pre_ool(l0, 0.5 * r0, c0, l1, 0.5 * r1, cos_rad_ool(r1), R);              // :16
        }
        if (want_cc) {
            acc(CF, i, j + 1, l3, p3);
            const double r3 = p3 * D2R;
            out[DY_CC] = haversine_pre_ool(l3, 0.5 * r3, cos_rad_ool(r3), l0, 0.5 * r0, c0, R);              // :18
        }
    }
    if (want_fc) out[AZ_FC] = out[DY_FC] * out[DX_FC];                                     // :34
    if (want_cf) out[AZ_CF] = out[DY_CF] * out[DX_CF];                                     // :35
}

// ONE metric of M1 at interior cell (i, j) (same expressions as metrics_cell; the frame kernel runs
// one thread per element and metric).  The product areas az_fc, az_cf evaluate their two spacings.
template <class Acc>
OSSG_HD double metric_one(const Acc& acc, int mtr, int i, int j, double R)
{
    // haversine(first point X, second point Y) between two points of one staggering
    auto hav = [&](int loc, int ix, int jx, int iy, int jy) {
        double lx, px, ly, py;
        acc(loc, ix, jx, lx, px);
        acc(loc, iy, jy, ly, py);
        const double rx = px * D2R, ry = py * D2R;
        return haversine_pre_ool(lx, 0.5 * rx, cos_rad_ool(rx), ly, 0.5 * ry, cos_rad_ool(ry), R);
    };
    auto quad = [&](int loc, int ia, int ja, int ib, int jb, int ic, int jc, int id, int jd) {
        double l, p;
        acc(loc, ia, ja, l, p); const Vec3 a = lat_lon_to_cartesian_ool(p, l);
        acc(loc, ib, jb, l, p); const Vec3 b = lat_lon_to_cartesian_ool(p, l);
        acc(loc, ic, jc, l, p); const Vec3 c = lat_lon_to_cartesian_ool(p, l);
        acc(loc, id, jd, l, p); const Vec3 d = lat_lon_to_cartesian_ool(p, l);
        return spherical_area_quadrilateral_ool(a, b, c, d) * (R * R);
    };
    switch (mtr) {
    case DX_CC: return hav(FC, i + 1, j, i, j);                              // utils.jl:13
    case DX_FC: return hav(CC, i, j, i - 1, j);                              // :14
    case DX_CF: return hav(FF, i + 1, j, i, j);                              // :15
    case DX_FF: return hav(CF, i, j, i - 1, j);                              // :16
    case DY_CC: return hav(CF, i, j + 1, i, j);                              // :18
    case DY_FC: return hav(FF, i, j + 1, i, j);                              // :19
    case DY_CF: return hav(CC, i, j, i, j - 1);                              // :20
    case DY_FF: return hav(FC, i, j, i, j - 1);                              // :21
    case AZ_CC: return quad(FF, i, j, i + 1, j, i + 1, j + 1, i, j + 1);     // :23-28
    case AZ_FC: return hav(FF, i, j + 1, i, j) * hav(CC, i, j, i - 1, j);    // :34  dy_fc * dx_fc
    case AZ_CF: return hav(CC, i, j, i, j - 1) * hav(FF, i + 1, j, i, j);    // :35  dy_cf * dx_cf
    default:    return quad(CC, i - 1, j - 1, i, j - 1, i, j, i - 1, j);     // :38-43  az_ff
    }
}

// Full value of the 12 metrics at padded global element (I, J): M1 evaluated at each class's M2
// source, or the lat-lon row for J <= 1 (M3).  Generic (slow) path used for halo elements and by
// tests; the tile kernel handles the bulk.
template <class Acc>
OSSG_HD void metrics_at_padded(const Acc& acc, const Tables& tb, const GridDims& g, double R, int I, int J,
                               double* out /*[N_METRIC]*/)
{
    if (J <= 1) {
        const double* row = tb.ll + (size_t)(J + g.Hy - 1) * LL_NCOL;
        out[DX_FF] = row[LL_DX_FF]; out[DX_FC] = row[LL_DX_FC]; out[DX_CF] = row[LL_DX_CF]; out[DX_CC] = row[LL_DX_CC];
        out[DY_FF] = row[LL_DY];    out[DY_FC] = row[LL_DY];    out[DY_CF] = row[LL_DY];    out[DY_CC] = row[LL_DY];
        out[AZ_FF] = row[LL_AZ_FF]; out[AZ_FC] = row[LL_AZ_FC]; out[AZ_CF] = row[LL_AZ_CF]; out[AZ_CC] = row[LL_AZ_CC];
        return;
    }
    int si[4], sj[4];
    for (int loc = 0; loc < 4; ++loc) {
        bool wrapped;
        halo_source(loc_x(loc), loc_y(loc), I, J, g.Nx, g.Ny, si[loc], sj[loc], wrapped);
    }
    unsigned done = 0;
    for (int loc = 0; loc < 4; ++loc) {
        if (done & (1u << loc)) continue;
        unsigned mask = 0;
        for (int l2 = loc; l2 < 4; ++l2)
            if (!(done & (1u << l2)) && si[l2] == si[loc] && sj[l2] == sj[loc]) mask |= 1u << l2;
        metrics_cell(acc, si[loc], sj[loc], R, mask, out);
        done |= mask;
    }
}

}  // namespace ossg
