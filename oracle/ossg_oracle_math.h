/*
 * oracle/ossg_oracle_math.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Elementary functions used by the CPU oracle (oracle/ossg_oracle.c).
 *
 * The reference calls Julia Base (>= 1.10, un-vendored, not pinned: there is no
 * Manifest.toml in /root/reference) for sind/cosd/tand/atan/asin/asinh/sinh/cosh/
 * sin/cos.  Julia's sind/cosd/atan/asin are ports of the public fdlibm kernels
 * (__kernel_sin, __kernel_cos, s_atan.c, e_asin.c); they are restated here with the
 * fdlibm coefficient sets.  Julia's results are *not* reproducible bit-for-bit in this
 * environment (no Julia; its exp/log/sinh/cosh are Julia-specific tables), so this
 * oracle defines its own deterministic evaluation order: every operation below is an
 * IEEE-754 correctly-rounded primitive (+ - * / sqrt fma) written out explicitly, and
 * the file must be compiled with -ffp-contract=off so that no other contraction occurs.
 * The CUDA product (csrc/ossg_math.cuh) states the same sequences independently; the
 * two are expected to agree bit-for-bit, and tests/ checks both against glibc libm and
 * 50-digit mpmath.
 *
 * "Hot" functions (evaluated per grid point): sin/cos kernels, sind, cosd, sin, cos,
 * atan, asin.  "Cold" functions (evaluated 2(Nx+Ny) times per grid: tand, asinh, sinh,
 * cosh) are computed in double-double arithmetic and rounded once, i.e. they are
 * correctly rounded except with vanishing probability.
 *
 * Build with -DOSSG_ORACLE_LIBM to route everything through glibc libm instead (used
 * for the CPU timing baseline and for the libm-sensitivity test).
 */
#ifndef OSSG_ORACLE_MATH_H
#define OSSG_ORACLE_MATH_H

#include <math.h>
#include <stdint.h>

/* ------------------------------------------------------------------ constants */
#define OM_DEG2RAD_HI 0x1.1df46a2529d39p-6   /* fl(pi/180)                      */
#define OM_DEG2RAD_LO 0x1.5c1d8becdd291p-62  /* pi/180 - OM_DEG2RAD_HI          */
#define OM_INVPIO2    0x1.45f306dc9c883p-1
#define OM_PIO2_1     0x1.921fb54400000p+0   /* first 33 bits of pi/2           */
#define OM_PIO2_2     0x1.0b4611a600000p-34  /* next 33 bits                    */
#define OM_PIO2_2T    0x1.3198a2e037073p-69  /* pi/2 - PIO2_1 - PIO2_2          */
#define OM_PIO2_HI    0x1.921fb54442d18p+0
#define OM_PIO2_LO    0x1.1a62633145c07p-54
#define OM_PIO4_HI    0x1.921fb54442d18p-1
#define OM_LN2_HI     0x1.62e42fefa39efp-1
#define OM_LN2_LO     0x1.abc9e3b39803fp-56
#define OM_INVLN2     0x1.71547652b82fep+0

/* fdlibm k_sin.c / k_cos.c */
#define OM_S1 -1.66666666666666324348e-01
#define OM_S2  8.33333333332248946124e-03
#define OM_S3 -1.98412698298579493134e-04
#define OM_S4  2.75573137070700676789e-06
#define OM_S5 -2.50507602534068634195e-08
#define OM_S6  1.58969099521155010221e-10
#define OM_C1  4.16666666666666019037e-02
#define OM_C2 -1.38888888888741095749e-03
#define OM_C3  2.48015872894767294178e-05
#define OM_C4 -2.75573143513906633035e-07
#define OM_C5  2.08757232129817482790e-09
#define OM_C6 -1.13596475577881948265e-11

#ifndef OSSG_ORACLE_LIBM
/* ------------------------------------------------------------------ kernels
 * |x| <= pi/4 (+ a few ulp), y = tail of x.  Horner steps use fma explicitly. */
static inline double om_ksin(double x, double y)
{
    double z = x * x;
    double w = z * z;
    double r = fma(z * w, fma(z, OM_S6, OM_S5), fma(z, fma(z, OM_S4, OM_S3), OM_S2));
    double v = z * x;
    double t = fma(-v, r, 0.5 * y);   /* 0.5*y - v*r            */
    t = fma(z, t, -y);                /* z*(...) - y            */
    t = fma(-v, OM_S1, t);            /* (...) - v*S1           */
    return x - t;
}

static inline double om_kcos(double x, double y)
{
    double z = x * x;
    double w = z * z;
    double r = fma(w * w, fma(z, fma(z, OM_C6, OM_C5), OM_C4),
                   z * fma(z, fma(z, OM_C3, OM_C2), OM_C1));
    double hz = 0.5 * z;
    double w1 = 1.0 - hz;
    return w1 + (((1.0 - w1) - hz) + fma(z, r, -(x * y)));
}

/* x (degrees, |x| <= 45) -> radians as hi + lo */
static inline void om_deg2rad_dd(double x, double *hi, double *lo)
{
    double h = x * OM_DEG2RAD_HI;
    double l = fma(x, OM_DEG2RAD_HI, -h);
    *hi = h;
    *lo = fma(x, OM_DEG2RAD_LO, l);
}

/* Julia Base sind/cosd semantics (base/special/trig.jl): reduction in degrees, exact at
 * multiples of 90, signed zeros: sind(-180) = -0, sind(0) = +0, cosd(+-90) = +0. */
static inline double om_sind(double x)
{
    if (isnan(x) || isinf(x)) return NAN;
    double rx = copysign(fmod(x, 360.0), x);
    double arx = fabs(rx);
    double hi, lo;
    if (rx == 0.0) return rx;
    if (arx < 45.0) {
        om_deg2rad_dd(rx, &hi, &lo);
        return om_ksin(hi, lo);
    } else if (arx <= 135.0) {
        om_deg2rad_dd(90.0 - arx, &hi, &lo);
        return copysign(om_kcos(hi, lo), rx);
    } else if (arx == 180.0) {
        return copysign(0.0, rx);
    } else if (arx < 225.0) {
        om_deg2rad_dd((180.0 - arx) * (rx < 0.0 ? -1.0 : 1.0), &hi, &lo); /* (180-arx)*sign(rx) */
        return om_ksin(hi, lo);
    } else if (arx <= 315.0) {
        om_deg2rad_dd(270.0 - arx, &hi, &lo);
        return -copysign(om_kcos(hi, lo), rx);
    } else {
        om_deg2rad_dd(rx - copysign(360.0, rx), &hi, &lo);
        return om_ksin(hi, lo);
    }
}

static inline double om_cosd(double x)
{
    if (isnan(x) || isinf(x)) return NAN;
    double rx = fabs(fmod(x, 360.0));
    double hi, lo;
    if (rx <= 45.0) {
        om_deg2rad_dd(rx, &hi, &lo);
        return om_kcos(hi, lo);
    } else if (rx < 135.0) {
        om_deg2rad_dd(90.0 - rx, &hi, &lo);
        return om_ksin(hi, lo);
    } else if (rx <= 225.0) {
        om_deg2rad_dd(180.0 - rx, &hi, &lo);
        return -om_kcos(hi, lo);
    } else if (rx < 315.0) {
        om_deg2rad_dd(rx - 270.0, &hi, &lo);
        return om_ksin(hi, lo);
    } else {
        om_deg2rad_dd(360.0 - rx, &hi, &lo);
        return om_kcos(hi, lo);
    }
}

/* sin/cos of radians: |x| <= pi/4 direct; otherwise two-stage Cody-Waite with the
 * 33+33+53-bit split of pi/2 (fdlibm e_rem_pio2.c "medium" path, second iteration
 * always taken).  Valid for |x| < 2^20*pi/2, far beyond this path's |x| <= 2*pi. */
static inline int om_rem_pio2(double x, double *y0, double *y1)
{
    double fn = rint(x * OM_INVPIO2);
    double t = fma(-fn, OM_PIO2_1, x);           /* exact: fn*PIO2_1 has <= 53 bits */
    double w = fn * OM_PIO2_2;
    double r = t - w;
    w = fma(fn, OM_PIO2_2T, -((t - r) - w));
    *y0 = r - w;
    *y1 = (r - *y0) - w;
    return (int)fn;
}

/* |x| < 2^-5: Taylor through x^7 (truncation z^4/9! < 2.6e-18 relative, below 0.03 ulp);
 * the haversine half-differences of every grid finer than ~3.5 degrees land here */
#define OM_SMALL_ARG 0x1p-5
static inline double om_sin_small(double x)
{
    const double T1 = -0x1.5555555555555p-3, T2 = 0x1.1111111111111p-7, T3 = -0x1.a01a01a01a01ap-13;
    double z = x * x;
    return fma(x * z, fma(z, fma(z, T3, T2), T1), x);
}

static inline double om_sin(double x)
{
    if (isnan(x) || isinf(x)) return NAN;
    if (fabs(x) < OM_SMALL_ARG) return om_sin_small(x);
    if (fabs(x) <= OM_PIO4_HI) return om_ksin(x, 0.0);
    double y0, y1;
    int n = om_rem_pio2(x, &y0, &y1);
    switch (n & 3) {
    case 0:  return  om_ksin(y0, y1);
    case 1:  return  om_kcos(y0, y1);
    case 2:  return -om_ksin(y0, y1);
    default: return -om_kcos(y0, y1);
    }
}

static inline double om_cos(double x)
{
    if (isnan(x) || isinf(x)) return NAN;
    if (fabs(x) <= OM_PIO4_HI) return om_kcos(x, 0.0);
    double y0, y1;
    int n = om_rem_pio2(x, &y0, &y1);
    switch (n & 3) {
    case 0:  return  om_kcos(y0, y1);
    case 1:  return -om_ksin(y0, y1);
    case 2:  return -om_kcos(y0, y1);
    default: return  om_ksin(y0, y1);
    }
}

/* fdlibm s_atan.c */
static inline double om_atan(double x)
{
    static const double atanhi[4] = { 0x1.dac670561bb4fp-2, 0x1.921fb54442d18p-1,
                                      0x1.f730bd281f69bp-1, 0x1.921fb54442d18p+0 };
    static const double atanlo[4] = { 0x1.a2b7f222f65e2p-56, 0x1.1a62633145c07p-55,
                                      0x1.007887af0cbbdp-56, 0x1.1a62633145c07p-54 };
    static const double aT[11] = {
        3.33333333333329318027e-01, -1.99999999998764832476e-01,
        1.42857142725034663711e-01, -1.11111104054623557880e-01,
        9.09088713343650656196e-02, -7.69187620504482999495e-02,
        6.66107313738753120669e-02, -5.83357013379057348645e-02,
        4.97687799461593236017e-02, -3.65315727442169155270e-02,
        1.62858201153657823623e-02 };
    if (isnan(x)) return x;
    double ax = fabs(x);
    int id;
    double t;
    if (ax >= 0x1p66) return copysign(atanhi[3] + atanlo[3], x);
    if (ax < 0.4375) {
        if (ax < 0x1p-27) return x;
        id = -1; t = x;
    } else if (ax < 0.6875) {
        id = 0; t = fma(2.0, ax, -1.0) / (2.0 + ax);
    } else if (ax < 1.1875) {
        id = 1; t = (ax - 1.0) / (ax + 1.0);
    } else if (ax < 2.4375) {
        id = 2; t = (ax - 1.5) / fma(1.5, ax, 1.0);
    } else {
        id = 3; t = -1.0 / ax;
    }
    double z = t * t;
    double w = z * z;
    double s1 = z * fma(w, fma(w, fma(w, fma(w, fma(w, aT[10], aT[8]), aT[6]), aT[4]), aT[2]), aT[0]);
    double s2 = w * fma(w, fma(w, fma(w, fma(w, aT[9], aT[7]), aT[5]), aT[3]), aT[1]);
    double s = s1 + s2;
    if (id < 0) return fma(-t, s, t);                 /* t - t*s */
    double r = atanhi[id] - (fma(t, s, -atanlo[id]) - t);
    return copysign(r, x);
}

/* fdlibm e_asin.c */
static inline double om_asin_R(double t)      /* R(t) = P(t)/Q(t) */
{
    const double pS0 = 1.66666666666666657415e-01, pS1 = -3.25565818622400915405e-01,
                 pS2 = 2.01212532134862925881e-01, pS3 = -4.00555345006794114027e-02,
                 pS4 = 7.91534994289814532176e-04, pS5 = 3.47933107596021167570e-05,
                 qS1 = -2.40339491173441421878e+00, qS2 = 2.02094576023350569471e+00,
                 qS3 = -6.88283971605453293030e-01, qS4 = 7.70381505559019352791e-02;
    double p = t * fma(t, fma(t, fma(t, fma(t, fma(t, pS5, pS4), pS3), pS2), pS1), pS0);
    double q = fma(t, fma(t, fma(t, fma(t, qS4, qS3), qS2), qS1), 1.0);
    return p / q;
}

static inline double om_asin(double x)
{
    if (isnan(x)) return x;
    double ax = fabs(x);
    if (ax >= 1.0) {
        if (ax == 1.0) return fma(x, OM_PIO2_HI, x * OM_PIO2_LO);
        return NAN;
    }
    if (ax < OM_SMALL_ARG) {
        /* Taylor through x^11 (truncation (10395/599040) t^6 < 1.6e-20 relative) */
        const double A1 = 0x1.5555555555555p-3, A2 = 0x1.3333333333333p-4, A3 = 0x1.6db6db6db6db7p-5,
                     A4 = 0x1.f1c71c71c71c7p-6, A5 = 0x1.6e8ba2e8ba2e9p-6;
        double t = x * x;
        double p = fma(t, fma(t, fma(t, fma(t, A5, A4), A3), A2), A1);
        return fma(x, t * p, x);
    }
    if (ax < 0.5) return fma(x, om_asin_R(x * x), x);
    double w = 1.0 - ax;
    double t = w * 0.5;
    double r = om_asin_R(t);
    double s = sqrt(t);
    double res;
    if (ax >= 0.975) {
        res = OM_PIO2_HI - (2.0 * fma(s, r, s) - OM_PIO2_LO);
    } else {
        union { double d; uint64_t u; } sw;
        sw.d = s;
        sw.u &= 0xffffffff00000000ull;
        double f = sw.d;
        double c = fma(-f, f, t) / (s + f);           /* (t - f*f)/(s+f), f*f exact */
        double p = fma(2.0 * s, r, -(OM_PIO2_LO - 2.0 * c));
        double q = OM_PIO4_HI - 2.0 * f;
        res = OM_PIO4_HI - (p - q);
    }
    return copysign(res, x);
}

/* ------------------------------------------------------------------ double-double
 * (cold functions only) */
typedef struct { double hi, lo; } om_dd;

static inline om_dd om_two_sum(double a, double b)
{
    om_dd r; r.hi = a + b; double bb = r.hi - a;
    r.lo = (a - (r.hi - bb)) + (b - bb); return r;
}
static inline om_dd om_quick_two_sum(double a, double b)
{
    om_dd r; r.hi = a + b; r.lo = b - (r.hi - a); return r;
}
static inline om_dd om_two_prod(double a, double b)
{
    om_dd r; r.hi = a * b; r.lo = fma(a, b, -r.hi); return r;
}
static inline om_dd om_dd_add(om_dd a, om_dd b)
{
    om_dd s = om_two_sum(a.hi, b.hi);
    om_dd t = om_two_sum(a.lo, b.lo);
    s.lo += t.hi;
    s = om_quick_two_sum(s.hi, s.lo);
    s.lo += t.lo;
    return om_quick_two_sum(s.hi, s.lo);
}
static inline om_dd om_dd_neg(om_dd a) { om_dd r = { -a.hi, -a.lo }; return r; }
static inline om_dd om_dd_sub(om_dd a, om_dd b) { return om_dd_add(a, om_dd_neg(b)); }
static inline om_dd om_dd_mul(om_dd a, om_dd b)
{
    om_dd p = om_two_prod(a.hi, b.hi);
    p.lo += fma(a.hi, b.lo, a.lo * b.hi);
    return om_quick_two_sum(p.hi, p.lo);
}
static inline om_dd om_dd_mul_d(om_dd a, double b)
{
    om_dd p = om_two_prod(a.hi, b);
    p.lo = fma(a.lo, b, p.lo);
    return om_quick_two_sum(p.hi, p.lo);
}
static inline om_dd om_dd_scale2(om_dd a, double pow2)   /* exact: pow2 is a power of two */
{
    om_dd r = { a.hi * pow2, a.lo * pow2 }; return r;
}
static inline om_dd om_dd_div(om_dd a, om_dd b)
{
    double q1 = a.hi / b.hi;
    om_dd r = om_dd_sub(a, om_dd_mul_d(b, q1));
    double q2 = r.hi / b.hi;
    r = om_dd_sub(r, om_dd_mul_d(b, q2));
    double q3 = r.hi / b.hi;
    om_dd q = om_quick_two_sum(q1, q2);
    om_dd d3 = { q3, 0.0 };
    return om_dd_add(q, d3);
}
static inline om_dd om_dd_sqrt(om_dd a)
{
    if (a.hi == 0.0) { om_dd z = { 0.0, 0.0 }; return z; }
    double x = 1.0 / sqrt(a.hi);
    double ax = a.hi * x;
    om_dd ax2 = om_two_prod(ax, ax);
    om_dd d = om_dd_sub(a, ax2);
    om_dd r = om_two_sum(ax, d.hi * (x * 0.5));
    /* one more Newton step in dd for full accuracy: r <- (r + a/r)/2 */
    om_dd q = om_dd_div(a, r);
    return om_dd_scale2(om_dd_add(r, q), 0.5);
}

/* exp(x)-1 and exp(x) for a dd argument, |x| < ~700 */
static inline om_dd om_dd_expm1_small(om_dd m)    /* |m| <= 4e-4: Taylor to m^11/11! */
{
    static const double inv_hi[10] = { 0x1.0000000000000p-1, 0x1.5555555555555p-3,
        0x1.5555555555555p-5, 0x1.1111111111111p-7, 0x1.6c16c16c16c17p-10,
        0x1.a01a01a01a01ap-13, 0x1.a01a01a01a01ap-16, 0x1.71de3a556c734p-19,
        0x1.27e4fb7789f5cp-22, 0x1.ae64567f544e4p-26 };
    static const double inv_lo[10] = { 0x0.0p+0, 0x1.5555555555555p-57,
        0x1.5555555555555p-59, 0x1.1111111111111p-63, -0x1.f49f49f49f49fp-65,
        0x1.a01a01a01a01ap-73, 0x1.a01a01a01a01ap-76, -0x1.c154f8ddc6c00p-73,
        0x1.cbbc05b4fa99ap-76, -0x1.c062e06d1f209p-80 };
    om_dd s = m;
    om_dd p = m;
    for (int k = 0; k < 10; ++k) {
        p = om_dd_mul(p, m);
        om_dd c = { inv_hi[k], inv_lo[k] };
        s = om_dd_add(s, om_dd_mul(p, c));
    }
    return s;
}

static inline om_dd om_dd_exp(om_dd x)
{
    om_dd one = { 1.0, 0.0 };
    if (x.hi == 0.0 && x.lo == 0.0) return one;
    double k = rint(x.hi * OM_INVLN2);
    om_dd ln2 = { OM_LN2_HI, OM_LN2_LO };
    om_dd r = om_dd_sub(x, om_dd_mul_d(ln2, k));
    om_dd m = om_dd_scale2(r, 0x1p-10);
    om_dd p = om_dd_expm1_small(m);
    for (int i = 0; i < 10; ++i)               /* expm1(2m) = 2p + p^2 */
        p = om_dd_add(om_dd_scale2(p, 2.0), om_dd_mul(p, p));
    om_dd e = om_dd_add(one, p);
    double sc = ldexp(1.0, (int)k);
    return om_dd_scale2(e, sc);
}

/* crude double log (own code, ~1e-15): seed for the Newton step in om_dd_log */
static inline double om_log_seed(double x)
{
    int e;
    double m = frexp(x, &e);                   /* m in [0.5,1) */
    if (m < 0x1.6a09e667f3bcdp-1) { m *= 2.0; e -= 1; }
    double s = (m - 1.0) / (m + 1.0);
    double z = s * s;
    double p = 1.0 / 21.0;
    for (int k = 19; k >= 1; k -= 2) p = fma(p, z, 1.0 / (double)k);
    return fma((double)e, OM_LN2_HI, 2.0 * s * p);
}

static inline om_dd om_dd_log(om_dd x)        /* x > 0 */
{
    om_dd one = { 1.0, 0.0 };
    /* one Newton step y <- y + x*exp(-y) - 1 from a ~1e-15 seed: error ~5e-31, far below the
     * half-ulp of the final double (a second step changed no bit in 10^6 samples) */
    om_dd y = { om_log_seed(x.hi), 0.0 };
    om_dd e = om_dd_exp(om_dd_neg(y));
    om_dd c = om_dd_sub(om_dd_mul(x, e), one);
    return om_dd_add(y, c);
}

static inline double om_asinh(double t)
{
    if (isnan(t) || isinf(t) || t == 0.0) return t;
    double a = fabs(t);
    om_dd one = { 1.0, 0.0 };
    om_dd a2 = om_two_prod(a, a);
    om_dd ad = { a, 0.0 };
    om_dd u = om_dd_add(ad, om_dd_sqrt(om_dd_add(a2, one)));
    om_dd y = om_dd_log(u);
    return copysign(y.hi, t);
}

static inline double om_sinh(double x)
{
    if (isnan(x) || isinf(x) || x == 0.0) return x;
    om_dd ax = { fabs(x), 0.0 };
    om_dd one = { 1.0, 0.0 };
    om_dd e = om_dd_exp(ax);
    om_dd s = om_dd_scale2(om_dd_sub(e, om_dd_div(one, e)), 0.5);
    return copysign(s.hi, x);
}

static inline double om_cosh(double x)
{
    if (isnan(x)) return x;
    if (isinf(x)) return INFINITY;
    om_dd ax = { fabs(x), 0.0 };
    om_dd one = { 1.0, 0.0 };
    om_dd e = om_dd_exp(ax);
    om_dd s = om_dd_scale2(om_dd_add(e, om_dd_div(one, e)), 0.5);
    return s.hi;
}

#else  /* OSSG_ORACLE_LIBM: glibc libm for everything (timing baseline / sensitivity) */

static inline void om_deg2rad_dd(double x, double *hi, double *lo)
{
    double h = x * OM_DEG2RAD_HI;
    double l = fma(x, OM_DEG2RAD_HI, -h);
    *hi = h;
    *lo = fma(x, OM_DEG2RAD_LO, l);
}
static inline double om_ksin(double x, double y) { return fma(y, cos(x), sin(x)); }
static inline double om_kcos(double x, double y) { return fma(-y, sin(x), cos(x)); }
static inline double om_sind(double x)
{
    double rx = copysign(fmod(x, 360.0), x);
    double arx = fabs(rx);
    double hi, lo;
    if (rx == 0.0) return rx;
    if (arx < 45.0) { om_deg2rad_dd(rx, &hi, &lo); return om_ksin(hi, lo); }
    else if (arx <= 135.0) { om_deg2rad_dd(90.0 - arx, &hi, &lo); return copysign(om_kcos(hi, lo), rx); }
    else if (arx == 180.0) return copysign(0.0, rx);
    else if (arx < 225.0) { om_deg2rad_dd((180.0 - arx) * (rx < 0.0 ? -1.0 : 1.0), &hi, &lo); return om_ksin(hi, lo); }
    else if (arx <= 315.0) { om_deg2rad_dd(270.0 - arx, &hi, &lo); return -copysign(om_kcos(hi, lo), rx); }
    else { om_deg2rad_dd(rx - copysign(360.0, rx), &hi, &lo); return om_ksin(hi, lo); }
}
static inline double om_cosd(double x)
{
    double rx = fabs(fmod(x, 360.0));
    double hi, lo;
    if (rx <= 45.0) { om_deg2rad_dd(rx, &hi, &lo); return om_kcos(hi, lo); }
    else if (rx < 135.0) { om_deg2rad_dd(90.0 - rx, &hi, &lo); return om_ksin(hi, lo); }
    else if (rx <= 225.0) { om_deg2rad_dd(180.0 - rx, &hi, &lo); return -om_kcos(hi, lo); }
    else if (rx < 315.0) { om_deg2rad_dd(rx - 270.0, &hi, &lo); return om_ksin(hi, lo); }
    else { om_deg2rad_dd(360.0 - rx, &hi, &lo); return om_kcos(hi, lo); }
}
static inline double om_sin(double x)   { return sin(x); }
static inline double om_cos(double x)   { return cos(x); }
static inline double om_atan(double x)  { return atan(x); }
static inline double om_asin(double x)  { return asin(x); }
static inline double om_asinh(double x) { return asinh(x); }
static inline double om_sinh(double x)  { return sinh(x); }
static inline double om_cosh(double x)  { return cosh(x); }
#endif

static inline double om_tand(double x) { return om_sind(x) / om_cosd(x); }

#endif /* OSSG_ORACLE_MATH_H */
