/*
 * oracle/ossg_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the data-parallel hot path of CliMA/OrthogonalSphericalShellGrids.jl
 * v0.2.1 (reference mounted read-only at /root/reference; nothing here reads it at run
 * time): TripolarGrid coordinate + metric generation and the zipper (north-fold) halo
 * fill.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference leg may load this library; the product (libossg_b200.so) never does.
 *
 * PARITY STATUS.  Julia is not installed here or on the GPU box, so the reference
 * cannot be executed.  This restatement is pinned against every known-answer the
 * reference holds for the path:
 *   - README.md:54-59  (six printed numbers of TripolarGrid(size=(60,30,1)))
 *   - test/test_zipper_boundary_conditions.jl:35-45,54,65-72 (exact halo / row-Ny values)
 *   - test/runtests.jl:32-39 (coordinate range inequalities)
 * Arithmetic that lives in un-vendored dependencies (Distances.jl 0.10 haversine,
 * Oceananigans 0.95-0.99 lat_lon_to_cartesian / spherical_area_quadrilateral /
 * generate_coordinate / LatitudeLongitudeGrid metrics / halo-fill ordering, Julia Base
 * degree trig) is restated from the published algorithms; each site is marked
 * "(external)".  Metrics are pinned only through the README numbers: the reference has
 * no other known-answer for them.
 *
 * Layout: every 2-D array is "padded": (Nx+2Hx) x (Ny+2Hy) doubles, column-major with i
 * fastest, Julia index (i,j) in [1-Hx,Nx+Hx] x [1-Hy,Ny+Hy] at offset
 * (i+Hx-1) + (j+Hy-1)*(Nx+2Hx)  -- the parent array of the reference's OffsetArrays.
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off -mfma; pthreads for the optional row-parallel loops).
 */
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <math.h>
#include <pthread.h>
#include <unistd.h>

#include "ossg_oracle_math.h"
#include "ossg_oracle.h"

/* ------------------------------------------------------------------ helpers */
static inline size_t pidx(const ossg_oracle_params *p, int i, int j)
{
    return (size_t)(i + p->Hx - 1) + (size_t)(j + p->Hy - 1) * (size_t)(p->Nx + 2 * p->Hx);
}
static inline size_t psize(const ossg_oracle_params *p)
{
    return (size_t)(p->Nx + 2 * p->Hx) * (size_t)(p->Ny + 2 * p->Hy);
}


/* ------------------------------------------------------------------ row-parallel loops
 * (pthreads; used only when the caller asks for more than one thread -- the reference's
 * KernelAbstractions CPU backend splits its ndrange over Julia threads the same way) */
static int g_threads = 1;
/* Reading of `generate_coordinate(FT, ...)` (src/tripolar_grid.jl:89) for FT = Float32: 0 = the 1-D
 * longitudes stay Float64 (default); 1 = they are Float32 values (rounded once) widened back to
 * Float64 before the map.  See DESIGN.md section 4; neither can be pinned without Julia. */
static int g_lambda_f32 = 0;
typedef void (*row_fn)(void *ctx, int j);
typedef struct { row_fn fn; void *ctx; int j0, j1; } row_job;
static void *row_worker(void *arg)
{
    row_job *job = (row_job *)arg;
    for (int j = job->j0; j < job->j1; ++j) job->fn(job->ctx, j);
    return NULL;
}
static void parallel_rows(row_fn fn, void *ctx, int jbegin, int jend)   /* [jbegin, jend) */
{
    int nt = g_threads, n = jend - jbegin;
    if (nt > n) nt = n;
    if (nt <= 1) { for (int j = jbegin; j < jend; ++j) fn(ctx, j); return; }
    pthread_t *th = malloc(sizeof(pthread_t) * nt);
    row_job *jobs = malloc(sizeof(row_job) * nt);
    for (int t = 0; t < nt; ++t) {
        jobs[t].fn = fn; jobs[t].ctx = ctx;
        jobs[t].j0 = jbegin + (int)((long long)n * t / nt);
        jobs[t].j1 = jbegin + (int)((long long)n * (t + 1) / nt);
        pthread_create(&th[t], NULL, row_worker, &jobs[t]);
    }
    for (int t = 0; t < nt; ++t) pthread_join(th[t], NULL);
    free(th); free(jobs);
}

/* src/OrthogonalSphericalShellGrids.jl:24   convert_to_0_360(x) = ((x % 360) + 360) % 360 */
static inline double convert_to_0_360(double x)
{
    return fmod(fmod(x, 360.0) + 360.0, 360.0);
}

/* Correctly rounded (a*na + b*nb)/den for small integers na, nb, den: the value a
 * Julia TwicePrecision `range` element rounds to (external: Base.range /
 * Oceananigans.Grids.generate_coordinate; src/tripolar_grid.jl:90,95). */
static double rational_point(double a, double na, double b, double nb, double den)
{
#ifndef OSSG_ORACLE_LIBM
    om_dd pa = om_two_prod(a, na), pb = om_two_prod(b, nb);
    om_dd s = om_dd_add(pa, pb);
    om_dd d = { den, 0.0 };
    om_dd q = om_dd_div(s, d);
    return q.hi;
#else
    return (double)(((long double)a * na + (long double)b * nb) / den);
#endif
}

/* C0: 1-D coordinates.  src/tripolar_grid.jl:90,95-97.
 * lamF[i-1], lamC[i-1] for i=1..Nx (pre-circshift), phiF[j-1], phiC[j-1] for j=1..Ny. */
void ossg_oracle_coordinate_vectors(const ossg_oracle_params *p,
                                    double *lamF, double *lamC, double *phiF, double *phiC)
{
    const int Nx = p->Nx, Ny = p->Ny;
    for (int i = 1; i <= Nx; ++i) {
        /* longitude = (-180, 180), Periodic: faces at -180 + (i-1)*360/Nx, centres half a cell on */
        lamF[i - 1] = rational_point(-180.0, (double)(Nx - (i - 1)), 180.0, (double)(i - 1), (double)Nx);
        lamC[i - 1] = rational_point(-180.0, (double)(2 * Nx - (2 * i - 1)), 180.0, (double)(2 * i - 1), (double)(2 * Nx));
        if (g_lambda_f32) { lamF[i - 1] = (double)(float)lamF[i - 1]; lamC[i - 1] = (double)(float)lamC[i - 1]; }
    }
    const double south = p->southernmost_latitude;
    for (int j = 1; j <= Ny; ++j)   /* range(south, 90, length = Ny) */
        phiC[j - 1] = rational_point(south, (double)(Ny - j), 90.0, (double)(j - 1), (double)(Ny - 1));
    const double dphi = phiC[1] - phiC[0];
    for (int j = 1; j <= Ny; ++j)
        phiF[j - 1] = phiC[j - 1] - dphi / 2;
}

/* C1: one staggering of _compute_tripolar_coordinates!, src/generate_tripolar_coordinates.jl:60-88.
 * i is the PRE-circshift index (1-based). */
static inline void tripolar_point(double sinh_psi, double cosh_psi, double lam1, int i, int Nlam,
                                  double first_pole_longitude, double focal_distance,
                                  double *lam_out, double *phi_out)
{
    double x = focal_distance * om_sind(lam1) * cosh_psi;                 /* :67 */
    double y = focal_distance * om_cosd(lam1) * sinh_psi;                 /* :68 */
    int on_the_north_pole = (x == 0.0) & (y == 0.0);                      /* :74 */
    double north_pole_value = (i == 1) ? -90.0 : 90.0;                    /* :75 */
    const double m180_pi = -180.0 / 0x1.921fb54442d18p+1;                 /* -180 / pi   */
    const double p360_pi = 360.0 / 0x1.921fb54442d18p+1;                  /*  360 / pi   */
    double lam = on_the_north_pole ? north_pole_value : m180_pi * om_atan(y / x);   /* :77 */
    double phi = 90.0 - p360_pi * om_atan(sqrt(y * y + x * x));           /* :78 */
    lam += (i <= Nlam / 2) ? -90.0 : 90.0;                                /* :82 */
    lam += first_pole_longitude + 90.0;                                   /* :86 */
    lam = convert_to_0_360(lam);                                          /* :87 */
    *lam_out = lam;
    *phi_out = phi;
}


typedef struct {
    const ossg_oracle_params *p; double a;
    const double *lamF, *lamC, *phiF, *phiC, *shF, *chF, *shC, *chC;
    double *dense[8];
} c1_ctx;

static void c1_row(void *vctx, int j)
{
    c1_ctx *c = (c1_ctx *)vctx;
    const ossg_oracle_params *p = c->p;
    const int Nx = p->Nx;
    const double a = c->a;
    double *lCC = c->dense[0], *lFC = c->dense[1], *lCF = c->dense[2], *lFF = c->dense[3];
    double *pCC = c->dense[4], *pFC = c->dense[5], *pCF = c->dense[6], *pFF = c->dense[7];
    for (int i = 1; i <= Nx; ++i) {
        size_t d = (size_t)(i - 1) + (size_t)(j - 1) * Nx;
#ifndef OSSG_ORACLE_LITERAL
        double sF = c->shF[j - 1], cF = c->chF[j - 1], sC = c->shC[j - 1], cC = c->chC[j - 1];
#else
        double psiF = om_asinh(om_tand((90.0 - c->phiF[j - 1]) / 2) / a);
        double psiC = om_asinh(om_tand((90.0 - c->phiC[j - 1]) / 2) / a);
        double sF = om_sinh(psiF), cF = om_cosh(psiF), sC = om_sinh(psiC), cC = om_cosh(psiC);
#endif
        /* staggering table, generate_tripolar_coordinates.jl:60-63 */
        tripolar_point(sF, cF, c->lamF[i - 1], i, Nx, p->first_pole_longitude, a, &lFF[d], &pFF[d]);
        tripolar_point(sC, cC, c->lamF[i - 1], i, Nx, p->first_pole_longitude, a, &lFC[d], &pFC[d]);
        tripolar_point(sF, cF, c->lamC[i - 1], i, Nx, p->first_pole_longitude, a, &lCF[d], &pCF[d]);
        tripolar_point(sC, cC, c->lamC[i - 1], i, Nx, p->first_pole_longitude, a, &lCC[d], &pCC[d]);
    }
}

/* zipper + periodic-x fill of one padded 2-D array with sign +1
 * (src/tripolar_grid.jl:147-186; zipper: src/zipper_boundary_condition.jl:70-138;
 * fill order north/south first, periodic x last over the whole y extent: external,
 * Oceananigans.BoundaryConditions.fill_halo_regions!). */
static void fill_2d(const ossg_oracle_params *p, double *a, int locx, int locy)
{
    ossg_oracle_zipper_f64(a, p->Nx, p->Ny, 1, p->Hx, p->Hy, 0, locx, locy, 1.0, 0, 1);
    ossg_oracle_periodic_x_f64(a, p->Nx, p->Ny, 1, p->Hx, p->Hy, 0);
}

int ossg_oracle_check_params(const ossg_oracle_params *p)
{
    if (p->Nx < 2 || p->Ny < 2 || p->Hx < 1 || p->Hy < 1) return -1;
    if (p->Nx % 2 != 0) return -2;           /* src/tripolar_grid.jl:81-83 */
    if (p->Hy > p->Ny - 1 || p->Hx > p->Nx) return -3;
    return 0;
}

/* C0-C3: the 8 padded coordinate arrays, order: lam cc, fc, cf, ff, phi cc, fc, cf, ff. */
int ossg_oracle_coords(const ossg_oracle_params *p, double *const coords[8])
{
    int rc = ossg_oracle_check_params(p);
    if (rc) return rc;
    const int Nx = p->Nx, Ny = p->Ny;
    const double a = om_tand((90.0 - p->north_poles_latitude) / 2);       /* src/tripolar_grid.jl:76 */

    double *lamF = malloc(sizeof(double) * Nx), *lamC = malloc(sizeof(double) * Nx);
    double *phiF = malloc(sizeof(double) * Ny), *phiC = malloc(sizeof(double) * Ny);
    ossg_oracle_coordinate_vectors(p, lamF, lamC, phiF, phiC);

    /* dense Nx x Ny work arrays (src/tripolar_grid.jl:102-110) */
    double *dense[8];
    for (int k = 0; k < 8; ++k) dense[k] = calloc((size_t)Nx * Ny, sizeof(double));

#ifndef OSSG_ORACLE_LITERAL
    /* psi depends on j only: hoisted out of the i loop (pure functions => same bits) */
    double *shF = malloc(sizeof(double) * Ny), *chF = malloc(sizeof(double) * Ny);
    double *shC = malloc(sizeof(double) * Ny), *chC = malloc(sizeof(double) * Ny);
    for (int j = 0; j < Ny; ++j) {
        double psiF = om_asinh(om_tand((90.0 - phiF[j]) / 2) / a);        /* generate_tripolar_coordinates.jl:66 */
        double psiC = om_asinh(om_tand((90.0 - phiC[j]) / 2) / a);
        shF[j] = om_sinh(psiF); chF[j] = om_cosh(psiF);
        shC[j] = om_sinh(psiC); chC[j] = om_cosh(psiC);
    }
#endif

    c1_ctx ctx = { p, a, lamF, lamC, phiF, phiC, NULL, NULL, NULL, NULL, { 0 } };
#ifndef OSSG_ORACLE_LITERAL
    ctx.shF = shF; ctx.chF = chF; ctx.shC = shC; ctx.chC = chC;
#endif
    for (int k = 0; k < 8; ++k) ctx.dense[k] = dense[k];
    parallel_rows(c1_row, &ctx, 1, Ny + 1);

    /* C2 circshift by (Nx/4, 0) (src/tripolar_grid.jl:121-130) and C3 set! into zeroed
     * padded fields (src/tripolar_grid.jl:154-176) */
    const int shift = Nx / 4;
    for (int k = 0; k < 8; ++k) {
        memset(coords[k], 0, sizeof(double) * psize(p));
        for (int j = 1; j <= Ny; ++j)
            for (int i = 1; i <= Nx; ++i) {
                int is = i - shift;                    /* new[i] = old[mod1(i - shift, Nx)] */
                if (is < 1) is += Nx;
                coords[k][pidx(p, i, j)] = dense[k][(size_t)(is - 1) + (size_t)(j - 1) * Nx];
            }
    }
    /* fill_halo_regions!, src/tripolar_grid.jl:178-186; locations: cc, fc, cf, ff */
    static const int locx[4] = { 0, 1, 0, 1 }, locy[4] = { 0, 0, 1, 1 };
    for (int k = 0; k < 8; ++k) fill_2d(p, coords[k], locx[k % 4], locy[k % 4]);

    for (int k = 0; k < 8; ++k) free(dense[k]);
#ifndef OSSG_ORACLE_LITERAL
    free(shF); free(chF); free(shC); free(chC);
#endif
    free(lamF); free(lamC); free(phiF); free(phiC);
    return 0;
}

/* (external) Distances.jl 0.10 Haversine: x = (lam, phi) of the first point, y of the second. */
static inline double haversine(double lam_x, double phi_x, double lam_y, double phi_y, double radius)
{
    const double d2r = 0x1.921fb54442d18p+1 / 180.0;       /* deg2rad(z) = z * (pi / 180) */
    double dlam = (lam_y - lam_x) * d2r;
    double phi1 = phi_x * d2r;
    double phi2 = phi_y * d2r;
    double dphi = phi2 - phi1;
    double s1 = om_sin(dphi / 2);
    double s2 = om_sin(dlam / 2);
    double a = s1 * s1 + om_cos(phi1) * om_cos(phi2) * (s2 * s2);
    double r = sqrt(a);
    if (!(r < 1.0)) r = (r != r) ? r : 1.0;                /* min(sqrt(a), 1) */
    return 2 * (radius * om_asin(r));
}

/* (external) Oceananigans.Grids.lat_lon_to_cartesian(lat, lon, 1) */
static inline void lat_lon_to_cartesian(double lat, double lon, double v[3])
{
    double cl = om_cosd(lat);
    v[0] = om_cosd(lon) * cl;
    v[1] = om_sind(lon) * cl;
    v[2] = om_sind(lat);
}
static inline double dot3(const double a[3], const double b[3])
{
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}
/* (external) Oceananigans.Grids.spherical_area_triangle (Eriksson 1990) */
static inline double spherical_area_triangle(const double a1[3], const double a2[3], const double a3[3])
{
    double c[3] = { a2[1] * a3[2] - a2[2] * a3[1],
                    a2[2] * a3[0] - a2[0] * a3[2],
                    a2[0] * a3[1] - a2[1] * a3[0] };
    double t = fabs(dot3(a1, c));
    t /= 1 + dot3(a1, a2) + dot3(a2, a3) + dot3(a1, a3);
    return 2 * om_atan(t);
}
/* (external) Oceananigans.Grids.spherical_area_quadrilateral */
static inline double spherical_area_quadrilateral(const double a[3], const double b[3],
                                                  const double c[3], const double d[3])
{
    return 0.5 * (spherical_area_triangle(a, b, c) + spherical_area_triangle(a, b, d) +
                  spherical_area_triangle(a, c, d) + spherical_area_triangle(b, c, d));
}

/* (external) LatitudeLongitudeGrid(size, latitude=(south,90), longitude=(-180,180), halo,
 * radius) precomputed metrics at row j (src/tripolar_grid.jl:277-300).  Rows 2-Hy..
 * are computed by Oceananigans' precompute kernel; row 1-Hy of the x-spacings and areas
 * is never written there and stays 0. */
typedef struct { double dx_fc, dx_cf, dx_ff, dx_cc, dy, az_fc, az_cf, az_ff, az_cc; } ll_row;

static ll_row latlon_row(const ossg_oracle_params *p, int j)
{
    const double pi = 0x1.921fb54442d18p+1;
    const double d2r = pi / 180.0;
    const int Ny = p->Ny;
    const double south = p->southernmost_latitude, R = p->radius;
    const double dlam = 360.0 / (double)p->Nx;
    const double dphi = (90.0 - south) / (double)Ny;
    ll_row r;
    r.dy = R * (dphi * d2r);
    if (j == 1 - p->Hy) {
        r.dx_fc = r.dx_cf = r.dx_ff = r.dx_cc = 0.0;
        r.az_fc = r.az_cf = r.az_ff = r.az_cc = 0.0;
        return r;
    }
    /* faces  phiF(j) = south + (j-1) dphi ; centres phiC(j) = phiF(j) + dphi/2 */
    double phiF_j  = rational_point(south, (double)(Ny - (j - 1)), 90.0, (double)(j - 1), (double)Ny);
    double phiF_j1 = rational_point(south, (double)(Ny - j), 90.0, (double)j, (double)Ny);
    double phiC_j  = rational_point(south, (double)(2 * Ny - (2 * j - 1)), 90.0, (double)(2 * j - 1), (double)(2 * Ny));
    double phiC_jm = rational_point(south, (double)(2 * Ny - (2 * j - 3)), 90.0, (double)(2 * j - 3), (double)(2 * Ny));
    double rdl = R * (dlam * d2r);
    r.dx_fc = rdl * om_cos(pi * phiC_j / 180);
    r.dx_cc = r.dx_fc;
    r.dx_cf = rdl * om_cos(pi * phiF_j / 180);
    r.dx_ff = r.dx_cf;
    double r2dl = R * R * (dlam * d2r);
    r.az_fc = r2dl * (om_sin(pi * phiF_j1 / 180) - om_sin(pi * phiF_j / 180));
    r.az_cc = r.az_fc;
    r.az_cf = r2dl * (om_sin(pi * phiC_j / 180) - om_sin(pi * phiC_jm / 180));
    r.az_ff = r.az_cf;
    return r;
}

#define P(i, j) pidx(p, (i), (j))
typedef struct { const ossg_oracle_params *p; const double *const *coords; double *const *metrics; } m1_ctx;

/* M1: one row of _calculate_metrics!, src/tripolar_grid_utils.jl:4-45 */
static void m1_row(void *vctx, int j)
{
    m1_ctx *cx = (m1_ctx *)vctx;
    const ossg_oracle_params *p = cx->p;
    const int Nx = p->Nx;
    const double R = p->radius;
    const double *lCC = cx->coords[0], *lFC = cx->coords[1], *lCF = cx->coords[2], *lFF = cx->coords[3];
    const double *pCC = cx->coords[4], *pFC = cx->coords[5], *pCF = cx->coords[6], *pFF = cx->coords[7];
    double *dxCC = cx->metrics[0], *dxFC = cx->metrics[1], *dxCF = cx->metrics[2], *dxFF = cx->metrics[3];
    double *dyCC = cx->metrics[4], *dyCF = cx->metrics[5], *dyFC = cx->metrics[6], *dyFF = cx->metrics[7];
    double *azCC = cx->metrics[8], *azFC = cx->metrics[9], *azCF = cx->metrics[10], *azFF = cx->metrics[11];
    for (int i = 1; i <= Nx; ++i) {
        size_t o = P(i, j);
        dxCC[o] = haversine(lFC[P(i + 1, j)], pFC[P(i + 1, j)], lFC[P(i, j)], pFC[P(i, j)], R);         /* :13 */
        dxFC[o] = haversine(lCC[P(i, j)], pCC[P(i, j)], lCC[P(i - 1, j)], pCC[P(i - 1, j)], R);         /* :14 */
        dxCF[o] = haversine(lFF[P(i + 1, j)], pFF[P(i + 1, j)], lFF[P(i, j)], pFF[P(i, j)], R);         /* :15 */
        dxFF[o] = haversine(lCF[P(i, j)], pCF[P(i, j)], lCF[P(i - 1, j)], pCF[P(i - 1, j)], R);         /* :16 */

        dyCC[o] = haversine(lCF[P(i, j + 1)], pCF[P(i, j + 1)], lCF[P(i, j)], pCF[P(i, j)], R);         /* :18 */
        dyFC[o] = haversine(lFF[P(i, j + 1)], pFF[P(i, j + 1)], lFF[P(i, j)], pFF[P(i, j)], R);         /* :19 */
        dyCF[o] = haversine(lCC[P(i, j)], pCC[P(i, j)], lCC[P(i, j - 1)], pCC[P(i, j - 1)], R);         /* :20 */
        dyFF[o] = haversine(lFC[P(i, j)], pFC[P(i, j)], lFC[P(i, j - 1)], pFC[P(i, j - 1)], R);         /* :21 */

        double a[3], b[3], c[3], d[3];
        lat_lon_to_cartesian(pFF[P(i, j)], lFF[P(i, j)], a);                                             /* :23-26 */
        lat_lon_to_cartesian(pFF[P(i + 1, j)], lFF[P(i + 1, j)], b);
        lat_lon_to_cartesian(pFF[P(i + 1, j + 1)], lFF[P(i + 1, j + 1)], c);
        lat_lon_to_cartesian(pFF[P(i, j + 1)], lFF[P(i, j + 1)], d);
        azCC[o] = spherical_area_quadrilateral(a, b, c, d) * (R * R);                                    /* :28 */

        azFC[o] = dyFC[o] * dxFC[o];                                                                     /* :34 */
        azCF[o] = dyCF[o] * dxCF[o];                                                                     /* :35 */

        lat_lon_to_cartesian(pCC[P(i - 1, j - 1)], lCC[P(i - 1, j - 1)], a);                             /* :38-41 */
        lat_lon_to_cartesian(pCC[P(i, j - 1)], lCC[P(i, j - 1)], b);
        lat_lon_to_cartesian(pCC[P(i, j)], lCC[P(i, j)], c);
        lat_lon_to_cartesian(pCC[P(i - 1, j)], lCC[P(i - 1, j)], d);
        azFF[o] = spherical_area_quadrilateral(a, b, c, d) * (R * R);                                    /* :43 */
    }
}

/* M1-M3.  coords: lam cc,fc,cf,ff, phi cc,fc,cf,ff (padded, halo-filled).
 * metrics out: dx cc,fc,cf,ff ; dy cc,cf,fc,ff ; az cc,fc,cf,ff (padded). */
int ossg_oracle_metrics(const ossg_oracle_params *p, const double *const coords[8],
                        double *const metrics[12])
{
    int rc = ossg_oracle_check_params(p);
    if (rc) return rc;
    const int Nx = p->Nx, Ny = p->Ny;
    double *dxCC = metrics[0], *dxFC = metrics[1], *dxCF = metrics[2], *dxFF = metrics[3];
    double *dyCC = metrics[4], *dyCF = metrics[5], *dyFC = metrics[6], *dyFF = metrics[7];
    double *azCC = metrics[8], *azFC = metrics[9], *azCF = metrics[10], *azFF = metrics[11];
    for (int k = 0; k < 12; ++k) memset(metrics[k], 0, sizeof(double) * psize(p));

    m1_ctx cx = { p, coords, metrics };
    parallel_rows(m1_row, &cx, 1, Ny + 1);        /* src/tripolar_grid.jl:220-227 */

    /* M2: halo fill of the 12 metrics with sign +1 (src/tripolar_grid.jl:230-273) */
    fill_2d(p, dxCC, 0, 0); fill_2d(p, dxFC, 1, 0); fill_2d(p, dxCF, 0, 1); fill_2d(p, dxFF, 1, 1);
    fill_2d(p, dyCC, 0, 0); fill_2d(p, dyFC, 1, 0); fill_2d(p, dyCF, 0, 1); fill_2d(p, dyFF, 1, 1);
    fill_2d(p, azCC, 0, 0); fill_2d(p, azFC, 1, 0); fill_2d(p, azCF, 0, 1); fill_2d(p, azFF, 1, 1);

    /* M3: continue_south!, rows j in [1-Hy, 1], all i (src/tripolar_grid.jl:287-300, 336-357) */
    for (int j = 1 - p->Hy; j <= 1; ++j) {
        ll_row r = latlon_row(p, j);
        for (int i = 1 - p->Hx; i <= Nx + p->Hx; ++i) {
            size_t o = P(i, j);
            dxFF[o] = r.dx_ff; dxFC[o] = r.dx_fc; dxCF[o] = r.dx_cf; dxCC[o] = r.dx_cc;   /* :287-290 */
            dyFF[o] = r.dy;    dyFC[o] = r.dy;    dyCF[o] = r.dy;    dyCC[o] = r.dy;      /* :292-295 */
            azFF[o] = r.az_ff; azFC[o] = r.az_fc; azCF[o] = r.az_cf; azCC[o] = r.az_cc;   /* :297-300 */
        }
    }
#undef P
    return 0;
}

/* Whole constructor: the 20 padded arrays in OrthogonalSphericalShellGrid field order
 * (src/tripolar_grid.jl:308-328): lam cc,fc,cf,ff ; phi cc,fc,cf,ff ; dx cc,fc,cf,ff ;
 * dy cc,cf,fc,ff ; az cc,fc,cf,ff. */
int ossg_oracle_build_f64(const ossg_oracle_params *p, double *const out[20])
{
    int rc = ossg_oracle_coords(p, out);
    if (rc) return rc;
    return ossg_oracle_metrics(p, (const double *const *)out, out + 8);
}

/* F: FT = Float32: compute in Float64, round once at the end (src/tripolar_grid.jl:308-328 map(FT, .)) */
int ossg_oracle_build_f32(const ossg_oracle_params *p, float *const out[20])
{
    size_t n = psize(p);
    double *tmp[20];
    for (int k = 0; k < 20; ++k) tmp[k] = malloc(sizeof(double) * n);
    int rc = ossg_oracle_build_f64(p, tmp);
    if (!rc)
        for (int k = 0; k < 20; ++k)
            for (size_t e = 0; e < n; ++e) out[k][e] = (float)tmp[k][e];
    for (int k = 0; k < 20; ++k) free(tmp[k]);
    return rc;
}

/* ------------------------------------------------------------------ Z: zipper fold
 * src/zipper_boundary_condition.jl:70-155.  c is the parent of a 3-D field,
 * (Nx+2Hx) x (Ny+2Hy) x (Nz+2Hz); the fold is applied for k-levels [k0, k0+nk) of the
 * parent (0-based), i in 1..Nx.  locx/locy: 0 = Center, 1 = Face. */
#define ZIPPER_BODY(T)                                                                         \
    const size_t sx = (size_t)(Nx + 2 * Hx), sxy = sx * (size_t)(Ny + 2 * Hy);                 \
    (void)Nz; (void)Hz;                                                                        \
    for (int k = k0; k < k0 + nk; ++k) {                                                       \
        T *ck = c + (size_t)k * sxy;                                                           \
        for (int i = 1; i <= Nx; ++i) {                                                        \
            double s = sign;                                                                   \
            int ip = locx ? Nx - i + 2 : Nx - i + 1;                                           \
            if (locx && ip > Nx) { s = fabs(s); ip -= Nx; }                                    \
            for (int j = 1; j <= Hy; ++j) {                                                    \
                int js = locy ? Ny - j + 1 : Ny - j;                                           \
                ck[(i + Hx - 1) + (size_t)(Ny + j + Hy - 1) * sx] =                            \
                    (T)(s * ck[(ip + Hx - 1) + (size_t)(js + Hy - 1) * sx]);                   \
            }                                                                                  \
            if (!locy && i > Nx / 2)                                                           \
                ck[(i + Hx - 1) + (size_t)(Ny + Hy - 1) * sx] =                                \
                    (T)(s * ck[(ip + Hx - 1) + (size_t)(Ny + Hy - 1) * sx]);                   \
        }                                                                                      \
    }

void ossg_oracle_zipper_f64(double *c, int Nx, int Ny, int Nz, int Hx, int Hy, int Hz,
                            int locx, int locy, double sign, int k0, int nk)
{
    ZIPPER_BODY(double)
}
void ossg_oracle_zipper_f32(float *c, int Nx, int Ny, int Nz, int Hx, int Hy, int Hz,
                            int locx, int locy, double sign, int k0, int nk)
{
    ZIPPER_BODY(float)
}

/* periodic west/east fill over the whole (y, z) parent extent (external: Oceananigans
 * periodic halo kernel; fills the NW/NE corners from the already-folded north halo). */
#define PERIODIC_BODY(T)                                                                       \
    const size_t sx = (size_t)(Nx + 2 * Hx);                                                   \
    const size_t rows = (size_t)(Ny + 2 * Hy) * (size_t)(Nz + 2 * Hz);                         \
    for (size_t r = 0; r < rows; ++r) {                                                        \
        T *row = c + r * sx;                                                                   \
        for (int h = 0; h < Hx; ++h) {                                                         \
            row[h] = row[Nx + h];                     /* west halo  <- east interior */        \
            row[Hx + Nx + h] = row[Hx + h];           /* east halo  <- west interior */        \
        }                                                                                      \
    }

void ossg_oracle_periodic_x_f64(double *c, int Nx, int Ny, int Nz, int Hx, int Hy, int Hz)
{
    PERIODIC_BODY(double)
}
void ossg_oracle_periodic_x_f32(float *c, int Nx, int Ny, int Nz, int Hx, int Hy, int Hz)
{
    PERIODIC_BODY(float)
}

/* ------------------------------------------------------------------ test hooks */
int ossg_oracle_uses_libm(void)
{
#ifdef OSSG_ORACLE_LIBM
    return 1;
#else
    return 0;
#endif
}

int ossg_oracle_max_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

void ossg_oracle_set_lambda_f32(int on) { g_lambda_f32 = on ? 1 : 0; }

void ossg_oracle_set_threads(int n)
{
    g_threads = n > 0 ? n : 1;
}

/* (external) Oceananigans.Grids.lat_lon_to_cartesian(lat, lon, radius) over n nodes
 * (test/test_tripolar_grid.jl:68, examples/visualize_tripolar_grid.jl:39-42) */
void ossg_oracle_cartesian_nodes(const double *lam, const double *phi, size_t n, double radius,
                                 double *x, double *y, double *z)
{
    for (size_t e = 0; e < n; ++e) {
        x[e] = radius * om_cosd(lam[e]) * om_cosd(phi[e]);
        y[e] = radius * om_sind(lam[e]) * om_cosd(phi[e]);
        z[e] = radius * om_sind(phi[e]);
    }
}

/* evaluate one elementary function over an array (tests compare with mpmath / glibc) */
int ossg_oracle_math_eval(int fn, const double *x, double *y, size_t n)
{
    for (size_t e = 0; e < n; ++e) {
        switch (fn) {
        case 0: y[e] = om_sind(x[e]); break;
        case 1: y[e] = om_cosd(x[e]); break;
        case 2: y[e] = om_tand(x[e]); break;
        case 3: y[e] = om_sin(x[e]); break;
        case 4: y[e] = om_cos(x[e]); break;
        case 5: y[e] = om_atan(x[e]); break;
        case 6: y[e] = om_asin(x[e]); break;
        case 7: y[e] = om_asinh(x[e]); break;
        case 8: y[e] = om_sinh(x[e]); break;
        case 9: y[e] = om_cosh(x[e]); break;
        default: return -1;
        }
    }
    return 0;
}
