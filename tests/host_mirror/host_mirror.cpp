// tests/host_mirror/host_mirror.cpp -- TEST-ONLY host build of the product's per-point headers.
//
// Compiles csrc/ossg_math.cuh and csrc/ossg_tripolar_core.cuh with g++ (OSSG_HD expands to
// `inline`) and drives them with plain loops, so that the index maps and formulas that the CUDA
// kernels execute can be compared bit-for-bit with the oracle on a machine without a GPU.
// It is NOT part of the shipped library and is never used as a fallback.
#include <cstddef>
#include <vector>
#include "../../orthogonalsphericalshellgrids.jl_b200/csrc/ossg_tripolar_core.cuh"

using namespace ossg;

struct HMParams {
    int Nx, Ny, Hx, Hy;
    double south, north_poles_latitude, first_pole_longitude, radius;
    int i0, j0, nx, ny;
};

static void build_tables(const HMParams& p, std::vector<double>& buf)
{
    const int Nx = p.Nx, Ny = p.Ny;
    buf.assign(tables_doubles(Nx, Ny, p.Hy), 0.0);
    double* xs = buf.data();
    double* ys = xs + 4 * (size_t)Nx;
    double* ll = ys + 4 * (size_t)Ny;
    const double a = focal_distance(p.north_poles_latitude);
    for (int i = 1; i <= Nx; ++i) {
        const int ip = preshift(i, Nx);
        const double lf = lam_face(ip, Nx), lc = lam_center(ip, Nx);
        xs[0 * (size_t)Nx + i - 1] = a * m::sind(lf);
        xs[1 * (size_t)Nx + i - 1] = a * m::cosd(lf);
        xs[2 * (size_t)Nx + i - 1] = a * m::sind(lc);
        xs[3 * (size_t)Nx + i - 1] = a * m::cosd(lc);
    }
    for (int j = 1; j <= Ny; ++j) {
        const double pf = phi_face(j, Ny, p.south), pc = phi_center(j, Ny, p.south);
        double psf, psc, shf, chf, shc, chc;                           // as the table kernel does
        m::asinh_sinh_cosh_cr(m::tand((90.0 - pf) / 2) / a, psf, shf, chf);
        m::asinh_sinh_cosh_cr(m::tand((90.0 - pc) / 2) / a, psc, shc, chc);
        ys[0 * (size_t)Ny + j - 1] = chf;
        ys[1 * (size_t)Ny + j - 1] = shf;
        ys[2 * (size_t)Ny + j - 1] = chc;
        ys[3 * (size_t)Ny + j - 1] = shc;
    }
    for (int r = 0; r <= p.Hy; ++r) latlon_row(r + 1 - p.Hy, Nx, Ny, p.Hy, p.south, p.radius, ll + (size_t)r * LL_NCOL);
}

static Tables view(const std::vector<double>& buf, int Nx, int Ny)
{
    const double* base = buf.data();
    Tables t;
    t.asF = base; t.acF = base + Nx; t.asC = base + 2 * (size_t)Nx; t.acC = base + 3 * (size_t)Nx;
    const double* y = base + 4 * (size_t)Nx;
    t.chF = y; t.shF = y + Ny; t.chC = y + 2 * (size_t)Ny; t.shC = y + 3 * (size_t)Ny;
    t.ll = y + 4 * (size_t)Ny;
    return t;
}

struct WindowAcc {
    double* const* c; int I0, J0, pitch;
    void operator()(int loc, int I, int J, double& lam, double& phi) const
    {
        const size_t o = (size_t)(J - J0) * pitch + (I - I0);
        lam = c[loc][o]; phi = c[4 + loc][o];
    }
};
struct ClosedAcc {
    Tables tb; GridDims g; double fpl;
    void operator()(int loc, int I, int J, double& lam, double& phi) const { coord_at(tb, g, fpl, loc, I, J, lam, phi); }
};

extern "C" {

// coords[8]: window [i0-Hx, i0+nx-1+Hx] x [j0-Hy, j0+ny-1+Hy] of the global padded arrays
int hm_coords(const HMParams* p, double* const coords[8])
{
    std::vector<double> buf; build_tables(*p, buf);
    const Tables tb = view(buf, p->Nx, p->Ny);
    const GridDims g{ p->Nx, p->Ny, p->Hx, p->Hy };
    const int I0 = p->i0 - p->Hx, J0 = p->j0 - p->Hy, wnx = p->nx + 2 * p->Hx, wny = p->ny + 2 * p->Hy;
    for (int jw = 0; jw < wny; ++jw)
        for (int iw = 0; iw < wnx; ++iw)
            for (int loc = 0; loc < 4; ++loc) {
                double lam, phi;
                coord_at(tb, g, p->first_pole_longitude, loc, I0 + iw, J0 + jw, lam, phi);
                coords[loc][(size_t)jw * wnx + iw] = lam;
                coords[4 + loc][(size_t)jw * wnx + iw] = phi;
            }
    return 0;
}

// metrics[12] over the same window; closed_form != 0: coordinates re-evaluated from the tables
// (tile windows), else read from `coords` (whole-grid windows only)
int hm_metrics(const HMParams* p, double* const coords[8], double* const metrics[12], int closed_form)
{
    std::vector<double> buf; build_tables(*p, buf);
    const Tables tb = view(buf, p->Nx, p->Ny);
    const GridDims g{ p->Nx, p->Ny, p->Hx, p->Hy };
    const int I0 = p->i0 - p->Hx, J0 = p->j0 - p->Hy, wnx = p->nx + 2 * p->Hx, wny = p->ny + 2 * p->Hy;
    for (int jw = 0; jw < wny; ++jw)
        for (int iw = 0; iw < wnx; ++iw) {
            double out[N_METRIC];
            if (closed_form) {
                ClosedAcc acc{ tb, g, p->first_pole_longitude };
                metrics_at_padded(acc, tb, g, p->radius, I0 + iw, J0 + jw, out);
            } else {
                WindowAcc acc{ coords, I0, J0, wnx };
                metrics_at_padded(acc, tb, g, p->radius, I0 + iw, J0 + jw, out);
            }
            for (int k = 0; k < N_METRIC; ++k) metrics[k][(size_t)jw * wnx + iw] = out[k];
        }
    return 0;
}

int hm_math_eval(int fn, const double* x, double* y, size_t n)
{
    for (size_t e = 0; e < n; ++e) {
        switch (fn) {
        case 0: y[e] = m::sind(x[e]); break;
        case 1: y[e] = m::cosd(x[e]); break;
        case 2: y[e] = m::tand(x[e]); break;
        case 3: y[e] = m::sin_rad(x[e]); break;
        case 4: y[e] = m::cos_rad(x[e]); break;
        case 5: y[e] = m::atan_f(x[e]); break;
        case 6: y[e] = m::asin_f(x[e]); break;
        case 7: y[e] = m::asinh_cr(x[e]); break;
        case 8: y[e] = m::sinh_cr(x[e]); break;
        case 9: y[e] = m::cosh_cr(x[e]); break;
        case 10: { double sn, cs; m::sincosd(x[e], sn, cs); y[e] = sn; break; }
        case 11: { double sn, cs; m::sincosd(x[e], sn, cs); y[e] = cs; break; }
        case 12: y[e] = m::atan_short(x[e]); break;          // callers restrict |x| < ATAN_SHORT
        case 16: { double sn, cs; bool bad = false; m::sincosd_lat_spec(x[e], sn, cs, bad); if (bad) m::sincosd(x[e], sn, cs); y[e] = sn; break; }
        case 17: { double sn, cs; bool bad = false; m::sincosd_lat_spec(x[e], sn, cs, bad); if (bad) m::sincosd(x[e], sn, cs); y[e] = cs; break; }
        case 18: { double sn, cs; bool bad = false; m::sincosd_lon_spec(x[e], sn, cs, bad); if (bad) m::sincosd(x[e], sn, cs); y[e] = sn; break; }
        case 19: { double sn, cs; bool bad = false; m::sincosd_lon_spec(x[e], sn, cs, bad); if (bad) m::sincosd(x[e], sn, cs); y[e] = cs; break; }
        case 20: y[e] = m::atan_lut(x[e], m::kATAN_LUT); break;
        case 21: { bool bad = false; const double r = m::atan_sl(x[e], m::kATAN_LUT, bad); y[e] = bad ? m::atan_f(x[e]) : r; break; }
        case 22: y[e] = m::cos_rad_inl(x[e]); break;
        case 13: { double p, sh, ch; m::asinh_sinh_cosh_cr(x[e], p, sh, ch); y[e] = p; break; }
        case 14: { double p, sh, ch; m::asinh_sinh_cosh_cr(x[e], p, sh, ch); y[e] = sh; break; }
        case 15: { double p, sh, ch; m::asinh_sinh_cosh_cr(x[e], p, sh, ch); y[e] = ch; break; }
        default: return -1;
        }
    }
    return 0;
}

// tile-kernel ("hot") forms against the general ones: in = n x 7 (lam_x, hphi1, c1, lam_y, hphi2, c2, R)
int hm_haversine(int hot, const double* in, double* out, size_t n)
{
    for (size_t e = 0; e < n; ++e) {
        const double* v = in + 7 * e;
        out[e] = hot ? haversine_hot(v[0], v[1], v[2], v[3], v[4], v[5], v[6])
                     : haversine_pre(v[0], v[1], v[2], v[3], v[4], v[5], v[6]);
    }
    return 0;
}
// in = n x 8 (lat, lon of the four corners a, b, c, d)
int hm_quadrilateral(int hot, const double* in, double* out, size_t n)
{
    for (size_t e = 0; e < n; ++e) {
        const double* v = in + 8 * e;
        const Vec3 a = lat_lon_to_cartesian(v[0], v[1]), b = lat_lon_to_cartesian(v[2], v[3]);
        const Vec3 c = lat_lon_to_cartesian(v[4], v[5]), d = lat_lon_to_cartesian(v[6], v[7]);
        out[e] = hot ? spherical_area_quadrilateral_hot(a, b, c, d) : spherical_area_quadrilateral(a, b, c, d);
    }
    return 0;
}

// halo_source exposed for index-map unit tests
int hm_halo_source(int locx, int locy, int I, int J, int Nx, int Ny, int* i, int* j, int* wrapped)
{
    bool w; int ii = 0, jj = 0;
    const bool ok = halo_source(locx, locy, I, J, Nx, Ny, ii, jj, w);
    *i = ii; *j = jj; *wrapped = w;
    return ok;
}

}  // extern "C"
