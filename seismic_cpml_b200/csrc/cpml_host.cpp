// Host-side helpers of libcpml_b200: the set-up and output phases that stay in the
// driver (profiles, source law, receiver search, Courant number, file writers).
// Pure host code; mirrors the behaviour of the reference's inline set-up code so that
// the C++ / Python / Fortran drivers in this repo all hand the same numbers to the GPU.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/cpml_b200.h"

namespace {

constexpr double kPi = 3.141592653589793238462643;   // 3D-iso :199

// Damping value set of one point of one side of the PML (3D-iso :466-484).
struct Damp { double d = 0.0, K = 1.0, alpha = 0.0; };

inline void damp_at(double depth_in_pml, double thickness, double d0, double npower,
                    double k_max, double alpha_max, Damp &out)
{
    if (depth_in_pml < 0.0) return;                  // outside this side's layer
    const double s = depth_in_pml / thickness;       // abscissa_normalized
    const double sp = std::pow(s, npower);
    out.d = d0 * sp;
    out.K = 1.0 + (k_max - 1.0) * sp;
    out.alpha = alpha_max * (1.0 - s);
}

inline void recursion_coefs(const Damp &p, double deltat, double &a, double &b)
{
    b = std::exp(-(p.d / p.K + p.alpha) * deltat);                       // :517
    a = 0.0;
    if (std::fabs(p.d) > 1.e-6) a = p.d * (b - 1.0) / (p.K * (p.d + p.K * p.alpha));   // :521
}

std::string join(const char *dir, const char *name)
{
    std::string s = (dir && *dir) ? dir : ".";
    if (s.back() != '/') s += '/';
    return s + name;
}

}  // namespace

extern "C" int32_t cpml_host_pml_profile(int32_t n, double delta, double deltat, int32_t npoints_pml,
                                         int32_t use_pml_min, int32_t use_pml_max, double cp, double rcoef,
                                         double npower, double k_max_pml, double alpha_max_pml,
                                         int32_t origin_top_uses_n, int32_t clamp_alpha,
                                         double *a, double *b, double *K,
                                         double *a_half, double *b_half, double *K_half)
{
    return cpml_host_pml_profile_visco(n, delta, deltat, npoints_pml, use_pml_min, use_pml_max, cp, 1.0, rcoef, npower,
                                       k_max_pml, alpha_max_pml, origin_top_uses_n, clamp_alpha, a, b, K, a_half, b_half, K_half);
}

// 3D-visco :533-801: the same profiles with d0 scaled by dsqrt(taumax) (:547); sqrt_taumax = 1 is
// the isotropic formula bit for bit.
extern "C" int32_t cpml_host_pml_profile_visco(int32_t n, double delta, double deltat, int32_t npoints_pml,
                                               int32_t use_pml_min, int32_t use_pml_max, double cp, double sqrt_taumax,
                                               double rcoef, double npower, double k_max_pml, double alpha_max_pml,
                                               int32_t origin_top_uses_n, int32_t clamp_alpha,
                                               double *a, double *b, double *K,
                                               double *a_half, double *b_half, double *K_half)
{
    if (n < 1 || !(delta > 0) || !(deltat > 0) || npoints_pml < 1 || !a || !b || !K || !a_half || !b_half || !K_half)
        return CPML_EINVAL;
    if (npower < 1) return CPML_EINVAL;                                   // :410
    const double thickness = npoints_pml * delta;                        // :402
    const double d0 = -(npower + 1) * cp * sqrt_taumax * std::log(rcoef) / (2.0 * thickness);   // :413 ; 3D-visco :547
    const double origin_min = thickness;                                  // :455
    const double origin_max = (origin_top_uses_n ? n : n - 1) * delta - thickness;   // :456 / 2D-4th :401

    for (int32_t idx = 0; idx < n; idx++) {
        const double pos = delta * static_cast<double>(idx);              // :461
        Damp full, half;
        if (use_pml_min) {
            damp_at(origin_min - pos, thickness, d0, npower, k_max_pml, alpha_max_pml, full);
            damp_at(origin_min - (pos + delta / 2.0), thickness, d0, npower, k_max_pml, alpha_max_pml, half);
        }
        if (use_pml_max) {   // a max-side hit overrides the min side, like the sequential ifs of :489-511
            damp_at(pos - origin_max, thickness, d0, npower, k_max_pml, alpha_max_pml, full);
            damp_at(pos + delta / 2.0 - origin_max, thickness, d0, npower, k_max_pml, alpha_max_pml, half);
        }
        if (clamp_alpha) {                                                // :514-515
            full.alpha = std::max(full.alpha, 0.0);
            half.alpha = std::max(half.alpha, 0.0);
        }
        recursion_coefs(full, deltat, a[idx], b[idx]);
        recursion_coefs(half, deltat, a_half[idx], b_half[idx]);
        K[idx] = full.K;
        K_half[idx] = half.K;
    }
    return CPML_OK;
}

extern "C" int32_t cpml_host_source_series(int32_t nstep, double deltat, double f0, double t0, double factor,
                                           double angle_force_deg, double *force_x, double *force_y)
{
    if (nstep < 1 || !force_x || !force_y) return CPML_EINVAL;
    const double a = kPi * kPi * f0 * f0;                                 // :1058
    const double rad = angle_force_deg * (kPi / 180.0);                   // :202
    const double sx = std::sin(rad), cy = std::cos(rad);
    for (int32_t it = 1; it <= nstep; it++) {
        const double t = static_cast<double>(it - 1) * deltat;           // :1059
        const double tau = t - t0;
        const double source_term = -factor * 2.0 * a * tau * std::exp(-a * (tau * tau));   // :1065
        force_x[it - 1] = sx * source_term;                               // :1070-1071
        force_y[it - 1] = cy * source_term;
    }
    return CPML_OK;
}

extern "C" int32_t cpml_host_find_receivers(int32_t nx, int32_t ny, double deltax, double deltay, int32_t nrec,
                                            double xdeb, double ydeb, double xfin, double yfin,
                                            int32_t *ix_rec, int32_t *iy_rec, double *dist)
{
    if (nx < 1 || ny < 1 || nrec < 1 || !ix_rec || !iy_rec) return CPML_EINVAL;
    const double xstep = nrec > 1 ? (xfin - xdeb) / static_cast<double>(nrec - 1) : 0.0;   // :683
    const double ystep = nrec > 1 ? (yfin - ydeb) / static_cast<double>(nrec - 1) : 0.0;
    for (int32_t r = 0; r < nrec; r++) {
        const double xr = xdeb + static_cast<double>(r) * xstep;          // :686-687
        const double yr = ydeb + static_cast<double>(r) * ystep;
        double best = 1.e+30;                                             // HUGEVAL :208
        // strict '<' with j outer, i inner: the first minimum wins (quirk B9)
        for (int32_t j = 1; j <= ny; j++)
            for (int32_t i = 1; i <= nx; i++) {
                const double ex = deltax * static_cast<double>(i - 1) - xr;
                const double ey = deltay * static_cast<double>(j - 1) - yr;
                const double dv = std::sqrt(ex * ex + ey * ey);
                if (dv < best) { best = dv; ix_rec[r] = i; iy_rec[r] = j; }
            }
        if (dist) dist[r] = best;
    }
    return CPML_OK;
}

// 3D-visco :832-853: explicit receiver targets; grid abscissa DELTAX*i when index_origin = 1
// (the viscoelastic program), DELTAX*(i-1) when 0.
extern "C" int32_t cpml_host_find_receivers_at(int32_t nx, int32_t ny, double deltax, double deltay, int32_t nrec,
                                               const double *xrec, const double *yrec, int32_t index_origin,
                                               int32_t *ix_rec, int32_t *iy_rec, double *dist)
{
    if (nx < 1 || ny < 1 || nrec < 1 || !xrec || !yrec || !ix_rec || !iy_rec) return CPML_EINVAL;
    if (index_origin != 0 && index_origin != 1) return CPML_EINVAL;
    for (int32_t r = 0; r < nrec; r++) {
        double best = 1.e+30;
        for (int32_t j = 1; j <= ny; j++)
            for (int32_t i = 1; i <= nx; i++) {
                const double ex = deltax * static_cast<double>(i - 1 + index_origin) - xrec[r];
                const double ey = deltay * static_cast<double>(j - 1 + index_origin) - yrec[r];
                const double dv = std::sqrt(ex * ex + ey * ey);
                if (dv < best) { best = dv; ix_rec[r] = i; iy_rec[r] = j; }
            }
        if (dist) dist[r] = best;
    }
    return CPML_OK;
}

extern "C" double cpml_host_courant(double cp, double deltat, double deltax, double deltay, double deltaz)
{
    double s = 1.0 / (deltax * deltax) + 1.0 / (deltay * deltay);        // :712 / 2D-2nd :513
    if (deltaz > 0.0) s += 1.0 / (deltaz * deltaz);
    return cp * deltat * std::sqrt(s);
}

// ---- gfortran list-directed output -------------------------------------------------
// The reference writes its seismogram, energy and timestamp files with `write(unit,*)` (3D-iso :1219-1229,
// :1254-1256, :1349): the text is whatever the compiler's list-directed formatting produces.  The reference Makefile
// builds with gfortran, whose run-time library (libgfortran/io/write.c, write_real / set_fnode_default) formats a
// default REAL(4) as 1PG16.9E2 and a REAL(8) as 1PG25.17E3 "with the same number of significant digits whether F or
// E editing is used" (9 and 17), a default INTEGER right-justified in 11 columns, starts every record with one blank
// and puts one blank before every item that follows another one (except a character item after a character item).
// Restated here so that the files are byte-for-byte what the gfortran build of the reference would write for the
// same numbers (tests/test_output_format.py checks hand-derived strings).
namespace {

// 1PGw.dEe of a list-directed item: d significant digits; F editing, followed by e+2 blanks, when the value rounded to
// d digits lies in [0.1, 10^d), E editing otherwise.
std::string gf_real(double v, int w, int d, int e)
{
    const int n = e + 2;
    char buf[96];
    std::string body;
    bool f_edit = true;
    if (std::isnan(v)) { body = "NaN"; f_edit = false; }
    else if (std::isinf(v)) { body = v < 0 ? "-Infinity" : "Infinity"; f_edit = false; }
    else if (v == 0.0) {
        snprintf(buf, sizeof buf, "%.*f", d - 1, v);
        body = buf;
    } else {
        snprintf(buf, sizeof buf, "%.*E", d - 1, v);              // correctly rounded to d significant digits
        const char *E = strchr(buf, 'E');
        const int ex = atoi(E + 1);
        if (ex >= -1 && ex < d) {
            snprintf(buf, sizeof buf, "%#.*f", d - (ex + 1), v);   // '#': gfortran keeps the point of "123456792."
            body = buf;
        } else {
            f_edit = false;
            body.assign(buf, (size_t)(E - buf));
            char ebuf[16];
            snprintf(ebuf, sizeof ebuf, "E%c%0*d", ex < 0 ? '-' : '+', e, std::abs(ex));
            body += ebuf;
        }
    }
    const int field = f_edit ? w - n : w;
    std::string out;
    if ((int)body.size() < field) out.assign((size_t)(field - (int)body.size()), ' ');
    out += body;
    if (f_edit) out.append((size_t)n, ' ');
    return out;
}

std::string gf_r4(double v) { return gf_real((double)(float)v, 16, 9, 2); }     // sngl(...)
std::string gf_r8(double v) { return gf_real(v, 25, 17, 3); }
std::string gf_i4(int v) { char b[16]; snprintf(b, sizeof b, "%11d", v); return b; }

}  // namespace

// The text gfortran writes for a list-directed REAL(4) (kind = 4, the value is first demoted like sngl()) or REAL(8)
// (kind = 8) item, without the record's leading blank / the item separator.  Exported for the tests and the drivers.
extern "C" int32_t cpml_host_format_real(double value, int32_t kind, char *out, int32_t capacity)
{
    if (!out || (kind != 4 && kind != 8)) return CPML_EINVAL;
    const std::string s = kind == 4 ? gf_r4(value) : gf_r8(value);
    if ((int)s.size() + 1 > capacity) return CPML_EINVAL;
    memcpy(out, s.c_str(), s.size() + 1);
    return CPML_OK;
}

// ---- writers ---------------------------------------------------------------------

extern "C" int32_t cpml_host_write_seismograms(const char *dir, const double *sisvx, const double *sisvy,
                                               int32_t nt, int32_t nrec, double deltat)
{
    if (!sisvx || !sisvy || nt < 1 || nrec < 0) return CPML_EINVAL;
    for (int comp = 0; comp < 2; comp++) {
        const double *sis = comp == 0 ? sisvx : sisvy;
        for (int32_t r = 1; r <= nrec; r++) {
            char name[64];
            snprintf(name, sizeof name, comp == 0 ? "Vx_file_%03d.dat" : "Vy_file_%03d.dat", r);   // :1346,1356
            FILE *f = fopen(join(dir, name).c_str(), "w");
            if (!f) return CPML_EINVAL;
            // time and amplitude, both demoted to single precision like sngl(...) :1349
            for (int32_t it = 1; it <= nt; it++)
                fprintf(f, " %s   %s\n", gf_r4((double)(it - 1) * deltat).c_str(),
                        gf_r4(sis[(size_t)(r - 1) * nt + (it - 1)]).c_str());
            fclose(f);
        }
    }
    return CPML_OK;
}

// timestampNNNNNN of the 3-D programs: progress of the simulation (3D-iso :1219-1229, 3D-visco :1469-1479)
extern "C" int32_t cpml_host_write_timestamp(const char *dir, int32_t it, double deltat, double vsolidnorm,
                                             double total_energy, double tcpu)
{
    if (it < 1) return CPML_EINVAL;
    char name[32];
    snprintf(name, sizeof name, "timestamp%06d", it);
    FILE *f = fopen(join(dir, name).c_str(), "w");
    if (!f) return CPML_EINVAL;
    const int int_tcpu = (int)tcpu, ihours = int_tcpu / 3600, iminutes = (int_tcpu - 3600 * ihours) / 60;
    const int iseconds = int_tcpu - 3600 * ihours - 60 * iminutes;
    fprintf(f, " Time step #  %s\n", gf_i4(it).c_str());                                   // write(IOUT,*) 'Time step # ',it
    fprintf(f, " Time:  %s  seconds\n", gf_r4((it - 1) * deltat).c_str());                 // 'Time: ',sngl(...),' seconds'
    fprintf(f, " Max norm velocity vector V (m/s) =  %s\n", gf_r8(vsolidnorm).c_str());
    fprintf(f, " Total energy =  %s\n", gf_r8(total_energy).c_str());
    fprintf(f, " Elapsed time in seconds =  %s\n", gf_r8(tcpu).c_str());
    fprintf(f, " Elapsed time in hh:mm:ss = %4d h %02d m %02d s\n", ihours, iminutes, iseconds);
    fprintf(f, " Mean elapsed time per time step in seconds =  %s\n", gf_r8(tcpu / (double)it).c_str());
    fclose(f);
    return CPML_OK;
}

// Vz_file_NNN.dat: not written by the reference (quirk B7: its plotgnu reads them, its programs record Vx and
// Vy only); same format as the other two, time axis minus t0 (0 for the isotropic program)
extern "C" int32_t cpml_host_write_seismograms_vz(const char *dir, const double *sisvz, int32_t nt, int32_t nrec,
                                                  double deltat, double t0)
{
    if (!sisvz || nt < 1 || nrec < 0) return CPML_EINVAL;
    for (int32_t r = 1; r <= nrec; r++) {
        char name[64];
        snprintf(name, sizeof name, "Vz_file_%03d.dat", r);
        FILE *f = fopen(join(dir, name).c_str(), "w");
        if (!f) return CPML_EINVAL;
        for (int32_t it = 1; it <= nt; it++)
            fprintf(f, " %s   %s\n", gf_r4((double)(it - 1) * deltat - t0).c_str(),
                    gf_r4(sisvz[(size_t)(r - 1) * nt + (it - 1)]).c_str());
        fclose(f);
    }
    return CPML_OK;
}

// write_seismograms of the viscoelastic programs: the time axis is shifted by -t0 (3D-visco :1603,1613;
// 2D-visco-4th :1178,1188); the 2-D programs also record the pressure, which their scheme holds half a
// time step later (:1164-1168), and name the Vy file after its staggered position (:1184)
extern "C" int32_t cpml_host_write_seismograms_visco(const char *dir, const double *sisvx, const double *sisvy,
                                                     const double *sispressure, int32_t nt, int32_t nrec,
                                                     double deltat, double t0)
{
    if (!sisvx || !sisvy || nt < 1 || nrec < 0) return CPML_EINVAL;
    const bool two_d = sispressure != nullptr;
    for (int comp = 0; comp < (two_d ? 3 : 2); comp++) {
        const double *sis = comp == 0 ? sisvx : comp == 1 ? sisvy : sispressure;
        const char *fmt = comp == 0 ? "Vx_file_%03d.dat"
                        : comp == 1 ? (two_d ? "Vy_file_half_a_grid_cell_away_from_Vx_%03d.dat" : "Vy_file_%03d.dat")
                                    : "pressure_file_%03d.dat";
        const double shift = comp == 2 ? deltat / 2.0 : 0.0;
        for (int32_t r = 1; r <= nrec; r++) {
            char name[96];
            snprintf(name, sizeof name, fmt, r);
            FILE *f = fopen(join(dir, name).c_str(), "w");
            if (!f) return CPML_EINVAL;
            for (int32_t it = 1; it <= nt; it++)
                fprintf(f, " %s   %s\n", gf_r4((double)(it - 1) * deltat - t0 + shift).c_str(),
                        gf_r4(sis[(size_t)(r - 1) * nt + (it - 1)]).c_str());
            fclose(f);
        }
    }
    return CPML_OK;
}

extern "C" int32_t cpml_host_write_energy_3d(const char *path, const double *total, int32_t nt, double deltat)
{
    if (!path || !total) return CPML_EINVAL;
    FILE *f = fopen(path, "w");
    if (!f) return CPML_EINVAL;
    for (int32_t it = 1; it <= nt; it++)                                  // :1254-1256
        fprintf(f, " %s %s\n", gf_r4((double)(it - 1) * deltat).c_str(), gf_r8(total[it - 1]).c_str());
    fclose(f);
    return CPML_OK;
}

extern "C" int32_t cpml_host_write_energy_2d(const char *path, const double *kinetic, const double *potential,
                                             int32_t nt, double deltat)
{
    if (!path || !kinetic || !potential) return CPML_EINVAL;
    FILE *f = fopen(path, "w");
    if (!f) return CPML_EINVAL;
    for (int32_t it = 1; it <= nt; it++)                                  // 2D-2nd :742-745
        fprintf(f, " %s %s %s %s\n", gf_r4((double)(it - 1) * deltat).c_str(), gf_r4(kinetic[it - 1]).c_str(),
                gf_r4(potential[it - 1]).c_str(), gf_r4(kinetic[it - 1] + potential[it - 1]).c_str());
    fclose(f);
    return CPML_OK;
}

extern "C" int32_t cpml_host_create_color_image(const char *dir, const double *img, int32_t nx, int32_t ny,
                                                int32_t it, int32_t isource, int32_t jsource,
                                                const int32_t *ix_rec, const int32_t *iy_rec, int32_t nrec,
                                                int32_t npoints_pml, int32_t use_pml_xmin, int32_t use_pml_xmax,
                                                int32_t use_pml_ymin, int32_t use_pml_ymax, int32_t field_number)
{
    if (!img || nx < 1 || ny < 1 || (field_number != 1 && field_number != 2)) return CPML_EINVAL;
    // display constants of create_color_image, 3D-iso :1377-1386
    const double power_display = 0.30, cutvect = 0.01;
    const int width_cross = 5, thickness_cross = 1, size_square = 3;

    char name[64];
    snprintf(name, sizeof name, field_number == 1 ? "image%06d_Vx.pnm" : "image%06d_Vy.pnm", it);   // :1406,1409
    FILE *f = fopen(join(dir, name).c_str(), "w");
    if (!f) return CPML_EINVAL;
    fprintf(f, "P3\n%d %d\n255\n", nx, ny);                                // :1415-1418

    double max_amplitude = 0.0;                                           // :1421
    for (size_t q = 0; q < (size_t)nx * ny; q++) max_amplitude = std::max(max_amplitude, std::fabs(img[q]));

    auto near = [](int v, int c, int w) { return v >= c - w && v <= c + w; };
    for (int iy = ny; iy >= 1; iy--) {                                    // top row first, :1424
        for (int ix = 1; ix <= nx; ix++) {
            const double v = img[(size_t)(iy - 1) * nx + (ix - 1)];
            double nv = v / max_amplitude;
            nv = std::min(1.0, std::max(-1.0, nv));
            int R, G, B;
            if ((near(ix, isource, width_cross) && near(iy, jsource, thickness_cross)) ||
                (near(ix, isource, thickness_cross) && near(iy, jsource, width_cross))) {
                R = 255; G = 157; B = 0;                                  // source cross
            } else if (ix <= 2 || ix >= nx - 1 || iy <= 2 || iy >= ny - 1) {
                R = G = B = 0;                                            // frame
            } else if ((use_pml_xmin && ix == npoints_pml) || (use_pml_xmax && ix == nx - npoints_pml) ||
                       (use_pml_ymin && iy == npoints_pml) || (use_pml_ymax && iy == ny - npoints_pml)) {
                R = 255; G = 150; B = 0;                                  // PML edges
            } else if (std::fabs(v) <= max_amplitude * cutvect) {
                R = G = B = 255;                                          // white background
            } else if (nv >= 0.0) {
                R = (int)std::lround(255.0 * std::pow(nv, power_display)); G = 0; B = 0;
            } else {
                R = 0; G = 0; B = (int)std::lround(255.0 * std::pow(std::fabs(nv), power_display));
            }
            for (int r = 0; r < nrec; r++)
                if (near(ix, ix_rec[r], size_square) && near(iy, iy_rec[r], size_square)) { R = 30; G = 180; B = 60; }
            fprintf(f, "%3d %3d %3d\n", R, G, B);                          // :1498
        }
    }
    fclose(f);
    return CPML_OK;
}


// plot_energy / plotgnu (/ plot_comparison) -- the Gnuplot scripts the programs leave next to their output files
// (3D-iso :1260-1313, 2D-2nd :748-806; the fourth-order and viscoelastic programs write the same text).  Each line
// is a list-directed character record (one leading blank), `write(20,*)` alone an empty record.  The receiver list is
// the reference's own hard-coded one (001 and 002), whatever NREC is.  program: 0 = 3-D (plotgnu also names the Vz
// files, which the 3-D reference never writes -- quirk B7 -- and this library does), 1 = 2-D.
extern "C" int32_t cpml_host_write_gnuplot_scripts(const char *dir, int32_t program)
{
    if (!dir || (program != 0 && program != 1)) return CPML_EINVAL;
    auto put = [](FILE *f, const char *line) { if (line[0]) fprintf(f, " %s\n", line); else fprintf(f, "\n"); };
    const bool d3 = program == 0;
    FILE *f = fopen(join(dir, "plot_energy").c_str(), "w");
    if (!f) return CPML_EINVAL;
    put(f, "# set term x11");
    put(f, "set term postscript landscape monochrome dashed \"Helvetica\" 22");
    put(f, "");
    put(f, "set xlabel \"Time (s)\"");
    put(f, "set ylabel \"Total energy\"");
    put(f, "");
    put(f, d3 ? "set output \"CPML3D_total_energy_semilog.eps\"" : "set output \"cpml_total_energy_semilog.eps\"");
    put(f, "set logscale y");
    put(f, d3 ? "plot \"energy.dat\" t 'Total energy' w l lc 1"
              : "plot \"energy.dat\" us 1:2 t 'Ec' w l lc 1, \"energy.dat\" us 1:3  t 'Ep' w l lc 3, \"energy.dat\" us 1:4 t 'Total energy' w l lc 4");
    put(f, "pause -1 \"Hit any key...\"");
    put(f, "");
    fclose(f);
    if (!d3) {
        f = fopen(join(dir, "plot_comparison").c_str(), "w");
        if (!f) return CPML_EINVAL;
        put(f, "# set term x11");
        put(f, "set term postscript landscape monochrome dashed \"Helvetica\" 22");
        put(f, "");
        put(f, "set xlabel \"Time (s)\"");
        put(f, "set ylabel \"Total energy\"");
        put(f, "");
        put(f, "set output \"compare_total_energy_semilog.eps\"");
        put(f, "set logscale y");
        put(f, "plot \"energy.dat\" us 1:4 t 'Total energy CPML' w l lc 1,  \"../collino/energy.dat\" us 1:4 t 'Total energy Collino' w l lc 2");
        put(f, "pause -1 \"Hit any key...\"");
        put(f, "");
        fclose(f);
    }
    f = fopen(join(dir, "plotgnu").c_str(), "w");
    if (!f) return CPML_EINVAL;
    put(f, "set term x11");
    put(f, "# set term postscript landscape monochrome dashed \"Helvetica\" 22");
    put(f, "");
    put(f, "set xlabel \"Time (s)\"");
    put(f, "set ylabel \"Amplitude (m / s)\"");
    put(f, "");
    for (int r = 1; r <= 2; r++)
        for (const char *c : {"Vx", "Vy", "Vz"}) {
            if (!d3 && c[1] == 'z') continue;
            char line[128];
            snprintf(line, sizeof line, "set output \"v_sigma_%s_receiver_%03d.eps\"", c, r);
            put(f, line);
            snprintf(line, sizeof line, "plot \"%s_file_%03d.dat\" t '%s C-PML' w l lc 1", c, r, c);
            put(f, line);
            put(f, "pause -1 \"Hit any key...\"");
            put(f, "");
        }
    fclose(f);
    return CPML_OK;
}
