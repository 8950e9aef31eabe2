// seismic_cpml_driver.cpp -- C++ host driver over the C ABI (include/cpml_b200.h).
//
// Stands in for the Fortran host programs where no Fortran compiler exists: it keeps the
// reference programs' parameter surface (same names, same defaults), their screen output,
// display schedule and output files, and runs the time loop on the GPU.
//
//   xseismic_cpml --program 3d_iso      == seismic_CPML_3D_isotropic_MPI_OpenMP.f90
//   xseismic_cpml --program 2d_second   == seismic_CPML_2D_isotropic_second_order.f90
//   xseismic_cpml --program 2d_fourth   == seismic_CPML_2D_isotropic_fourth_order.f90
//   xseismic_cpml --program 3d_visco    == seismic_CPML_3D_viscoelastic_MPI.f90
//   xseismic_cpml --program 2d_visco_second | 2d_visco_fourth
//                                       == seismic_CPML_2D_velocity_and_stress_{second,fourth}_order_viscoelastic.f90
// (the viscoelastic programs also take NPROC= QKappa_att= QMu_att= f0_attenuation= / Qp= Qs= xsource=
//  ysource= COMPUTE_ENERGY= ; their relaxation times come from the SolvOpt fit like in the reference)
// Parameters are the Fortran `parameter` names given as NAME=value on the command line
// (the reference edits them in the source and recompiles): NX= NY= NZ= NSTEP= DELTAX=
// DELTAT= NPOINTS_PML= ISOURCE= JSOURCE= NREC= IT_DISPLAY= f0= factor= ANGLE_FORCE= cp= rho=
// xdeb= ydeb= xfin= yfin= K_MAX_PML= ; NGPU= decomposes the 3-D programs into that many z-slabs, one GPU each, the
// way the reference uses NPROC MPI ranks (SAME_DEVICE=1: all slabs on device 0); --out DIR selects the output directory,
// --no-images skips the PNM snapshots.
//
// Build: see drivers/Makefile (g++ -O2 -std=c++17 -Iinclude ... -lcpml_b200)
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "cpml_b200.h"

namespace {

constexpr double PI = 3.141592653589793238462643;
constexpr double STABILITY_THRESHOLD = 1.e25;

struct Args {
    std::string program = "3d_iso", out = ".";
    bool images = true;
    std::map<std::string, double> kv;
    double get(const char *k, double dflt) const { auto it = kv.find(k); return it == kv.end() ? dflt : it->second; }
    int geti(const char *k, int dflt) const { return (int)std::lround(get(k, dflt)); }
};


struct Profile {
    std::vector<double> a, b, K, a_half, b_half, K_half;
    explicit Profile(int n) : a(n), b(n), K(n), a_half(n), b_half(n), K_half(n) {}
};

// One simulation: a single handle (2-D programs; 3-D with NGPU=1) or, for the 3-D programs with NGPU >= 2, the
// whole z-slab decomposition behind a cpml_multi handle -- what the reference does with NPROC MPI ranks
// (3D-iso :337-346, :770-796), here from this one process.  NGPU= selects the number of slabs, one GPU each;
// SAME_DEVICE=1 puts every slab on device 0 (for boxes with a single GPU).
struct Sim {
    cpml_handle *h = nullptr;
    cpml_multi *m = nullptr;
    [[noreturn]] void die(const char *where) const
    {
        fprintf(stderr, " libcpml_b200 error in %s: %s\n", where, m ? cpml_multi_last_error(m) : cpml_last_error(h));
        exit(1);
    }
    void ck(int32_t rc, const char *where) const { if (rc != CPML_OK) die(where); }
    void create(const cpml_config &cfg, int ngpu, bool same_device)
    {
        if (cfg.ndim == 3 && ngpu >= 2) {
            std::vector<int32_t> dev(ngpu, 0);
            if (cpml_multi_create(&cfg, ngpu, same_device ? dev.data() : nullptr, &m) != CPML_OK) {
                fprintf(stderr, " libcpml_b200 error in cpml_multi_create: %s\n", cpml_multi_last_error(nullptr));
                exit(1);
            }
            printf(" z-slab decomposition over %d GPU slabs (NZ_LOCAL = %d)\n\n", ngpu, cfg.nz / ngpu);
        } else if (cpml_create(&cfg, &h) != CPML_OK) {
            fprintf(stderr, " libcpml_b200 error in cpml_create: %s\n", cpml_last_error(nullptr));
            exit(1);
        }
    }
    void set_profiles(int axis, const double *a, const double *b, const double *K, const double *ah, const double *bh, const double *Kh, int n)
    { ck(m ? cpml_multi_set_profiles(m, axis, a, b, K, ah, bh, Kh, n) : cpml_set_profiles(h, axis, a, b, K, ah, bh, Kh, n), "set_profiles"); }
    void set_material_2d(const double *lam, const double *mu, const double *rho) { ck(cpml_set_material_2d(h, lam, mu, rho), "set_material_2d"); }
    void set_attenuation(int n, const double *a, const double *b, const double *c, const double *d)
    { ck(m ? cpml_multi_set_attenuation(m, n, a, b, c, d) : cpml_set_attenuation(h, n, a, b, c, d), "set_attenuation"); }
    void set_source_series(const double *fx, const double *fy, int n)
    { ck(m ? cpml_multi_set_source_series(m, fx, fy, n) : cpml_set_source_series(h, fx, fy, n), "set_source_series"); }
    void set_receivers(const int32_t *ix, const int32_t *iy, int n)
    { ck(m ? cpml_multi_set_receivers(m, ix, iy, n) : cpml_set_receivers(h, ix, iy, n), "set_receivers"); }
    void run(int a, int b) { ck(m ? cpml_multi_run(m, a, b) : cpml_run(h, a, b), "run"); }
    void get_maxnorm(double *v) { ck(m ? cpml_multi_get_maxnorm(m, v) : cpml_get_maxnorm(h, v), "get_maxnorm"); }
    void get_energy(double *t, double *k, double *p) { ck(m ? cpml_multi_get_energy(m, t, k, p) : cpml_get_energy(h, t, k, p), "get_energy"); }
    void get_seismograms(double *sx, double *sy) { ck(m ? cpml_multi_get_seismograms(m, sx, sy) : cpml_get_seismograms(h, sx, sy), "get_seismograms"); }
    void get_seismograms_vz(double *sz) { ck(m ? cpml_multi_get_seismograms_vz(m, sz) : cpml_get_seismograms_vz(h, sz), "get_seismograms_vz"); }
    void get_pressure_seismograms(double *sp) { ck(cpml_get_pressure_seismograms(h, sp), "get_pressure_seismograms"); }
    void get_plane(int f, int k, double *out) { ck(m ? cpml_multi_get_plane(m, f, k, out) : cpml_get_plane(h, f, k, out), "get_plane"); }
    void destroy() { if (m) cpml_multi_destroy(m); else cpml_destroy(h); m = nullptr; h = nullptr; }
};

void set_profiles(Sim &S, int axis, const Profile &p, int n)
{
    S.set_profiles(axis, p.a.data(), p.b.data(), p.K.data(), p.a_half.data(), p.b_half.data(), p.K_half.data(), n);
}

// compute_attenuation_coeffs for the two modes (3D-visco :433-443, 2D-visco-4th :366-376)
void fit_attenuation(int n_sls, double q1, double q2, double f0_att, std::vector<double> (&tau)[4])
{
    const double f_min = std::exp(std::log(f0_att) - std::log(12.0) / 2.0), f_max = 12.0 * f_min;
    for (auto &t : tau) t.assign(n_sls, 0.0);
    double info[4];
    if (cpml_host_attenuation_fit(n_sls, q1, f0_att, f_min, f_max, tau[0].data(), tau[1].data(), info) != CPML_OK ||
        cpml_host_attenuation_fit(n_sls, q2, f0_att, f_min, f_max, tau[2].data(), tau[3].data(), info) != CPML_OK) {
        fprintf(stderr, "attenuation fit failed\n");
        exit(1);
    }
    const char *names[4] = {"tau_epsilon_nu1", "tau_sigma_nu1", "tau_epsilon_nu2", "tau_sigma_nu2"};
    for (int k = 0; k < 4; k++) {
        printf(" %s =", names[k]);
        for (double v : tau[k]) printf(" %.16g", v);
        printf("\n");
    }
    printf("\n");
}

// ---- the viscoelastic programs ------------------------------------------------------------------
int run_visco(const Args &A)
{
    const bool is3d = A.program == "3d_visco";
    const int order = is3d ? 4 : (A.program == "2d_visco_fourth" ? 4 : 2);
    // parameter block: 3D-visco :152-244 ; 2D-visco-4th :140-230
    const int NX = A.geti("NX", is3d ? 210 : 2001), NY = A.geti("NY", is3d ? 800 : 2001), NZ = is3d ? A.geti("NZ", 220) : 1;
    const int NPROC = A.geti("NPROC", 4);
    const double DELTAX = A.get("DELTAX", is3d ? 4.0 : 1.5), DELTAY = A.get("DELTAY", DELTAX), DELTAZ = A.get("DELTAZ", DELTAX);
    const double cp = A.get("cp", is3d ? 3000.0 : 2000.0), cs = A.get("cs", is3d ? 2000.0 : cp / 1.732), rho = A.get("rho", 2000.0);
    const int NSTEP = A.geti("NSTEP", is3d ? 100000 : 5200);
    const double DELTAT = A.get("DELTAT", is3d ? 4.e-4 : 2.2e-4);
    const double f0 = A.get("f0", is3d ? 18.0 : 35.0), t0 = A.get("t0", 1.20 / f0), factor = A.get("factor", is3d ? 1.e7 : 1.0);
    const int NPOINTS_PML = A.geti("NPOINTS_PML", 10);
    const double xs2d = A.get("xsource", 1500.0), ys2d = A.get("ysource", 1500.0);
    const int ISOURCE = A.geti("ISOURCE", is3d ? NPOINTS_PML + 20 : (int)(xs2d / DELTAX + 1));
    const int JSOURCE = A.geti("JSOURCE", is3d ? NY / 5 + 1 : (int)(ys2d / DELTAY + 1));
    const double ANGLE_FORCE = A.get("ANGLE_FORCE", 0.0);
    const int NREC = A.geti("NREC", is3d ? 3 : 1);
    const int IT_DISPLAY = A.geti("IT_DISPLAY", is3d ? 10000 : 200);
    const double NPOWER = 2.0, K_MAX_PML = A.get("K_MAX_PML", is3d ? 7.0 : 1.0), ALPHA_MAX_PML = 2.0 * PI * (f0 / 2.0);
    const double Rcoef = is3d ? 0.0001 : 0.001;
    const int N_SLS = is3d ? 2 : 3;
    const int COMPUTE_ENERGY = is3d ? 1 : A.geti("COMPUTE_ENERGY", 0);

    printf("\n %s viscoelastic finite-difference code in velocity and stress formulation with C-PML\n\n", is3d ? "3D" : "2D");
    printf(" NX = %d\n NY = %d\n", NX, NY);
    if (is3d) printf(" NZ = %d\n", NZ);
    printf("\n Total number of grid points = %lld\n\n", (long long)NX * NY * NZ);

    std::vector<double> tau[4];
    if (is3d) fit_attenuation(N_SLS, A.get("QKappa_att", 20.0), A.get("QMu_att", 10.0), A.get("f0_attenuation", 16.0), tau);
    else      fit_attenuation(N_SLS, A.get("Qp", 65.0), A.get("Qs", 55.0), f0, tau);

    Profile px(NX), py(NY), pz(is3d ? NZ : 1);
    double courant;
    double cp_cfl = 0.0;
    if (is3d) {
        double taumax = 0.0;                                               // :450-456
        for (int l = 0; l < N_SLS; l++) {
            taumax = std::max(taumax, 1.0 / (tau[1][l] / tau[0][l]));
            taumax = std::max(taumax, 1.0 / (tau[3][l] / tau[2][l]));
        }
        const double sq = std::sqrt(taumax);
        auto prof = [&](int n, double d, int clamp, Profile &p) {
            cpml_host_pml_profile_visco(n, d, DELTAT, NPOINTS_PML, 1, 1, cp, sq, Rcoef, NPOWER, K_MAX_PML, ALPHA_MAX_PML, 0, clamp,
                                        p.a.data(), p.b.data(), p.K.data(), p.a_half.data(), p.b_half.data(), p.K_half.data());
        };
        prof(NX, DELTAX, 1, px); prof(NY, DELTAY, 0, py); prof(NZ, DELTAZ, 0, pz);
        cp_cfl = cp * sq;
        courant = cpml_host_courant(cp_cfl, DELTAT, DELTAX, DELTAY, DELTAZ);                            // :856
        if (courant > 1.0) { fprintf(stderr, "time step is too large, simulation will be unstable\n"); return 1; }
    } else {
        auto prof = [&](int n, double d, int clamp, Profile &p) {
            cpml_host_pml_profile(n, d, DELTAT, NPOINTS_PML, 1, 1, cp, Rcoef, NPOWER, K_MAX_PML, ALPHA_MAX_PML, 0, clamp,
                                  p.a.data(), p.b.data(), p.K.data(), p.a_half.data(), p.b_half.data(), p.K_half.data());
        };
        prof(NX, DELTAX, 1, px); prof(NY, DELTAY, 0, py);
        courant = cp * DELTAT / DELTAX;                                                                 // :654
        if (DELTAX == DELTAY && courant > (order == 4 ? 0.606 : 1.0 / std::sqrt(2.0))) {
            fprintf(stderr, "time step is too large, simulation will be unstable\n"); return 1;
        }
    }
    printf(" Courant number is %.15g\n\n", courant);

    // source time function: first derivative of a Gaussian (3-D, :1310-1335) / Ricker over the cell area (2-D, :931-958)
    std::vector<double> force_x(NSTEP), force_y(NSTEP);
    if (is3d) cpml_host_source_series(NSTEP, DELTAT, f0, t0, factor, ANGLE_FORCE, force_x.data(), force_y.data());
    else {
        const double a = PI * PI * f0 * f0, rad = ANGLE_FORCE * (PI / 180.0);
        for (int it = 1; it <= NSTEP; it++) {
            const double t = (double)(it - 1) * DELTAT;
            double term = factor * (1.0 - 2.0 * a * ((t - t0) * (t - t0))) * std::exp(-a * ((t - t0) * (t - t0)));
            term = term / (DELTAX * DELTAY);
            force_x[it - 1] = std::sin(rad) * term;
            force_y[it - 1] = std::cos(rad) * term;
        }
    }
    // receivers: explicit targets (3-D, :825-853) / a line (2-D, :187-199)
    std::vector<int32_t> ix_rec(NREC), iy_rec(NREC);
    std::vector<double> dist(NREC);
    if (is3d) {
        const double xs = ISOURCE * DELTAX, ys = JSOURCE * DELTAY;                                      // :207-208
        std::vector<double> xr = {xs + 500.0, xs, xs + 500.0}, yr = {ys + 500.0, ys + 2260.0, ys + 2260.0};
        xr.resize(NREC, xs); yr.resize(NREC, ys);
        for (int r = 0; r < NREC; r++) {
            char kx[16], ky[16];
            snprintf(kx, sizeof kx, "xrec%d", r + 1); snprintf(ky, sizeof ky, "yrec%d", r + 1);
            xr[r] = A.get(kx, xr[r]); yr[r] = A.get(ky, yr[r]);
        }
        cpml_host_find_receivers_at(NX, NY, DELTAX, DELTAY, NREC, xr.data(), yr.data(), 1, ix_rec.data(), iy_rec.data(), dist.data());
    } else {
        cpml_host_find_receivers(NX, NY, DELTAX, DELTAY, NREC, A.get("xdeb", 2301.0), A.get("ydeb", 2301.0),
                                 A.get("xfin", 2301.0), A.get("yfin", 2301.0), ix_rec.data(), iy_rec.data(), dist.data());
    }
    for (int r = 0; r < NREC; r++)
        printf(" receiver %d closest grid point found at distance %g in i,j = %d %d\n", r + 1, dist[r], ix_rec[r], iy_rec[r]);
    printf("\n");

    cpml_config cfg{};
    cfg.ndim = is3d ? 3 : 2; cfg.order = order; cfg.rheology = 1;
    cfg.nx = NX; cfg.ny = NY; cfg.nz = NZ; cfg.nstep = NSTEP; cfg.npoints_pml = NPOINTS_PML; cfg.nrec = NREC;
    cfg.isource = ISOURCE; cfg.jsource = JSOURCE; cfg.ksource = 0; cfg.nslabs = 1; cfg.slab_rank = 0; cfg.device = -1;
    cfg.energy_bug_compat = 1;
    cfg.deltax = DELTAX; cfg.deltay = DELTAY; cfg.deltaz = DELTAZ; cfg.deltat = DELTAT;
    if (is3d) {
        cfg.emulate_nproc = NPROC;
        cfg.lambda = rho * (cp * cp - 2.0 * cs * cs); cfg.mu = rho * cs * cs; cfg.rho = rho; cfg.cp = cp_cfl;
    } else {
        cfg.compute_energy = COMPUTE_ENERGY;
    }
    Sim S;
    S.create(cfg, A.geti("NGPU", 1), A.geti("SAME_DEVICE", 0) != 0);
    set_profiles(S, CPML_AXIS_X, px, NX);
    set_profiles(S, CPML_AXIS_Y, py, NY);
    if (is3d) set_profiles(S, CPML_AXIS_Z, pz, NZ);
    else {
        const size_t n = (size_t)NX * NY;                       // homogeneous unrelaxed medium, :596-602
        const double mu = rho * cs * cs;
        std::vector<double> lam(n, rho * cp * cp - 2.0 * mu), muv(n, mu), rh(n, rho);
        S.set_material_2d(lam.data(), muv.data(), rh.data());
    }
    S.set_attenuation(N_SLS, tau[0].data(), tau[1].data(), tau[2].data(), tau[3].data());
    S.set_source_series(force_x.data(), force_y.data(), NSTEP);
    S.set_receivers(ix_rec.data(), iy_rec.data(), NREC);

    std::vector<double> sisvx((size_t)NSTEP * NREC), sisvy((size_t)NSTEP * NREC), sisp((size_t)NSTEP * NREC);
    std::vector<double> e_tot(NSTEP), e_kin(NSTEP), e_pot(NSTEP), plane((size_t)NX * NY);
    const auto t_start = std::chrono::steady_clock::now();
    auto write_all = [&]() {
        S.get_seismograms(sisvx.data(), sisvy.data());
        if (!is3d) S.get_pressure_seismograms(sisp.data());
        cpml_host_write_seismograms_visco(A.out.c_str(), sisvx.data(), sisvy.data(), is3d ? nullptr : sisp.data(), NSTEP, NREC, DELTAT, t0);
        if (is3d) {     // Vz_file_NNN.dat: an extension, the reference records Vx and Vy only
            S.get_seismograms_vz(sisp.data());
            cpml_host_write_seismograms_vz(A.out.c_str(), sisp.data(), NSTEP, NREC, DELTAT, t0);
        }
    };
    int it_begin = 1;
    while (it_begin <= NSTEP) {
        int it_end = std::min(NSTEP, (it_begin / IT_DISPLAY + 1) * IT_DISPLAY);
        if (it_begin <= 5 && it_end > 5) it_end = 5;
        S.run(it_begin, it_end);
        const int it = it_end;
        if (it % IT_DISPLAY == 0 || it == 5) {
            double vnorm = 0.0;
            S.get_maxnorm(&vnorm);
            const double tcpu = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
            printf(" Time step # %d out of %d\n Time: %g seconds\n Max norm velocity vector V (m/s) = %.15g\n", it, NSTEP,
                   (double)(float)((it - 1) * DELTAT), vnorm);
            if (COMPUTE_ENERGY) {
                S.get_energy(e_tot.data(), e_kin.data(), e_pot.data());
                printf(" Total energy = %.15g\n", e_tot[it - 1]);
            }
            printf(" Elapsed time in seconds = %g\n Mean elapsed time per time step in seconds = %g\n\n", tcpu, tcpu / it);
            if (vnorm > STABILITY_THRESHOLD || !std::isfinite(vnorm)) { fprintf(stderr, "code became unstable and blew up\n"); return 1; }
            if (is3d) cpml_host_write_timestamp(A.out.c_str(), it, DELTAT, vnorm, e_tot[it - 1], tcpu);   // 3D-visco :1469-1479
            write_all();
            if (A.images)
                for (int f = 0; f < 2; f++) {
                    S.get_plane(f, is3d ? NZ / 2 : 0, plane.data());
                    cpml_host_create_color_image(A.out.c_str(), plane.data(), NX, NY, it, ISOURCE, JSOURCE, ix_rec.data(),
                                                 iy_rec.data(), NREC, NPOINTS_PML, 1, 1, 1, 1, f + 1);
                }
        }
        it_begin = it_end + 1;
    }
    write_all();
    if (COMPUTE_ENERGY) {
        S.get_energy(e_tot.data(), e_kin.data(), e_pot.data());
        const std::string epath = A.out + "/energy.dat";
        cpml_host_write_energy_2d(epath.c_str(), e_kin.data(), e_pot.data(), NSTEP, DELTAT);   // time, kinetic, potential, total
    }
    cpml_host_write_gnuplot_scripts(A.out.c_str(), is3d ? 0 : 1);          // plot_energy, plotgnu (3D-iso :1260-1313)
    const double tcpu = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
    printf(" Total elapsed time = %g s, %.3f Gpts/s\n", tcpu, (double)NX * NY * NZ * NSTEP / tcpu / 1e9);
    S.destroy();
    printf("\n End of the simulation\n\n");
    return 0;
}

}  // namespace

int main(int argc, char **argv)
{
    Args A;
    for (int i = 1; i < argc; i++) {
        std::string s = argv[i];
        if (s == "--program" && i + 1 < argc) A.program = argv[++i];
        else if (s == "--out" && i + 1 < argc) A.out = argv[++i];
        else if (s == "--no-images") A.images = false;
        else if (s.find('=') != std::string::npos) A.kv[s.substr(0, s.find('='))] = atof(s.c_str() + s.find('=') + 1);
        else { fprintf(stderr, "unknown argument %s\n", s.c_str()); return 2; }
    }
    if (A.program == "3d_visco" || A.program == "2d_visco_second" || A.program == "2d_visco_fourth") return run_visco(A);
    const bool is3d = A.program == "3d_iso";
    const bool fourth = A.program == "2d_fourth";
    if (!is3d && !fourth && A.program != "2d_second") { fprintf(stderr, "unknown program %s\n", A.program.c_str()); return 2; }

    // ---- parameter block (3D-iso :124-218 ; 2D-2nd :138-218 ; 2D-4th :132-219)
    const int NX = A.geti("NX", 101), NY = A.geti("NY", 641), NZ = is3d ? A.geti("NZ", 640) : 1;
    const double DELTAX = A.get("DELTAX", 10.0), DELTAY = A.get("DELTAY", DELTAX), DELTAZ = A.get("DELTAZ", DELTAX);
    const double cp = A.get("cp", 3300.0), cs = A.get("cs", cp / 1.732), rho = A.get("rho", 2800.0);
    const int NSTEP = A.geti("NSTEP", is3d ? 2500 : fourth ? 4000 : 2000);
    const double DELTAT = A.get("DELTAT", is3d ? 1.6e-3 : fourth ? 2.e-3 / 2 : 2.e-3);
    const double f0 = A.get("f0", 7.0), t0 = A.get("t0", 1.20 / f0), factor = A.get("factor", 1.e7);
    const int NPOINTS_PML = A.geti("NPOINTS_PML", 10);
    const int ISOURCE = A.geti("ISOURCE", NX - 2 * NPOINTS_PML - 1), JSOURCE = A.geti("JSOURCE", 2 * NY / 3 + 1);
    const double xsource = (ISOURCE - 1) * DELTAX, ysource = (JSOURCE - 1) * DELTAY;
    const double ANGLE_FORCE = A.get("ANGLE_FORCE", 135.0);
    const int NREC = A.geti("NREC", 2);
    const double xdeb = A.get("xdeb", xsource - 100.0), ydeb = A.get("ydeb", 2300.0);
    const double xfin = A.get("xfin", xsource), yfin = A.get("yfin", 300.0);
    const int IT_DISPLAY = A.geti("IT_DISPLAY", fourth ? 200 : 100);
    const double NPOWER = 2.0, K_MAX_PML = A.get("K_MAX_PML", 1.0), ALPHA_MAX_PML = 2.0 * PI * (f0 / 2.0), Rcoef = 0.001;
    const int use_pml = 1;

    printf("\n %s elastic finite-difference code in velocity and stress formulation with C-PML\n\n", is3d ? "3D" : "2D");
    printf(" NX = %d\n NY = %d\n", NX, NY);
    if (is3d) printf(" NZ = %d\n", NZ);
    printf("\n size of the model along X = %g\n size of the model along Y = %g\n", (NX - 1) * DELTAX, (NY - 1) * DELTAY);
    printf("\n Total number of grid points = %lld\n\n", (long long)NX * NY * NZ);

    // ---- set-up phase
    Profile px(NX), py(NY), pz(is3d ? NZ : 1);
    cpml_host_pml_profile(NX, DELTAX, DELTAT, NPOINTS_PML, use_pml, use_pml, cp, Rcoef, NPOWER, K_MAX_PML, ALPHA_MAX_PML,
                          0, 1, px.a.data(), px.b.data(), px.K.data(), px.a_half.data(), px.b_half.data(), px.K_half.data());
    cpml_host_pml_profile(NY, DELTAY, DELTAT, NPOINTS_PML, use_pml, use_pml, cp, Rcoef, NPOWER, K_MAX_PML, ALPHA_MAX_PML,
                          fourth ? 1 : 0, 0, py.a.data(), py.b.data(), py.K.data(), py.a_half.data(), py.b_half.data(), py.K_half.data());
    if (is3d)
        cpml_host_pml_profile(NZ, DELTAZ, DELTAT, NPOINTS_PML, use_pml, use_pml, cp, Rcoef, NPOWER, K_MAX_PML, ALPHA_MAX_PML,
                              0, 0, pz.a.data(), pz.b.data(), pz.K.data(), pz.a_half.data(), pz.b_half.data(), pz.K_half.data());
    printf(" d0_x = %.15g\n\n", -(NPOWER + 1) * cp * std::log(Rcoef) / (2.0 * NPOINTS_PML * DELTAX));
    printf(" Position of the source:\n\n x = %g\n y = %g\n\n", xsource, ysource);

    std::vector<double> force_x(NSTEP), force_y(NSTEP);
    cpml_host_source_series(NSTEP, DELTAT, f0, t0, factor, ANGLE_FORCE, force_x.data(), force_y.data());
    std::vector<int32_t> ix_rec(NREC), iy_rec(NREC);
    std::vector<double> dist(NREC);
    printf(" There are %d receivers\n\n", NREC);
    cpml_host_find_receivers(NX, NY, DELTAX, DELTAY, NREC, xdeb, ydeb, xfin, yfin, ix_rec.data(), iy_rec.data(), dist.data());
    for (int r = 0; r < NREC; r++)
        printf(" receiver %d closest grid point found at distance %g in i,j = %d %d\n", r + 1, dist[r], ix_rec[r], iy_rec[r]);
    const double courant = cpml_host_courant(cp, DELTAT, DELTAX, DELTAY, is3d ? DELTAZ : 0.0);
    printf("\n Courant number is %.15g\n\n", courant);
    if (courant > 1.0) { fprintf(stderr, "time step is too large, simulation will be unstable\n"); return 1; }

    // ---- hand over to the GPU
    cpml_config cfg{};
    cfg.ndim = is3d ? 3 : 2; cfg.order = fourth ? 4 : 2;
    cfg.nx = NX; cfg.ny = NY; cfg.nz = NZ; cfg.nstep = NSTEP; cfg.npoints_pml = NPOINTS_PML; cfg.nrec = NREC;
    cfg.isource = ISOURCE; cfg.jsource = JSOURCE; cfg.ksource = 0; cfg.nslabs = 1; cfg.slab_rank = 0; cfg.device = -1;
    cfg.energy_bug_compat = 1;
    cfg.deltax = DELTAX; cfg.deltay = DELTAY; cfg.deltaz = DELTAZ; cfg.deltat = DELTAT;
    cfg.lambda = rho * (cp * cp - 2.0 * cs * cs); cfg.mu = rho * cs * cs; cfg.lambdaplustwomu = rho * cp * cp;
    cfg.rho = rho; cfg.cp = cp;
    Sim S;
    S.create(cfg, A.geti("NGPU", 1), A.geti("SAME_DEVICE", 0) != 0);
    S.set_profiles(CPML_AXIS_X, px.a.data(), px.b.data(), px.K.data(), px.a_half.data(), px.b_half.data(), px.K_half.data(), NX);
    S.set_profiles(CPML_AXIS_Y, py.a.data(), py.b.data(), py.K.data(), py.a_half.data(), py.b_half.data(), py.K_half.data(), NY);
    if (is3d)
        S.set_profiles(CPML_AXIS_Z, pz.a.data(), pz.b.data(), pz.K.data(), pz.a_half.data(), pz.b_half.data(), pz.K_half.data(), NZ);
    else {
        const size_t n = (size_t)NX * NY;     // homogeneous medium of 2D-2nd :468-474
        std::vector<double> lam(n, cfg.lambda), mu(n, cfg.mu), rh(n, rho);
        S.set_material_2d(lam.data(), mu.data(), rh.data());
    }
    S.set_source_series(force_x.data(), force_y.data(), NSTEP);
    S.set_receivers(ix_rec.data(), iy_rec.data(), NREC);

    std::vector<double> sisvx((size_t)NSTEP * NREC), sisvy((size_t)NSTEP * NREC);
    std::vector<double> e_tot(NSTEP), e_kin(NSTEP), e_pot(NSTEP), plane((size_t)NX * NY);
    const auto t_start = std::chrono::steady_clock::now();

    // ---- time loop: the GPU runs to the next display step (:1183), then the driver outputs
    int it_begin = 1;
    while (it_begin <= NSTEP) {
        int it_end = std::min(NSTEP, (it_begin / IT_DISPLAY + 1) * IT_DISPLAY);
        if (it_begin <= 5 && it_end > 5) it_end = 5;
        S.run(it_begin, it_end);
        const int it = it_end;
        if (it % IT_DISPLAY == 0 || it == 5) {
            double vnorm = 0.0;
            S.get_maxnorm(&vnorm);
            S.get_energy(e_tot.data(), e_kin.data(), e_pot.data());
            const double tcpu = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
            printf(" Time step # %d out of %d\n Time: %g seconds\n Max norm velocity vector V (m/s) = %.15g\n"
                   " Total energy = %.15g\n Elapsed time in seconds = %g\n Mean elapsed time per time step in seconds = %g\n\n",
                   it, NSTEP, (double)(float)((it - 1) * DELTAT), vnorm, e_tot[it - 1], tcpu, tcpu / it);
            if (vnorm > STABILITY_THRESHOLD || !std::isfinite(vnorm)) { fprintf(stderr, "code became unstable and blew up\n"); return 1; }
            if (is3d) cpml_host_write_timestamp(A.out.c_str(), it, DELTAT, vnorm, e_tot[it - 1], tcpu);   // :1219-1229
            S.get_seismograms(sisvx.data(), sisvy.data());
            cpml_host_write_seismograms(A.out.c_str(), sisvx.data(), sisvy.data(), NSTEP, NREC, DELTAT);
            if (A.images)
                for (int f = 0; f < 2; f++) {
                    S.get_plane(f, is3d ? NZ / 2 : 0, plane.data());
                    cpml_host_create_color_image(A.out.c_str(), plane.data(), NX, NY, it, ISOURCE, JSOURCE, ix_rec.data(),
                                                 iy_rec.data(), NREC, NPOINTS_PML, use_pml, use_pml, use_pml, use_pml, f + 1);
                }
        }
        it_begin = it_end + 1;
    }

    // ---- final output (:1247-1257 ; 2D-2nd :737-746)
    S.get_seismograms(sisvx.data(), sisvy.data());
    cpml_host_write_seismograms(A.out.c_str(), sisvx.data(), sisvy.data(), NSTEP, NREC, DELTAT);
    if (is3d) {         // Vz_file_NNN.dat: an extension, the reference records Vx and Vy only
        S.get_seismograms_vz(sisvx.data());
        cpml_host_write_seismograms_vz(A.out.c_str(), sisvx.data(), NSTEP, NREC, DELTAT, 0.0);
    }
    S.get_energy(e_tot.data(), e_kin.data(), e_pot.data());
    const std::string epath = A.out + "/energy.dat";
    if (is3d) cpml_host_write_energy_3d(epath.c_str(), e_tot.data(), NSTEP, DELTAT);
    else      cpml_host_write_energy_2d(epath.c_str(), e_kin.data(), e_pot.data(), NSTEP, DELTAT);
    cpml_host_write_gnuplot_scripts(A.out.c_str(), is3d ? 0 : 1);          // plot_energy, plotgnu (3D-iso :1260-1313)
    const double tcpu = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
    printf(" Total elapsed time = %g s, %.3f Gpts/s\n", tcpu, (double)NX * NY * NZ * NSTEP / tcpu / 1e9);
    S.destroy();
    printf("\n End of the simulation\n\n");
    return 0;
}
