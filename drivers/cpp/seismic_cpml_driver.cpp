// seismic_cpml_driver.cpp -- C++ host driver over the C ABI (include/cpml_b200.h).
//
// Stands in for the Fortran host programs where no Fortran compiler exists: it keeps the
// reference programs' parameter surface (same names, same defaults), their screen output,
// display schedule and output files, and runs the time loop on the GPU.
//
//   xseismic_cpml --program 3d_iso      == seismic_CPML_3D_isotropic_MPI_OpenMP.f90
//   xseismic_cpml --program 2d_second   == seismic_CPML_2D_isotropic_second_order.f90
//   xseismic_cpml --program 2d_fourth   == seismic_CPML_2D_isotropic_fourth_order.f90
// Parameters are the Fortran `parameter` names given as NAME=value on the command line
// (the reference edits them in the source and recompiles): NX= NY= NZ= NSTEP= DELTAX=
// DELTAT= NPOINTS_PML= ISOURCE= JSOURCE= NREC= IT_DISPLAY= f0= factor= ANGLE_FORCE= cp= rho=
// xdeb= ydeb= xfin= yfin= K_MAX_PML= ; --out DIR selects the output directory,
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

[[noreturn]] void die(cpml_handle *h, const char *where)
{
    fprintf(stderr, " libcpml_b200 error in %s: %s\n", where, cpml_last_error(h));
    exit(1);
}
#define CHECK(h, call) do { if ((call) != CPML_OK) die(h, #call); } while (0)

struct Profile {
    std::vector<double> a, b, K, a_half, b_half, K_half;
    explicit Profile(int n) : a(n), b(n), K(n), a_half(n), b_half(n), K_half(n) {}
};

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
    cpml_handle *h = nullptr;
    CHECK(nullptr, cpml_create(&cfg, &h));
    CHECK(h, cpml_set_profiles(h, CPML_AXIS_X, px.a.data(), px.b.data(), px.K.data(), px.a_half.data(), px.b_half.data(), px.K_half.data(), NX));
    CHECK(h, cpml_set_profiles(h, CPML_AXIS_Y, py.a.data(), py.b.data(), py.K.data(), py.a_half.data(), py.b_half.data(), py.K_half.data(), NY));
    if (is3d)
        CHECK(h, cpml_set_profiles(h, CPML_AXIS_Z, pz.a.data(), pz.b.data(), pz.K.data(), pz.a_half.data(), pz.b_half.data(), pz.K_half.data(), NZ));
    else {
        const size_t n = (size_t)NX * NY;     // homogeneous medium of 2D-2nd :468-474
        std::vector<double> lam(n, cfg.lambda), mu(n, cfg.mu), rh(n, rho);
        CHECK(h, cpml_set_material_2d(h, lam.data(), mu.data(), rh.data()));
    }
    CHECK(h, cpml_set_source_series(h, force_x.data(), force_y.data(), NSTEP));
    CHECK(h, cpml_set_receivers(h, ix_rec.data(), iy_rec.data(), NREC));

    std::vector<double> sisvx((size_t)NSTEP * NREC), sisvy((size_t)NSTEP * NREC);
    std::vector<double> e_tot(NSTEP), e_kin(NSTEP), e_pot(NSTEP), plane((size_t)NX * NY);
    const auto t_start = std::chrono::steady_clock::now();

    // ---- time loop: the GPU runs to the next display step (:1183), then the driver outputs
    int it_begin = 1;
    while (it_begin <= NSTEP) {
        int it_end = std::min(NSTEP, (it_begin / IT_DISPLAY + 1) * IT_DISPLAY);
        if (it_begin <= 5 && it_end > 5) it_end = 5;
        CHECK(h, cpml_run(h, it_begin, it_end));
        const int it = it_end;
        if (it % IT_DISPLAY == 0 || it == 5) {
            double vnorm = 0.0;
            CHECK(h, cpml_get_maxnorm(h, &vnorm));
            CHECK(h, cpml_get_energy(h, e_tot.data(), e_kin.data(), e_pot.data()));
            const double tcpu = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
            printf(" Time step # %d out of %d\n Time: %g seconds\n Max norm velocity vector V (m/s) = %.15g\n"
                   " Total energy = %.15g\n Elapsed time in seconds = %g\n Mean elapsed time per time step in seconds = %g\n\n",
                   it, NSTEP, (double)(float)((it - 1) * DELTAT), vnorm, e_tot[it - 1], tcpu, tcpu / it);
            if (vnorm > STABILITY_THRESHOLD || !std::isfinite(vnorm)) { fprintf(stderr, "code became unstable and blew up\n"); return 1; }
            CHECK(h, cpml_get_seismograms(h, sisvx.data(), sisvy.data()));
            cpml_host_write_seismograms(A.out.c_str(), sisvx.data(), sisvy.data(), NSTEP, NREC, DELTAT);
            if (A.images)
                for (int f = 0; f < 2; f++) {
                    CHECK(h, cpml_get_plane(h, f, is3d ? NZ / 2 : 0, plane.data()));
                    cpml_host_create_color_image(A.out.c_str(), plane.data(), NX, NY, it, ISOURCE, JSOURCE, ix_rec.data(),
                                                 iy_rec.data(), NREC, NPOINTS_PML, use_pml, use_pml, use_pml, use_pml, f + 1);
                }
        }
        it_begin = it_end + 1;
    }

    // ---- final output (:1247-1257 ; 2D-2nd :737-746)
    CHECK(h, cpml_get_seismograms(h, sisvx.data(), sisvy.data()));
    cpml_host_write_seismograms(A.out.c_str(), sisvx.data(), sisvy.data(), NSTEP, NREC, DELTAT);
    CHECK(h, cpml_get_energy(h, e_tot.data(), e_kin.data(), e_pot.data()));
    const std::string epath = A.out + "/energy.dat";
    if (is3d) cpml_host_write_energy_3d(epath.c_str(), e_tot.data(), NSTEP, DELTAT);
    else      cpml_host_write_energy_2d(epath.c_str(), e_kin.data(), e_pot.data(), NSTEP, DELTAT);
    const double tcpu = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
    printf(" Total elapsed time = %g s, %.3f Gpts/s\n", tcpu, (double)NX * NY * NZ * NSTEP / tcpu / 1e9);
    cpml_destroy(h);
    printf("\n End of the simulation\n\n");
    return 0;
}
