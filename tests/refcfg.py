"""Reference configurations used by the tests, built with the ORACLE's own set-up
functions (oracle/cpml_oracle.c) so that oracle tests do not depend on the product."""
import math

import numpy as np

from oracle import oracle as O

PI = 3.141592653589793238462643


def cfg2d(order=2, nx=101, ny=641, nstep=None, npml=10, material="homogeneous", nrec=2,
          ydeb=2300.0, yfin=300.0, k_max=1.0):
    """seismic_CPML_2D_isotropic_{second,fourth}_order.f90 defaults (:138-218)."""
    dx = 10.0
    dt = 2e-3 if order == 2 else 2e-3 / 2
    if nstep is None:
        nstep = 2000 if order == 2 else 4000
    cp = 3300.0
    cs = cp / 1.732
    rho = 2800.0
    f0 = 7.0
    t0 = 1.2 / f0
    amax = 2.0 * PI * (f0 / 2.0)
    px = O.pml_profile(nx, dx, dt, npml, cp=cp, alpha_max_pml=amax, clamp_alpha=True, k_max_pml=k_max)
    py = O.pml_profile(ny, dx, dt, npml, cp=cp, alpha_max_pml=amax, origin_top_uses_n=(order == 4),
                       k_max_pml=k_max)
    fx, fy = O.source_series(nstep, dt, f0, t0, 1e7, 135.0)
    isrc = nx - 2 * npml - 1
    jsrc = 2 * ny // 3 + 1
    xs = (isrc - 1) * dx
    ix, iy, _ = O.find_receivers(nx, ny, dx, dx, nrec, xs - 100.0, ydeb, xs, yfin)
    lam = np.full((ny, nx), rho * (cp * cp - 2.0 * cs * cs))
    mu = np.full((ny, nx), rho * cs * cs)
    r = np.full((ny, nx), rho)
    if material == "layered":     # slower, lighter layer in the upper third + a smooth gradient in x
        jj, ii = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
        scale = np.where(jj > 2 * ny // 3, 0.64, 1.0) * (1.0 + 0.1 * ii / nx)
        lam, mu = lam * scale, mu * scale
        r = r * np.where(jj > 2 * ny // 3, 0.9, 1.0)
    return dict(order=order, nx=nx, ny=ny, deltax=dx, deltay=dx, deltat=dt, nstep=nstep,
                npoints_pml=npml, isource=isrc, jsource=jsrc, lam=lam.ravel(), mu=mu.ravel(),
                rho=r.ravel(), prof_x=px, prof_y=py, force_x=fx, force_y=fy, ix_rec=ix, iy_rec=iy)


def cfg3d(nx=37, ny=45, nz=40, nstep=150, npml=6, nrec=2, dt=1.6e-3, k_max=1.0):
    """seismic_CPML_3D_isotropic_MPI_OpenMP.f90 (:124-218) on a reduced grid."""
    dx = 10.0
    cp = 3300.0
    cs = cp / 1.732
    rho = 2800.0
    f0 = 7.0
    t0 = 1.2 / f0
    amax = 2.0 * PI * (f0 / 2.0)
    px = O.pml_profile(nx, dx, dt, npml, cp=cp, alpha_max_pml=amax, clamp_alpha=True, k_max_pml=k_max)
    py = O.pml_profile(ny, dx, dt, npml, cp=cp, alpha_max_pml=amax, k_max_pml=k_max)
    pz = O.pml_profile(nz, dx, dt, npml, cp=cp, alpha_max_pml=amax, k_max_pml=k_max)
    fx, fy = O.source_series(nstep, dt, f0, t0, 1e7, 135.0)
    isrc = nx - 2 * npml - 1
    jsrc = 2 * ny // 3 + 1
    xs = (isrc - 1) * dx
    ix, iy, _ = O.find_receivers(nx, ny, dx, dx, nrec, xs - 100.0, (ny // 3) * dx, xs, 3 * dx)
    return dict(nx=nx, ny=ny, nz=nz, deltax=dx, deltay=dx, deltaz=dx, deltat=dt,
                lam=rho * (cp * cp - 2.0 * cs * cs), mu=rho * cs * cs, lambdaplustwomu=rho * cp * cp,
                rho=rho, nstep=nstep, npoints_pml=npml, isource=isrc, jsource=jsrc,
                prof_x=px, prof_y=py, prof_z=pz, force_x=fx, force_y=fy, ix_rec=ix, iy_rec=iy)


def rel_l2(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = math.sqrt(float(np.sum(b * b)))
    num = math.sqrt(float(np.sum((a - b) ** 2)))
    return num / den if den > 0 else num


# Carcione (1993) two-mechanism relaxation times quoted in the reference as fixed alternatives to
# the SolvOpt fit (seismic_CPML_3D_viscoelastic_MPI.f90:402-413): Qkappa ~ 20, Qmu ~ 10.
TAU_CARCIONE_1993 = dict(tau_epsilon_nu1=(0.0334, 0.0028), tau_sigma_nu1=(0.0303, 0.0025),
                         tau_epsilon_nu2=(0.0352, 0.0029), tau_sigma_nu2=(0.0287, 0.0024))


def visco_taumax_taumin(tau):
    """3D-visco :450-456."""
    tau1 = tau["tau_sigma_nu1"][0] / tau["tau_epsilon_nu1"][0]
    tau2 = tau["tau_sigma_nu2"][0] / tau["tau_epsilon_nu2"][0]
    tau3 = tau["tau_sigma_nu1"][1] / tau["tau_epsilon_nu1"][1]
    tau4 = tau["tau_sigma_nu2"][1] / tau["tau_epsilon_nu2"][1]
    inv = [1.0 / tau1, 1.0 / tau2, 1.0 / tau3, 1.0 / tau4]
    return max(inv), min(inv)


def cfgv3d(nx=38, ny=46, nz=40, nstep=120, npml=6, dt=4e-4, tau=None, rec_scale=None):
    """seismic_CPML_3D_viscoelastic_MPI.f90 (:152-244) on a reduced grid, relaxation times given."""
    tau = dict(TAU_CARCIONE_1993 if tau is None else tau)
    dx = 4.0
    cp, cs, rho = 3000.0, 2000.0, 2000.0
    f0 = 18.0
    t0 = 1.2 / f0
    amax = 2.0 * PI * (f0 / 2.0)
    taumax, _ = visco_taumax_taumin(tau)
    sq = math.sqrt(taumax)
    kw = dict(cp=cp, sqrt_taumax=sq, alpha_max_pml=amax)
    px = O.pml_profile_visco(nx, dx, dt, npml, clamp_alpha=True, **kw)
    py = O.pml_profile_visco(ny, dx, dt, npml, **kw)
    pz = O.pml_profile_visco(nz, dx, dt, npml, **kw)
    fx, fy = O.source_series(nstep, dt, f0, t0, 1e7, 0.0)           # ANGLE_FORCE = 0 (:212)
    isrc = min(npml + 20, nx - npml - 3)                            # :205
    jsrc = ny // 5 + 1                                              # :206
    xs, ys = isrc * dx, jsrc * dx                                   # :207-208
    sc = rec_scale if rec_scale is not None else min(1.0, (nx - isrc - 2) * dx / 500.0, (ny - jsrc - 2) * dx / 2260.0)
    xrec = [xs + 500.0 * sc, xs, xs + 500.0 * sc]                   # :832-837
    yrec = [ys + 500.0 * sc, ys + 2260.0 * sc, ys + 2260.0 * sc]
    ix, iy, _ = O.find_receivers_visco(nx, ny, dx, dx, xrec, yrec)
    return dict(nx=nx, ny=ny, nz=nz, deltax=dx, deltay=dx, deltaz=dx, deltat=dt,
                lam=rho * (cp * cp - 2.0 * cs * cs), mu=rho * cs * cs, rho=rho, nstep=nstep,
                npoints_pml=npml, isource=isrc, jsource=jsrc, prof_x=px, prof_y=py, prof_z=pz,
                force_x=fx, force_y=fy, ix_rec=ix, iy_rec=iy, cp_eff=cp * sq, **tau)


# Relaxation times of the N_SLS = 3 Zener solids hard-coded in the reference's analytical-solution
# program for Qp = 65 / Qs = 55 style media (analytical_solution_viscoelastic_2D_plane_strain_Carcione_
# correct_with_1_over_L.f90:124-128: "classical least squares" constants, f0 = 35 Hz); used here as fixed
# inputs that bypass the SolvOpt fit.
TAU_2D_VISCO = dict(tau_epsilon_nu1=(2.408158185753685e-002, 4.699608990861351e-003, 9.567997872435925e-004),
                    tau_sigma_nu1=(2.256014638636808e-002, 4.508471279712252e-003, 8.937876403768840e-004),
                    tau_epsilon_nu2=(2.430544480527216e-002, 4.728107829226396e-003, 9.667252695863502e-004),
                    tau_sigma_nu2=(2.250919779429490e-002, 4.501388007338097e-003, 8.917332095369118e-004))


def cfgv2d(order=4, nx=81, ny=97, nstep=200, npml=8, tau=None, material="homogeneous", k_max=1.0, dt=2.2e-4):
    """seismic_CPML_2D_velocity_and_stress_{second,fourth}_order_viscoelastic.f90 (:140-230) on a reduced grid."""
    tau = dict(TAU_2D_VISCO if tau is None else tau)
    dx = 1.5
    cp, rho0 = 2000.0, 2000.0
    cs = cp / 1.732
    f0 = 35.0
    t0 = 1.2 / f0
    amax = 2.0 * PI * (f0 / 2.0)
    px = O.pml_profile(nx, dx, dt, npml, cp=cp, alpha_max_pml=amax, clamp_alpha=True, k_max_pml=k_max)
    py = O.pml_profile(ny, dx, dt, npml, cp=cp, alpha_max_pml=amax, k_max_pml=k_max)
    fx, fy = O.source_series_ricker(nstep, dt, f0, t0, 1.0, 0.0, dx, dx)
    isrc, jsrc = nx // 2 + 1, ny // 2 + 1
    xs, ys = (isrc - 1) * dx, (jsrc - 1) * dx
    ix, iy, _ = O.find_receivers(nx, ny, dx, dx, 2, xs + 20 * dx, ys + 20 * dx, xs + 10 * dx, ys - 25 * dx)
    mu = np.full((ny, nx), rho0 * cs * cs)
    lam = rho0 * cp * cp - 2.0 * mu
    r = np.full((ny, nx), rho0)
    if material == "layered":
        jj, ii = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
        scale = np.where(jj > 2 * ny // 3, 0.7, 1.0) * (1.0 + 0.1 * ii / nx)
        lam, mu = lam * scale, mu * scale
        r = r * np.where(jj > 2 * ny // 3, 0.9, 1.0)
    return dict(order=order, nx=nx, ny=ny, deltax=dx, deltay=dx, deltat=dt, nstep=nstep, npoints_pml=npml,
                isource=isrc, jsource=jsrc, lam=lam.ravel(), mu=mu.ravel(), rho=r.ravel(), prof_x=px, prof_y=py,
                force_x=fx, force_y=fy, ix_rec=ix, iy_rec=iy, **tau)
