"""Oracle-side set-up of the bench workloads (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

bench.py's `--impl reference` / `cpu_baseline` legs build their inputs here, with the ORACLE's own set-up
functions (oracle/cpml_oracle.c: profiles, source series, receiver search), so that the CPU arm never touches
the product library.  The parameter blocks restate the reference programs' `parameter` constants:
  3-D isotropic   seismic_CPML_3D_isotropic_MPI_OpenMP.f90:124-218
  2-D isotropic   seismic_CPML_2D_isotropic_{second,fourth}_order.f90:138-218
  3-D visco       seismic_CPML_3D_viscoelastic_MPI.f90:152-244
  2-D visco       seismic_CPML_2D_velocity_and_stress_fourth_order_viscoelastic.f90:140-230
"""
from __future__ import annotations

import math

import numpy as np

from . import oracle as O

PI = 3.141592653589793238462643

# fixed relaxation times the reference itself quotes (3D-visco :402-413; analytical program :124-128): the
# run time of the loop does not depend on their values, and they bypass the SolvOpt fit
TAU_CARCIONE_1993 = dict(tau_epsilon_nu1=(0.0334, 0.0028), tau_sigma_nu1=(0.0303, 0.0025),
                         tau_epsilon_nu2=(0.0352, 0.0029), tau_sigma_nu2=(0.0287, 0.0024))
TAU_2D_VISCO = dict(tau_epsilon_nu1=(2.408158185753685e-002, 4.699608990861351e-003, 9.567997872435925e-004),
                    tau_sigma_nu1=(2.256014638636808e-002, 4.508471279712252e-003, 8.937876403768840e-004),
                    tau_epsilon_nu2=(2.430544480527216e-002, 4.728107829226396e-003, 9.667252695863502e-004),
                    tau_sigma_nu2=(2.250919779429490e-002, 4.501388007338097e-003, 8.917332095369118e-004))


def iso3d(nx, ny, nz, nstep, dx=10.0, dt=1.6e-3, npml=10, nrec=2, ydeb=2300.0, yfin=300.0):
    """Arguments of oracle.run_3d_iso for the 3-D isotropic program on an nx x ny x nz grid (:124-218)."""
    cp, rho, f0 = 3300.0, 2800.0, 7.0
    cs = cp / 1.732
    t0 = 1.2 / f0
    amax = 2.0 * PI * (f0 / 2.0)
    px = O.pml_profile(nx, dx, dt, npml, cp=cp, alpha_max_pml=amax, clamp_alpha=True)
    py = O.pml_profile(ny, dx, dt, npml, cp=cp, alpha_max_pml=amax)
    pz = O.pml_profile(nz, dx, dt, npml, cp=cp, alpha_max_pml=amax)
    fx, fy = O.source_series(nstep, dt, f0, t0, 1e7, 135.0)
    isrc, jsrc = nx - 2 * npml - 1, 2 * ny // 3 + 1
    xs = (isrc - 1) * dx
    ydeb, yfin = min(ydeb, (ny - 1) * dx), min(yfin, (ny - 1) * dx)
    ix, iy, _ = O.find_receivers(nx, ny, dx, dx, nrec, xs - 100.0, ydeb, xs, yfin)
    return dict(nx=nx, ny=ny, nz=nz, deltax=dx, deltay=dx, deltaz=dx, deltat=dt,
                lam=rho * (cp * cp - 2.0 * cs * cs), mu=rho * cs * cs, lambdaplustwomu=rho * cp * cp, rho=rho,
                nstep=nstep, npoints_pml=npml, isource=isrc, jsource=jsrc,
                prof_x=px, prof_y=py, prof_z=pz, force_x=fx, force_y=fy, ix_rec=ix, iy_rec=iy)


def iso2d(order, nx, ny, nstep, dx=10.0, npml=10, nrec=2):
    """Arguments of oracle.run_2d (homogeneous medium of the shipped programs)."""
    dt = 2e-3 if order == 2 else 2e-3 / 2
    cp, rho, f0 = 3300.0, 2800.0, 7.0
    cs = cp / 1.732
    t0 = 1.2 / f0
    amax = 2.0 * PI * (f0 / 2.0)
    px = O.pml_profile(nx, dx, dt, npml, cp=cp, alpha_max_pml=amax, clamp_alpha=True)
    py = O.pml_profile(ny, dx, dt, npml, cp=cp, alpha_max_pml=amax, origin_top_uses_n=(order == 4))
    fx, fy = O.source_series(nstep, dt, f0, t0, 1e7, 135.0)
    isrc, jsrc = nx - 2 * npml - 1, 2 * ny // 3 + 1
    xs = (isrc - 1) * dx
    ix, iy, _ = O.find_receivers(nx, ny, dx, dx, nrec, xs - 100.0, min(2300.0, (ny - 1) * dx), xs, 300.0)
    return dict(order=order, nx=nx, ny=ny, deltax=dx, deltay=dx, deltat=dt, nstep=nstep, npoints_pml=npml,
                isource=isrc, jsource=jsrc, lam=np.full(nx * ny, rho * (cp * cp - 2.0 * cs * cs)),
                mu=np.full(nx * ny, rho * cs * cs), rho=np.full(nx * ny, rho), prof_x=px, prof_y=py,
                force_x=fx, force_y=fy, ix_rec=ix, iy_rec=iy)


def visco3d(nx, ny, nz, nstep, dx=4.0, dt=4e-4, npml=10):
    """Arguments of oracle.run_3d_visco (3D-visco :152-244), relaxation times of Carcione (1993)."""
    tau = dict(TAU_CARCIONE_1993)
    cp, cs, rho, f0 = 3000.0, 2000.0, 2000.0, 18.0
    t0 = 1.2 / f0
    amax = 2.0 * PI * (f0 / 2.0)
    inv = [tau["tau_epsilon_nu1"][0] / tau["tau_sigma_nu1"][0], tau["tau_epsilon_nu2"][0] / tau["tau_sigma_nu2"][0],
           tau["tau_epsilon_nu1"][1] / tau["tau_sigma_nu1"][1], tau["tau_epsilon_nu2"][1] / tau["tau_sigma_nu2"][1]]
    sq = math.sqrt(max(inv))                                       # :450-456, :547
    kw = dict(cp=cp, sqrt_taumax=sq, alpha_max_pml=amax)
    px = O.pml_profile_visco(nx, dx, dt, npml, clamp_alpha=True, **kw)
    py = O.pml_profile_visco(ny, dx, dt, npml, **kw)
    pz = O.pml_profile_visco(nz, dx, dt, npml, **kw)
    fx, fy = O.source_series(nstep, dt, f0, t0, 1e7, 0.0)
    isrc = min(npml + 20, nx - npml - 3)
    jsrc = ny // 5 + 1
    xs, ys = isrc * dx, jsrc * dx
    sc = min(1.0, (nx - isrc - 2) * dx / 500.0, (ny - jsrc - 2) * dx / 2260.0)
    ix, iy, _ = O.find_receivers_visco(nx, ny, dx, dx, [xs + 500.0 * sc, xs, xs + 500.0 * sc],
                                       [ys + 500.0 * sc, ys + 2260.0 * sc, ys + 2260.0 * sc])
    return dict(nx=nx, ny=ny, nz=nz, deltax=dx, deltay=dx, deltaz=dx, deltat=dt,
                lam=rho * (cp * cp - 2.0 * cs * cs), mu=rho * cs * cs, rho=rho, nstep=nstep, npoints_pml=npml,
                isource=isrc, jsource=jsrc, prof_x=px, prof_y=py, prof_z=pz, force_x=fx, force_y=fy,
                ix_rec=ix, iy_rec=iy, **tau)


def visco2d(order, nx, ny, nstep, dx=1.5, dt=2.2e-4, npml=10):
    """Arguments of oracle.run_2d_visco (2D-visco-4th :140-230), relaxation times of the analytical program."""
    tau = dict(TAU_2D_VISCO)
    cp, rho0, f0 = 2000.0, 2000.0, 35.0
    cs = cp / 1.732
    t0 = 1.2 / f0
    amax = 2.0 * PI * (f0 / 2.0)
    px = O.pml_profile(nx, dx, dt, npml, cp=cp, alpha_max_pml=amax, clamp_alpha=True)
    py = O.pml_profile(ny, dx, dt, npml, cp=cp, alpha_max_pml=amax)
    fx, fy = O.source_series_ricker(nstep, dt, f0, t0, 1.0, 0.0, dx, dx)
    isrc, jsrc = nx // 2 + 1, ny // 2 + 1
    xs, ys = (isrc - 1) * dx, (jsrc - 1) * dx
    ix, iy, _ = O.find_receivers(nx, ny, dx, dx, 2, xs + 20 * dx, ys + 20 * dx, xs + 10 * dx, ys - 25 * dx)
    mu = np.full(nx * ny, rho0 * cs * cs)
    return dict(order=order, nx=nx, ny=ny, deltax=dx, deltay=dx, deltat=dt, nstep=nstep, npoints_pml=npml,
                isource=isrc, jsource=jsrc, lam=rho0 * cp * cp - 2.0 * mu, mu=mu, rho=np.full(nx * ny, rho0),
                prof_x=px, prof_y=py, force_x=fx, force_y=fy, ix_rec=ix, iy_rec=iy, **tau)
