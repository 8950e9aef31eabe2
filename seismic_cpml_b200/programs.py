"""Host-side mirror of the reference programs: same parameter names, same set-up phase,
same loop structure, same outputs -- with the loop body replaced by calls through the
C ABI (libcpml_b200.so).

  Program3DIso   <-> seismic_CPML_3D_isotropic_MPI_OpenMP.f90
  Program2DIso   <-> seismic_CPML_2D_isotropic_second_order.f90 (order=2)
                     seismic_CPML_2D_isotropic_fourth_order.f90 (order=4)
  Program3DVisco <-> seismic_CPML_3D_viscoelastic_MPI.f90
  Program2DVisco <-> seismic_CPML_2D_velocity_and_stress_second_order_viscoelastic.f90 (order=2)
                     seismic_CPML_2D_velocity_and_stress_fourth_order_viscoelastic.f90 (order=4)

The reference configures itself through compile-time `parameter` constants; here they
are dataclass fields with the Fortran names and the Fortran defaults.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field

import numpy as np

from . import lib as _lib

PI = 3.141592653589793238462643          # 3D-iso :199
STABILITY_THRESHOLD = 1.0e25             # 3D-iso :211


class UnstableError(RuntimeError):
    """'code became unstable and blew up' (3D-iso :1195, 2D-2nd :726)."""


@dataclass
class Params3DIso:
    """Parameter block of seismic_CPML_3D_isotropic_MPI_OpenMP.f90:124-218."""
    NX: int = 101
    NY: int = 641
    NZ: int = 640
    DELTAX: float = 10.0
    DELTAY: float | None = None          # = DELTAX (:135)
    DELTAZ: float | None = None
    cp: float = 3300.0
    cs: float | None = None              # = cp / 1.732 (:140)
    rho: float = 2800.0
    NSTEP: int = 2500
    DELTAT: float = 1.6e-3
    f0: float = 7.0
    t0: float | None = None              # = 1.20 / f0 (:154)
    factor: float = 1.0e7
    USE_PML_XMIN: bool = True
    USE_PML_XMAX: bool = True
    USE_PML_YMIN: bool = True
    USE_PML_YMAX: bool = True
    USE_PML_ZMIN: bool = True
    USE_PML_ZMAX: bool = True
    NPOINTS_PML: int = 10
    ISOURCE: int | None = None           # = NX - 2*NPOINTS_PML - 1 (:181)
    JSOURCE: int | None = None           # = 2*NY/3 + 1 (:182)
    KSOURCE: int = 0                     # 0 = NZ/2, the reference's cut plane (:346)
    ANGLE_FORCE: float = 135.0
    NREC: int = 2
    xdeb: float | None = None            # = xsource - 100 (:190)
    ydeb: float = 2300.0
    xfin: float | None = None            # = xsource (:192)
    yfin: float = 300.0
    IT_DISPLAY: int = 100
    NPOWER: float = 2.0
    K_MAX_PML: float = 1.0
    ALPHA_MAX_PML: float | None = None   # = 2 pi (f0/2) (:218)
    Rcoef: float = 0.001                 # :407
    energy_bug_compat: bool = True       # quirk B2 (:1169-1172)

    def __post_init__(self):
        if self.DELTAY is None: self.DELTAY = self.DELTAX
        if self.DELTAZ is None: self.DELTAZ = self.DELTAX
        if self.cs is None: self.cs = self.cp / 1.732
        if self.t0 is None: self.t0 = 1.20 / self.f0
        if self.ISOURCE is None: self.ISOURCE = self.NX - 2 * self.NPOINTS_PML - 1
        if self.JSOURCE is None: self.JSOURCE = 2 * self.NY // 3 + 1
        if self.xdeb is None: self.xdeb = self.xsource - 100.0
        if self.xfin is None: self.xfin = self.xsource
        if self.ALPHA_MAX_PML is None: self.ALPHA_MAX_PML = 2.0 * PI * (self.f0 / 2.0)

    # derived constants, :142-144, :183-184
    @property
    def mu(self): return self.rho * self.cs * self.cs
    @property
    def lam(self): return self.rho * (self.cp * self.cp - 2.0 * self.cs * self.cs)
    @property
    def lambdaplustwomu(self): return self.rho * self.cp * self.cp
    @property
    def xsource(self): return (self.ISOURCE - 1) * self.DELTAX
    @property
    def ysource(self): return (self.JSOURCE - 1) * self.DELTAY


@dataclass
class Params2DIso:
    """Parameter block of seismic_CPML_2D_isotropic_{second,fourth}_order.f90:138-218."""
    order: int = 2
    NX: int = 101
    NY: int = 641
    DELTAX: float = 10.0
    DELTAY: float | None = None
    USE_PML_XMIN: bool = True
    USE_PML_XMAX: bool = True
    USE_PML_YMIN: bool = True
    USE_PML_YMAX: bool = True
    NPOINTS_PML: int = 10
    cp: float = 3300.0
    cs: float | None = None
    density: float = 2800.0
    NSTEP: int | None = None             # 2000 (2nd) / 4000 (4th, 2D-4th :156)
    DELTAT: float | None = None          # 2e-3 (2nd) / 1e-3 (4th, 2D-4th :162)
    f0: float = 7.0
    t0: float | None = None
    factor: float = 1.0e7
    ISOURCE: int | None = None
    JSOURCE: int | None = None
    ANGLE_FORCE: float = 135.0
    NREC: int = 2
    xdeb: float | None = None
    ydeb: float = 2300.0
    xfin: float | None = None
    yfin: float = 300.0
    IT_DISPLAY: int | None = None        # 100 (2nd) / 200 (4th, 2D-4th :187)
    NPOWER: float = 2.0
    K_MAX_PML: float = 1.0
    ALPHA_MAX_PML: float | None = None
    Rcoef: float = 0.001
    # quirk B4: the fourth-order program puts the top PML origin at NY*DELTAY - L
    # (2D-4th :401); None = follow the program selected by `order`
    yorigintop_uses_NY: bool | None = None

    def __post_init__(self):
        if self.order not in (2, 4):
            raise ValueError("order must be 2 or 4")
        fourth = self.order == 4
        if self.DELTAY is None: self.DELTAY = self.DELTAX
        if self.cs is None: self.cs = self.cp / 1.732
        if self.NSTEP is None: self.NSTEP = 4000 if fourth else 2000
        if self.DELTAT is None: self.DELTAT = 2.0e-3 / 2 if fourth else 2.0e-3
        if self.t0 is None: self.t0 = 1.20 / self.f0
        if self.ISOURCE is None: self.ISOURCE = self.NX - 2 * self.NPOINTS_PML - 1
        if self.JSOURCE is None: self.JSOURCE = 2 * self.NY // 3 + 1
        if self.xdeb is None: self.xdeb = self.xsource - 100.0
        if self.xfin is None: self.xfin = self.xsource
        if self.IT_DISPLAY is None: self.IT_DISPLAY = 200 if fourth else 100
        if self.ALPHA_MAX_PML is None: self.ALPHA_MAX_PML = 2.0 * PI * (self.f0 / 2.0)
        if self.yorigintop_uses_NY is None: self.yorigintop_uses_NY = fourth

    @property
    def xsource(self): return (self.ISOURCE - 1) * self.DELTAX
    @property
    def ysource(self): return (self.JSOURCE - 1) * self.DELTAY


# Two-mechanism relaxation times the reference quotes as fixed alternatives to its SolvOpt fit
# (3D-visco :402-413, Carcione 1993: Qkappa ~ 20, Qmu ~ 10) -- the same attenuation the program's
# QKappa_att = 20, QMu_att = 10 ask SolvOpt for (:192).
TAU_CARCIONE_1993 = dict(tau_epsilon_nu1=(0.0334, 0.0028), tau_sigma_nu1=(0.0303, 0.0025),
                         tau_epsilon_nu2=(0.0352, 0.0029), tau_sigma_nu2=(0.0287, 0.0024))


def _fit_missing_tau(p, q_nu1, q_nu2, f0_att):
    """The two compute_attenuation_coeffs calls of the viscoelastic programs (3D-visco :433-443,
    2D-visco-4th :366-376): mode nu1 (dilatation: QKappa, or Qp in the 2-D programs) and mode nu2
    (shear), each fitted over f_max / f_min = 12 around f0.  Only pairs left at None are fitted."""
    for mode, q in (("nu1", q_nu1), ("nu2", q_nu2)):
        te, ts = getattr(p, "tau_epsilon_" + mode), getattr(p, "tau_sigma_" + mode)
        if (te is None) != (ts is None):
            raise ValueError(f"give both tau_epsilon_{mode} and tau_sigma_{mode}, or neither")
        if te is None:
            te, ts = _lib.host_attenuation_fit(p.N_SLS, q, f0_att)
            setattr(p, "tau_epsilon_" + mode, te)
            setattr(p, "tau_sigma_" + mode, ts)


@dataclass
class Params3DVisco:
    """Parameter block of seismic_CPML_3D_viscoelastic_MPI.f90:152-244.  Relaxation times left at
    None are derived from QKappa_att, QMu_att, f0_attenuation by the SolvOpt fit like the reference
    does at :433-443 (`lib.host_attenuation_fit`); explicit tuples (e.g. TAU_CARCIONE_1993) bypass it."""
    NX: int = 210
    NY: int = 800
    NZ: int = 220
    NPROC: int = 4                       # :158 -- the incomplete halo exchange makes the result depend on it
    DELTAX: float = 4.0
    DELTAY: float | None = None
    DELTAZ: float | None = None
    cp: float = 3000.0
    cs: float = 2000.0
    rho: float = 2000.0
    NSTEP: int = 100000
    DELTAT: float = 4.0e-4
    f0: float = 18.0
    t0: float | None = None              # = 1.20 / f0 (:183)
    factor: float = 1.0e7
    N_SLS: int = 2
    QKappa_att: float = 20.0             # :191
    QMu_att: float = 10.0
    f0_attenuation: float = 16.0         # :192
    tau_epsilon_nu1: tuple | None = None
    tau_sigma_nu1: tuple | None = None
    tau_epsilon_nu2: tuple | None = None
    tau_sigma_nu2: tuple | None = None
    USE_PML_XMIN: bool = True
    USE_PML_XMAX: bool = True
    USE_PML_YMIN: bool = True
    USE_PML_YMAX: bool = True
    USE_PML_ZMIN: bool = True
    USE_PML_ZMAX: bool = True
    NPOINTS_PML: int = 10
    ISOURCE: int | None = None           # = NPOINTS_PML + 20 (:205)
    JSOURCE: int | None = None           # = NY / 5 + 1 (:206)
    KSOURCE: int = 0
    ANGLE_FORCE: float = 0.0
    NREC: int = 3
    xrec: tuple | None = None            # :832-837
    yrec: tuple | None = None
    IT_DISPLAY: int = 10000
    NPOWER: float = 2.0
    K_MAX_PML: float = 7.0
    ALPHA_MAX_PML: float | None = None
    Rcoef: float = 0.0001                # :541
    energy_bug_compat: bool = True
    sigmazz_isotropic: bool = False      # True: the isotropic memory-variable term in sigmazz, NOT the reference (quirk B14)

    def __post_init__(self):
        if self.N_SLS != 2:
            raise ValueError("the reference loop is written for N_SLS = 2")
        _fit_missing_tau(self, self.QKappa_att, self.QMu_att, self.f0_attenuation)
        if self.DELTAY is None: self.DELTAY = self.DELTAX
        if self.DELTAZ is None: self.DELTAZ = self.DELTAX
        if self.t0 is None: self.t0 = 1.20 / self.f0
        if self.ISOURCE is None: self.ISOURCE = self.NPOINTS_PML + 20
        if self.JSOURCE is None: self.JSOURCE = self.NY // 5 + 1
        if self.xrec is None: self.xrec = (self.xsource + 500.0, self.xsource, self.xsource + 500.0)
        if self.yrec is None: self.yrec = (self.ysource + 500.0, self.ysource + 2260.0, self.ysource + 2260.0)
        if self.ALPHA_MAX_PML is None: self.ALPHA_MAX_PML = 2.0 * PI * (self.f0 / 2.0)

    @property
    def mu(self): return self.rho * self.cs * self.cs                                   # :171
    @property
    def lam(self): return self.rho * (self.cp * self.cp - 2.0 * self.cs * self.cs)      # :172
    @property
    def xsource(self): return self.ISOURCE * self.DELTAX                                # :207
    @property
    def ysource(self): return self.JSOURCE * self.DELTAY
    @property
    def taumax(self):                                                                   # :450-455
        return max(self._inv_tau())
    @property
    def taumin(self):
        return min(self._inv_tau())

    def _inv_tau(self):
        tau1 = self.tau_sigma_nu1[0] / self.tau_epsilon_nu1[0]
        tau2 = self.tau_sigma_nu2[0] / self.tau_epsilon_nu2[0]
        tau3 = self.tau_sigma_nu1[1] / self.tau_epsilon_nu1[1]
        tau4 = self.tau_sigma_nu2[1] / self.tau_epsilon_nu2[1]
        return (1.0 / tau1, 1.0 / tau2, 1.0 / tau3, 1.0 / tau4)


# "Classical least-squares constants" for N_SLS = 3, Qp = 65, Qs = 55, f0 = 35 Hz that the reference
# hard-codes in its analytical-solution program (analytical_solution_viscoelastic_2D_plane_strain_
# Carcione_correct_with_1_over_L.f90:124-128) -- the medium of the 2-D viscoelastic programs.
TAU_2D_VISCO = dict(tau_epsilon_nu1=(2.408158185753685e-002, 4.699608990861351e-003, 9.567997872435925e-004),
                    tau_sigma_nu1=(2.256014638636808e-002, 4.508471279712252e-003, 8.937876403768840e-004),
                    tau_epsilon_nu2=(2.430544480527216e-002, 4.728107829226396e-003, 9.667252695863502e-004),
                    tau_sigma_nu2=(2.250919779429490e-002, 4.501388007338097e-003, 8.917332095369118e-004))


@dataclass
class Params2DVisco:
    """Parameter block of seismic_CPML_2D_velocity_and_stress_{second,fourth}_order_viscoelastic.f90
    (:140-230).  Relaxation times left at None are fitted to Qp, Qs around f0 with SolvOpt like the
    reference does at :366-376; the defaults reproduce TAU_2D_VISCO to the last digit."""
    order: int = 4
    VISCOELASTIC_ATTENUATION: bool = True
    NX: int = 2001
    NY: int = 2001
    DELTAX: float = 1.5
    DELTAY: float | None = None
    USE_PML_XMIN: bool = True
    USE_PML_XMAX: bool = True
    USE_PML_YMIN: bool = True
    USE_PML_YMAX: bool = True
    NPOINTS_PML: int = 10
    cp_unrelaxed: float = 2000.0
    cs_unrelaxed: float | None = None    # = cp_unrelaxed / 1.732 (:163)
    density: float = 2000.0
    DELTAT: float = 2.2e-4
    NSTEP: int = 5200
    f0: float = 35.0
    t0: float | None = None
    factor: float = 1.0
    xsource: float = 1500.0
    ysource: float = 1500.0
    ISOURCE: int | None = None           # = xsource / DELTAX + 1 (:187)
    JSOURCE: int | None = None
    ANGLE_FORCE: float = 0.0
    NREC: int = 1
    xdeb: float = 2301.0
    ydeb: float = 2301.0
    xfin: float = 2301.0
    yfin: float = 2301.0
    COMPUTE_ENERGY: bool = False
    IT_DISPLAY: int = 200
    NPOWER: float = 2.0
    K_MAX_PML: float = 1.0
    ALPHA_MAX_PML: float | None = None
    Rcoef: float = 0.001
    N_SLS: int = 3
    Qp: float = 65.0                     # :316-317
    Qs: float = 55.0
    tau_epsilon_nu1: tuple | None = None
    tau_sigma_nu1: tuple | None = None
    tau_epsilon_nu2: tuple | None = None
    tau_sigma_nu2: tuple | None = None

    def __post_init__(self):
        if self.order not in (2, 4):
            raise ValueError("order must be 2 or 4")
        if self.N_SLS != 3:
            raise ValueError("the 2-D viscoelastic programs use N_SLS = 3")
        if self.DELTAY is None: self.DELTAY = self.DELTAX
        if self.cs_unrelaxed is None: self.cs_unrelaxed = self.cp_unrelaxed / 1.732
        if self.t0 is None: self.t0 = 1.20 / self.f0
        if self.ISOURCE is None: self.ISOURCE = int(self.xsource / self.DELTAX + 1)
        if self.JSOURCE is None: self.JSOURCE = int(self.ysource / self.DELTAY + 1)
        if self.ALPHA_MAX_PML is None: self.ALPHA_MAX_PML = 2.0 * PI * (self.f0 / 2.0)
        if not self.VISCOELASTIC_ATTENUATION:        # dummy values of :374-380
            self.tau_epsilon_nu1 = self.tau_sigma_nu1 = self.tau_epsilon_nu2 = self.tau_sigma_nu2 = (1.0, 1.0, 1.0)
        _fit_missing_tau(self, self.Qp, self.Qs, self.f0)

    @property
    def cp(self): return self.cp_unrelaxed


@dataclass
class Setup:
    """What the reference builds before `do it = 1,NSTEP`."""
    prof_x: dict
    prof_y: dict
    prof_z: dict | None
    force_x: np.ndarray
    force_y: np.ndarray
    ix_rec: np.ndarray
    iy_rec: np.ndarray
    dist_rec: np.ndarray
    courant: float
    material: tuple | None = None        # 2-D: (lambda, mu, rho) arrays, NX*NY, i fastest
    log: list = field(default_factory=list)


def _profile(n, delta, p, use_min, use_max, *, top_uses_n=False, clamp=False):
    return _lib.host_pml_profile(n, delta, p.DELTAT, p.NPOINTS_PML, use_min, use_max, cp=p.cp,
                                 rcoef=p.Rcoef, npower=p.NPOWER, k_max_pml=p.K_MAX_PML,
                                 alpha_max_pml=p.ALPHA_MAX_PML, origin_top_uses_n=top_uses_n,
                                 clamp_alpha=clamp)


def setup_3d(p: Params3DIso) -> Setup:
    """3D-iso :399-717."""
    prof_x = _profile(p.NX, p.DELTAX, p, p.USE_PML_XMIN, p.USE_PML_XMAX, clamp=True)
    prof_y = _profile(p.NY, p.DELTAY, p, p.USE_PML_YMIN, p.USE_PML_YMAX)
    prof_z = _profile(p.NZ, p.DELTAZ, p, p.USE_PML_ZMIN, p.USE_PML_ZMAX)
    fx, fy = _lib.host_source_series(p.NSTEP, p.DELTAT, p.f0, p.t0, p.factor, p.ANGLE_FORCE)
    ix, iy, dist = _lib.host_find_receivers(p.NX, p.NY, p.DELTAX, p.DELTAY, p.NREC,
                                            p.xdeb, p.ydeb, p.xfin, p.yfin)
    courant = _lib.host_courant(p.cp, p.DELTAT, p.DELTAX, p.DELTAY, p.DELTAZ)
    if courant > 1.0:
        raise _lib.CpmlError(_lib.CPML_ECFL, "time step is too large, simulation will be unstable")
    return Setup(prof_x, prof_y, prof_z, fx, fy, ix, iy, dist, courant)


def setup_2d(p: Params2DIso, material=None) -> Setup:
    """2D-2nd :283-516 (2D-4th: same lines + 1)."""
    prof_x = _profile(p.NX, p.DELTAX, p, p.USE_PML_XMIN, p.USE_PML_XMAX, clamp=True)
    prof_y = _profile(p.NY, p.DELTAY, p, p.USE_PML_YMIN, p.USE_PML_YMAX, top_uses_n=p.yorigintop_uses_NY)
    fx, fy = _lib.host_source_series(p.NSTEP, p.DELTAT, p.f0, p.t0, p.factor, p.ANGLE_FORCE)
    ix, iy, dist = _lib.host_find_receivers(p.NX, p.NY, p.DELTAX, p.DELTAY, p.NREC,
                                            p.xdeb, p.ydeb, p.xfin, p.yfin)
    courant = _lib.host_courant(p.cp, p.DELTAT, p.DELTAX, p.DELTAY)
    if courant > 1.0:
        raise _lib.CpmlError(_lib.CPML_ECFL, "time step is too large, simulation will be unstable")
    if material is None:                 # homogeneous medium of 2D-2nd :468-474
        n = p.NX * p.NY
        material = (np.full(n, p.density * (p.cp * p.cp - 2.0 * p.cs * p.cs)),
                    np.full(n, p.density * p.cs * p.cs), np.full(n, p.density))
    return Setup(prof_x, prof_y, None, fx, fy, ix, iy, dist, courant, material=material)


def setup_3d_visco(p: Params3DVisco) -> Setup:
    """3D-visco :533-858."""
    sq = math.sqrt(p.taumax)
    kw = dict(cp=p.cp, sqrt_taumax=sq, rcoef=p.Rcoef, npower=p.NPOWER, k_max_pml=p.K_MAX_PML,
              alpha_max_pml=p.ALPHA_MAX_PML)
    prof_x = _lib.host_pml_profile_visco(p.NX, p.DELTAX, p.DELTAT, p.NPOINTS_PML, p.USE_PML_XMIN, p.USE_PML_XMAX,
                                         clamp_alpha=True, **kw)
    prof_y = _lib.host_pml_profile_visco(p.NY, p.DELTAY, p.DELTAT, p.NPOINTS_PML, p.USE_PML_YMIN, p.USE_PML_YMAX, **kw)
    prof_z = _lib.host_pml_profile_visco(p.NZ, p.DELTAZ, p.DELTAT, p.NPOINTS_PML, p.USE_PML_ZMIN, p.USE_PML_ZMAX, **kw)
    fx, fy = _lib.host_source_series(p.NSTEP, p.DELTAT, p.f0, p.t0, p.factor, p.ANGLE_FORCE)
    ix, iy, dist = _lib.host_find_receivers_at(p.NX, p.NY, p.DELTAX, p.DELTAY, p.xrec[:p.NREC], p.yrec[:p.NREC], 1)
    courant = _lib.host_courant(p.cp * sq, p.DELTAT, p.DELTAX, p.DELTAY, p.DELTAZ)      # :856
    if courant > 1.0:
        raise _lib.CpmlError(_lib.CPML_ECFL, "time step is too large, simulation will be unstable")
    return Setup(prof_x, prof_y, prof_z, fx, fy, ix, iy, dist, courant)


def host_source_series_ricker(p: "Params2DVisco"):
    """2D-visco-4th :931-958: Ricker wavelet divided by the area of a grid cell."""
    a = PI * PI * p.f0 * p.f0
    rad = p.ANGLE_FORCE * (PI / 180.0)
    fx, fy = np.zeros(p.NSTEP), np.zeros(p.NSTEP)
    for it in range(1, p.NSTEP + 1):
        t = float(it - 1) * p.DELTAT
        term = p.factor * (1.0 - 2.0 * a * ((t - p.t0) * (t - p.t0))) * math.exp(-a * ((t - p.t0) * (t - p.t0)))
        term = term / (p.DELTAX * p.DELTAY)
        fx[it - 1] = math.sin(rad) * term
        fy[it - 1] = math.cos(rad) * term
    return fx, fy


def setup_2d_visco(p: "Params2DVisco", material=None) -> Setup:
    """2D-visco-4th :401-660."""
    prof_x = _profile(p.NX, p.DELTAX, p, p.USE_PML_XMIN, p.USE_PML_XMAX, clamp=True)
    prof_y = _profile(p.NY, p.DELTAY, p, p.USE_PML_YMIN, p.USE_PML_YMAX)
    fx, fy = host_source_series_ricker(p)
    ix, iy, dist = _lib.host_find_receivers(p.NX, p.NY, p.DELTAX, p.DELTAY, p.NREC, p.xdeb, p.ydeb, p.xfin, p.yfin)
    courant = p.cp_unrelaxed * p.DELTAT / p.DELTAX                      # :654
    limit = 0.606 if p.order == 4 else 1.0 / math.sqrt(2.0)               # :656 / second-order file :650
    if p.DELTAX == p.DELTAY and courant > limit:
        raise _lib.CpmlError(_lib.CPML_ECFL, "time step is too large, simulation will be unstable")
    if material is None:                 # :596-602
        n = p.NX * p.NY
        mu = p.density * p.cs_unrelaxed * p.cs_unrelaxed
        material = (np.full(n, p.density * p.cp_unrelaxed * p.cp_unrelaxed - 2.0 * mu), np.full(n, mu), np.full(n, p.density))
    return Setup(prof_x, prof_y, None, fx, fy, ix, iy, dist, courant, material=material)


def make_solver_2d_visco(p: "Params2DVisco", s: Setup, *, device=-1) -> _lib.Solver:
    sol = _lib.Solver(ndim=2, order=p.order, rheology=1, compute_energy=p.COMPUTE_ENERGY, nx=p.NX, ny=p.NY,
                      nstep=p.NSTEP, npoints_pml=p.NPOINTS_PML, nrec=p.NREC, isource=p.ISOURCE, jsource=p.JSOURCE,
                      device=device, deltax=p.DELTAX, deltay=p.DELTAY, deltat=p.DELTAT, cp=0.0)
    sol.set_profiles(_lib.AXIS_X, s.prof_x)
    sol.set_profiles(_lib.AXIS_Y, s.prof_y)
    sol.set_material_2d(*s.material)
    sol.set_attenuation(p.tau_epsilon_nu1, p.tau_sigma_nu1, p.tau_epsilon_nu2, p.tau_sigma_nu2)
    sol.set_source_series(s.force_x, s.force_y)
    sol.set_receivers(s.ix_rec, s.iy_rec)
    return sol


def make_solver_3d_visco(p: Params3DVisco, s: Setup, *, nslabs=1, slab_rank=0, device=-1) -> _lib.Solver:
    sol = _lib.Solver(ndim=3, order=4, rheology=1, emulate_nproc=p.NPROC, nx=p.NX, ny=p.NY, nz=p.NZ,
                      nstep=p.NSTEP, npoints_pml=p.NPOINTS_PML, nrec=p.NREC, isource=p.ISOURCE,
                      jsource=p.JSOURCE, ksource=p.KSOURCE, nslabs=nslabs, slab_rank=slab_rank, device=device,
                      energy_bug_compat=p.energy_bug_compat, sigmazz_isotropic=p.sigmazz_isotropic,
                      deltax=p.DELTAX, deltay=p.DELTAY,
                      deltaz=p.DELTAZ, deltat=p.DELTAT, lam=p.lam, mu=p.mu, rho=p.rho,
                      cp=p.cp * math.sqrt(p.taumax))
    sol.set_profiles(_lib.AXIS_X, s.prof_x)
    sol.set_profiles(_lib.AXIS_Y, s.prof_y)
    sol.set_profiles(_lib.AXIS_Z, s.prof_z)
    sol.set_attenuation(p.tau_epsilon_nu1, p.tau_sigma_nu1, p.tau_epsilon_nu2, p.tau_sigma_nu2)
    sol.set_source_series(s.force_x, s.force_y)
    sol.set_receivers(s.ix_rec, s.iy_rec)
    return sol


def make_solver_3d(p: Params3DIso, s: Setup, *, nslabs=1, slab_rank=0, device=-1, precision=0) -> _lib.Solver:
    """precision=1: the single-precision build the reference endorses (3D-iso :114-116), one GPU only."""
    sol = _lib.Solver(ndim=3, order=2, nx=p.NX, ny=p.NY, nz=p.NZ, nstep=p.NSTEP, precision=precision,
                      npoints_pml=p.NPOINTS_PML, nrec=p.NREC, isource=p.ISOURCE, jsource=p.JSOURCE,
                      ksource=p.KSOURCE, nslabs=nslabs, slab_rank=slab_rank, device=device,
                      energy_bug_compat=p.energy_bug_compat, deltax=p.DELTAX, deltay=p.DELTAY,
                      deltaz=p.DELTAZ, deltat=p.DELTAT, lam=p.lam, mu=p.mu,
                      lambdaplustwomu=p.lambdaplustwomu, rho=p.rho, cp=p.cp)
    sol.set_profiles(_lib.AXIS_X, s.prof_x)
    sol.set_profiles(_lib.AXIS_Y, s.prof_y)
    sol.set_profiles(_lib.AXIS_Z, s.prof_z)
    sol.set_source_series(s.force_x, s.force_y)
    sol.set_receivers(s.ix_rec, s.iy_rec)
    return sol


def make_solver_2d(p: Params2DIso, s: Setup, *, device=-1) -> _lib.Solver:
    sol = _lib.Solver(ndim=2, order=p.order, nx=p.NX, ny=p.NY, nstep=p.NSTEP,
                      npoints_pml=p.NPOINTS_PML, nrec=p.NREC, isource=p.ISOURCE, jsource=p.JSOURCE,
                      device=device, deltax=p.DELTAX, deltay=p.DELTAY, deltat=p.DELTAT, cp=p.cp)
    sol.set_profiles(_lib.AXIS_X, s.prof_x)
    sol.set_profiles(_lib.AXIS_Y, s.prof_y)
    sol.set_material_2d(*s.material)
    sol.set_source_series(s.force_x, s.force_y)
    sol.set_receivers(s.ix_rec, s.iy_rec)
    return sol


class _ProgramBase:
    """The driver loop shared by the programs: run IT_DISPLAY steps on the GPU, then do
    what the reference does at `mod(it,IT_DISPLAY) == 0 .or. it == 5` (:1183)."""

    def __init__(self, params, setup, solver, output_dir=None, verbose=False):
        self.p, self.s, self.solver = params, setup, solver
        self.output_dir = output_dir
        self.verbose = verbose
        self.display_log = []            # (it, time, max norm, total energy)
        self._t_start = None

    def _display_steps(self):
        p = self.p
        stops = sorted({it for it in range(1, p.NSTEP + 1) if it % p.IT_DISPLAY == 0 or it == 5} | {p.NSTEP})
        return stops

    def _snapshot_fields(self):
        raise NotImplementedError

    def run(self, nstep=None):
        p = self.p
        last = p.NSTEP if nstep is None else min(nstep, p.NSTEP)
        it0 = 1
        import time
        self._t_start = time.time()
        for stop in self._display_steps():
            if stop > last:
                break
            self.solver.run(it0, stop)
            it0 = stop + 1
            if stop % p.IT_DISPLAY == 0 or stop == 5:
                self._display(stop)
        if it0 <= last:
            self.solver.run(it0, last)
        return self.results()

    def _display(self, it):
        p = self.p
        vnorm = self.solver.get_maxnorm()
        total = self._total_energy()[it - 1]
        self.display_log.append((it, (it - 1) * p.DELTAT, vnorm, total))
        if self.verbose:
            print(f" Time step # {it} out of {p.NSTEP}")
            print(f" Time: {np.float32((it - 1) * p.DELTAT)} seconds")
            print(f" Max norm velocity vector V (m/s) = {vnorm}")
            print(f" Total energy = {total}")
        if vnorm > STABILITY_THRESHOLD or not math.isfinite(vnorm):
            raise UnstableError("code became unstable and blew up")
        if self.output_dir is not None:
            os.makedirs(self.output_dir, exist_ok=True)
            L = _lib.load()
            if hasattr(p, "NZ"):         # timestampNNNNNN: the 3-D programs only (3D-iso :1219-1229)
                import time
                L.cpml_host_write_timestamp(self.output_dir.encode(), it, p.DELTAT, vnorm, float(total),
                                            time.time() - (self._t_start or time.time()))
            self.write_seismograms()
            vx, vy = self._snapshot_fields()
            for img, num in ((vx, 1), (vy, 2)):
                img = np.ascontiguousarray(img)
                L.cpml_host_create_color_image(self.output_dir.encode(), _lib._d(img), p.NX, p.NY, it,
                                               p.ISOURCE, p.JSOURCE, _lib._i(self.s.ix_rec),
                                               _lib._i(self.s.iy_rec), p.NREC, p.NPOINTS_PML,
                                               int(p.USE_PML_XMIN), int(p.USE_PML_XMAX),
                                               int(p.USE_PML_YMIN), int(p.USE_PML_YMAX), num)

    def write_seismograms(self):
        sx, sy = self.solver.get_seismograms()
        _lib.load().cpml_host_write_seismograms(self.output_dir.encode(), _lib._d(sx), _lib._d(sy),
                                                self.p.NSTEP, self.p.NREC, self.p.DELTAT)


class Program3DIso(_ProgramBase):
    """seismic_CPML_3D_isotropic_MPI_OpenMP.f90 on one GPU (whole grid, nslabs = 1)."""

    def __init__(self, params: Params3DIso | None = None, output_dir=None, verbose=False, device=-1):
        params = params or Params3DIso()
        s = setup_3d(params)
        super().__init__(params, s, make_solver_3d(params, s, device=device), output_dir, verbose)
        self.ksource = params.KSOURCE or params.NZ // 2

    def _total_energy(self):
        return self.solver.get_energy()[0]

    def _snapshot_fields(self):          # vx, vy(:,:,NZ_LOCAL) of the cut-plane rank, :1236-1239
        return self.solver.get_plane(0, self.ksource), self.solver.get_plane(1, self.ksource)

    def results(self):
        sx, sy = self.solver.get_seismograms()
        sz = self.solver.get_seismograms_vz()          # extension: the reference records Vx and Vy only (quirk B7)
        total = self._total_energy()
        if self.output_dir is not None:  # :1247-1257
            os.makedirs(self.output_dir, exist_ok=True)
            self.write_seismograms()
            _lib.load().cpml_host_write_seismograms_vz(self.output_dir.encode(), _lib._d(sz), self.p.NSTEP, self.p.NREC,
                                                       self.p.DELTAT, 0.0)
            _lib.load().cpml_host_write_gnuplot_scripts(self.output_dir.encode(), 0)       # :1260-1313
            _lib.load().cpml_host_write_energy_3d(os.path.join(self.output_dir, "energy.dat").encode(),
                                                  _lib._d(total), self.p.NSTEP, self.p.DELTAT)
        return dict(sisvx=sx, sisvy=sy, sisvz=sz, total_energy=total, display_log=self.display_log)


class Program3DVisco(_ProgramBase):
    """seismic_CPML_3D_viscoelastic_MPI.f90 on one GPU (whole grid; the reference's NPROC only
    enters through the taps its halo exchange drops, emulate_nproc)."""

    def __init__(self, params: Params3DVisco | None = None, output_dir=None, verbose=False, device=-1):
        params = params or Params3DVisco()
        s = setup_3d_visco(params)
        super().__init__(params, s, make_solver_3d_visco(params, s, device=device), output_dir, verbose)
        self.ksource = params.KSOURCE or params.NZ // 2

    def _total_energy(self):
        return self.solver.get_energy()[0]

    def _snapshot_fields(self):          # vx, vy(1:NX,1:NY,NZ_LOCAL) of the cut-plane rank, :1492-1495
        return self.solver.get_plane(0, self.ksource), self.solver.get_plane(1, self.ksource)

    def results(self):
        sx, sy = self.solver.get_seismograms()
        sz = self.solver.get_seismograms_vz()          # extension, quirk B7
        total, ek, ep = self.solver.get_energy()
        if self.output_dir is not None:
            os.makedirs(self.output_dir, exist_ok=True)
            _lib.load().cpml_host_write_seismograms_visco(self.output_dir.encode(), _lib._d(sx), _lib._d(sy), None,
                                                          self.p.NSTEP, self.p.NREC, self.p.DELTAT, self.p.t0)   # :1596-1616
            _lib.load().cpml_host_write_seismograms_vz(self.output_dir.encode(), _lib._d(sz), self.p.NSTEP, self.p.NREC,
                                                       self.p.DELTAT, self.p.t0)
            with open(os.path.join(self.output_dir, "energy.dat"), "w") as f:      # :1478-1483, four columns
                for it in range(self.p.NSTEP):
                    f.write(f" {np.float32(it * self.p.DELTAT)} {np.float32(ek[it])} {np.float32(ep[it])} {np.float32(total[it])}\n")
        return dict(sisvx=sx, sisvy=sy, sisvz=sz, total_energy=total, energy_kinetic=ek, energy_potential=ep,
                    display_log=self.display_log)


class Program2DIso(_ProgramBase):
    """seismic_CPML_2D_isotropic_{second,fourth}_order.f90."""

    def __init__(self, params: Params2DIso | None = None, material=None, output_dir=None,
                 verbose=False, device=-1):
        params = params or Params2DIso()
        s = setup_2d(params, material)
        super().__init__(params, s, make_solver_2d(params, s, device=device), output_dir, verbose)

    def _total_energy(self):
        return self.solver.get_energy()[0]

    def _snapshot_fields(self):          # 2D-2nd :728-731
        return self.solver.get_plane(0), self.solver.get_plane(1)

    def results(self):
        sx, sy = self.solver.get_seismograms()
        total, ek, ep = self.solver.get_energy()
        if self.output_dir is not None:  # 2D-2nd :737-746
            os.makedirs(self.output_dir, exist_ok=True)
            self.write_seismograms()
            _lib.load().cpml_host_write_gnuplot_scripts(self.output_dir.encode(), 1)       # 2D-2nd :748-806
            _lib.load().cpml_host_write_energy_2d(os.path.join(self.output_dir, "energy.dat").encode(),
                                                  _lib._d(ek), _lib._d(ep), self.p.NSTEP, self.p.DELTAT)
        return dict(sisvx=sx, sisvy=sy, energy_kinetic=ek, energy_potential=ep,
                    display_log=self.display_log)


class Program2DVisco(_ProgramBase):
    """seismic_CPML_2D_velocity_and_stress_{second,fourth}_order_viscoelastic.f90."""

    def __init__(self, params: Params2DVisco | None = None, material=None, output_dir=None, verbose=False, device=-1):
        params = params or Params2DVisco()
        s = setup_2d_visco(params, material)
        super().__init__(params, s, make_solver_2d_visco(params, s, device=device), output_dir, verbose)

    def _total_energy(self):
        return self.solver.get_energy()[0]

    def _snapshot_fields(self):          # :1083-1086
        return self.solver.get_plane(0), self.solver.get_plane(1)

    def results(self):
        sx, sy = self.solver.get_seismograms()
        sp = self.solver.get_pressure_seismograms()
        total, ek, ep = self.solver.get_energy()
        if self.output_dir is not None:
            # write_seismograms of 2D-visco-4th :1145-1193: the files the reference's
            # plotall_fit_is_perfect_for_viscoelastic_fourth_order.gnu reads
            os.makedirs(self.output_dir, exist_ok=True)
            _lib.load().cpml_host_write_seismograms_visco(self.output_dir.encode(), _lib._d(sx), _lib._d(sy), _lib._d(sp),
                                                          self.p.NSTEP, self.p.NREC, self.p.DELTAT, self.p.t0)
        return dict(sisvx=sx, sisvy=sy, sispressure=sp, energy_kinetic=ek, energy_potential=ep,
                    display_log=self.display_log)
