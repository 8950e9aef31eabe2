"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  Parity is pinned by an execution of the reference source (oracle/f90_exec.py,
tests/test_reference_vectors.py) -- see oracle/cpml_oracle.h.

Two builds of the same C file (oracle/Makefile):
  golden -- gcc -O2 -ffp-contract=off, serial: the parity checker;
  golden_omp -- the same strict arithmetic with the OpenMP loops on: identical fields and seismograms, energy sums in
            another order (used by the one large reference-vector test);
  timed  -- gcc -O3 -march=x86-64-v3 -fopenmp: the CPU baseline that is timed.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS: dict[str, C.CDLL] = {}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class Oracle2DConfig(C.Structure):
    _fields_ = [("order", C.c_int), ("nx", C.c_int), ("ny", C.c_int),
                ("deltax", C.c_double), ("deltay", C.c_double), ("deltat", C.c_double),
                ("nstep", C.c_int), ("npoints_pml", C.c_int),
                ("isource", C.c_int), ("jsource", C.c_int), ("nrec", C.c_int)]


class Oracle3DConfig(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("nproc", C.c_int),
                ("deltax", C.c_double), ("deltay", C.c_double), ("deltaz", C.c_double),
                ("deltat", C.c_double),
                ("lambda_", C.c_double), ("mu", C.c_double), ("lambdaplustwomu", C.c_double),
                ("rho", C.c_double),
                ("nstep", C.c_int), ("npoints_pml", C.c_int),
                ("isource", C.c_int), ("jsource", C.c_int), ("nrec", C.c_int),
                ("energy_bug_compat", C.c_int)]


class OracleV3DConfig(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("nproc", C.c_int),
                ("deltax", C.c_double), ("deltay", C.c_double), ("deltaz", C.c_double),
                ("deltat", C.c_double),
                ("lambda_", C.c_double), ("mu", C.c_double), ("rho", C.c_double),
                ("nstep", C.c_int), ("npoints_pml", C.c_int),
                ("isource", C.c_int), ("jsource", C.c_int), ("nrec", C.c_int),
                ("tau_epsilon_nu1", C.c_double * 2), ("tau_sigma_nu1", C.c_double * 2),
                ("tau_epsilon_nu2", C.c_double * 2), ("tau_sigma_nu2", C.c_double * 2),
                ("complete_halos", C.c_int), ("sigmazz_isotropic", C.c_int)]


class OracleV2DConfig(C.Structure):
    _fields_ = [("order", C.c_int), ("nx", C.c_int), ("ny", C.c_int),
                ("deltax", C.c_double), ("deltay", C.c_double), ("deltat", C.c_double),
                ("nstep", C.c_int), ("npoints_pml", C.c_int),
                ("isource", C.c_int), ("jsource", C.c_int), ("nrec", C.c_int),
                ("viscoelastic_attenuation", C.c_int), ("compute_energy", C.c_int),
                ("tau_epsilon_nu1", C.c_double * 3), ("tau_sigma_nu1", C.c_double * 3),
                ("tau_epsilon_nu2", C.c_double * 3), ("tau_sigma_nu2", C.c_double * 3)]


def build(force: bool = False) -> None:
    """Compile both oracle libraries (no-op when they are already there)."""
    names = ["liboracle_golden.so", "liboracle_golden_omp.so", "liboracle_timed.so"]
    srcs = ["cpml_oracle.c", "cpml_oracle_visco.c", "cpml_oracle_visco2d.c", "cpml_oracle.h", "oracle_internal.h", "Makefile"]
    if not force and all(os.path.exists(os.path.join(_HERE, n)) for n in names):
        # prebuilt libraries travel to the GPU box; rebuild only when a source is newer
        newest = max(os.path.getmtime(os.path.join(_HERE, f)) for f in srcs)
        if all(os.path.getmtime(os.path.join(_HERE, n)) >= newest for n in names):
            return
    subprocess.run(["make", "-C", _HERE, "-B" if force else "-s", "all"], check=True,
                   stdout=subprocess.DEVNULL)


def lib(kind: str = "golden") -> C.CDLL:
    if kind not in _LIBS:
        path = os.path.join(_HERE, f"liboracle_{kind}.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.oracle_pml_profile.restype = None
        L.oracle_pml_profile.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int,
                                         C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                         C.c_int, C.c_int] + [_dp] * 6
        L.oracle_source_series.restype = None
        L.oracle_source_series.argtypes = [C.c_int] + [C.c_double] * 5 + [_dp, _dp]
        L.oracle_find_receivers.restype = None
        L.oracle_find_receivers.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_int,
                                            C.c_double, C.c_double, C.c_double, C.c_double,
                                            _ip, _ip, _dp]
        L.oracle_run_2d.restype = C.c_int
        L.oracle_run_2d.argtypes = [C.POINTER(Oracle2DConfig)] + [_dp] * 3 + [_dp] * 12 + [_dp] * 2 \
            + [_ip] * 2 + [_dp] * 4 + [_dp] * 5 + [_dp]
        L.oracle_run_3d_iso.restype = C.c_int
        L.oracle_run_3d_iso.argtypes = [C.POINTER(Oracle3DConfig)] + [_dp] * 18 + [_dp] * 2 + [_ip] * 2 \
            + [_dp] * 3 + [_dp] * 2 + [_dp] * 2
        L.oracle_pml_profile_visco.restype = None
        L.oracle_pml_profile_visco.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int,
                                               C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                               C.c_double, C.c_int, C.c_int] + [_dp] * 6
        L.oracle_find_receivers_visco.restype = None
        L.oracle_find_receivers_visco.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_int,
                                                  _dp, _dp, _ip, _ip, _dp]
        L.oracle_run_3d_visco.restype = C.c_int
        L.oracle_run_3d_visco.argtypes = [C.POINTER(OracleV3DConfig)] + [_dp] * 18 + [_dp] * 2 + [_ip] * 2 \
            + [_dp] * 5 + [_dp] * 2
        L.oracle_source_series_ricker.restype = None
        L.oracle_source_series_ricker.argtypes = [C.c_int] + [C.c_double] * 7 + [_dp, _dp]
        L.oracle_run_2d_visco.restype = C.c_int
        L.oracle_run_2d_visco.argtypes = [C.POINTER(OracleV2DConfig)] + [_dp] * 3 + [_dp] * 12 + [_dp] * 2 \
            + [_ip] * 2 + [_dp] * 5 + [_dp] * 3
        L.oracle_set_num_threads.restype = None
        L.oracle_set_num_threads.argtypes = [C.c_int]
        L.oracle_num_threads.restype = C.c_int
        L.oracle_set_warmup_steps.argtypes = [C.c_int]
        L.oracle_last_loop_seconds.restype = C.c_double
        L.oracle_set_ftz.argtypes = [C.c_int]
        _LIBS[kind] = L
    return _LIBS[kind]


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def pml_profile(n, delta, deltat, npoints_pml, use_min=True, use_max=True, *, cp, rcoef=0.001,
                npower=2.0, k_max_pml=1.0, alpha_max_pml, origin_top_uses_n=False,
                clamp_alpha=False, kind="golden"):
    """Returns dict(a, b, K, a_half, b_half, K_half) of length-n arrays."""
    out = {k: np.zeros(n) for k in ("a", "b", "K", "a_half", "b_half", "K_half")}
    lib(kind).oracle_pml_profile(n, delta, deltat, npoints_pml, int(use_min), int(use_max),
                                 cp, rcoef, npower, k_max_pml, alpha_max_pml,
                                 int(origin_top_uses_n), int(clamp_alpha),
                                 *[_d(out[k]) for k in ("a", "b", "K", "a_half", "b_half", "K_half")])
    return out


def source_series(nstep, deltat, f0, t0, factor, angle_force_deg, kind="golden"):
    fx, fy = np.zeros(nstep), np.zeros(nstep)
    lib(kind).oracle_source_series(nstep, deltat, f0, t0, factor, angle_force_deg, _d(fx), _d(fy))
    return fx, fy


def find_receivers(nx, ny, deltax, deltay, nrec, xdeb, ydeb, xfin, yfin, kind="golden"):
    ix, iy = np.zeros(nrec, dtype=np.int32), np.zeros(nrec, dtype=np.int32)
    dist = np.zeros(nrec)
    lib(kind).oracle_find_receivers(nx, ny, deltax, deltay, nrec, xdeb, ydeb, xfin, yfin,
                                    _i(ix), _i(iy), _d(dist))
    return ix, iy, dist


_PK = ("a", "b", "K", "a_half", "b_half", "K_half")


def run_2d(*, order, nx, ny, deltax, deltay, deltat, nstep, npoints_pml, isource, jsource,
           lam, mu, rho, prof_x, prof_y, force_x, force_y, ix_rec, iy_rec,
           want_fields=False, kind="golden"):
    """lam/mu/rho: arrays of nx*ny values, i fastest (shape (ny, nx) in C order)."""
    nrec = len(ix_rec)
    cfg = Oracle2DConfig(order, nx, ny, deltax, deltay, deltat, nstep, npoints_pml,
                         isource, jsource, nrec)
    lam, mu, rho = _f64(lam).ravel(), _f64(mu).ravel(), _f64(rho).ravel()
    assert lam.size == nx * ny and mu.size == nx * ny and rho.size == nx * ny
    px = [_f64(prof_x[k]) for k in _PK]
    py = [_f64(prof_y[k]) for k in _PK]
    fx, fy = _f64(force_x), _f64(force_y)
    assert fx.size >= nstep and fy.size >= nstep
    ixr = np.ascontiguousarray(ix_rec, dtype=np.int32)
    iyr = np.ascontiguousarray(iy_rec, dtype=np.int32)
    sisvx = np.zeros((nrec, nstep))
    sisvy = np.zeros((nrec, nstep))
    ek, ep = np.zeros(nstep), np.zeros(nstep)
    fields = [np.zeros((ny, nx)) if want_fields else None for _ in range(5)]
    vnorm = C.c_double(0.0)
    rc = lib(kind).oracle_run_2d(C.byref(cfg), _d(lam), _d(mu), _d(rho),
                                 *[_d(p) for p in px], *[_d(p) for p in py],
                                 _d(fx), _d(fy), _i(ixr), _i(iyr),
                                 _d(sisvx), _d(sisvy), _d(ek), _d(ep),
                                 *[_d(f) for f in fields], C.byref(vnorm))
    if rc != 0:
        raise RuntimeError(f"oracle_run_2d failed rc={rc}")
    out = dict(sisvx=sisvx, sisvy=sisvy, energy_kinetic=ek, energy_potential=ep,
               velocnorm=vnorm.value)
    if want_fields:
        out.update(dict(zip(("vx", "vy", "sigmaxx", "sigmayy", "sigmaxy"), fields)))
    return out


def run_3d_iso(*, nx, ny, nz, nproc, deltax, deltay, deltaz, deltat, lam, mu, lambdaplustwomu, rho,
               nstep, npoints_pml, isource, jsource, prof_x, prof_y, prof_z, force_x, force_y,
               ix_rec, iy_rec, energy_bug_compat=True, want_planes=False, want_fields=False,
               kind="golden"):
    nrec = len(ix_rec)
    cfg = Oracle3DConfig(nx, ny, nz, nproc, deltax, deltay, deltaz, deltat, lam, mu,
                         lambdaplustwomu, rho, nstep, npoints_pml, isource, jsource, nrec,
                         int(energy_bug_compat))
    px = [_f64(prof_x[k]) for k in _PK]
    py = [_f64(prof_y[k]) for k in _PK]
    pz = [_f64(prof_z[k]) for k in _PK]
    assert px[0].size == nx and py[0].size == ny and pz[0].size == nz
    fx, fy = _f64(force_x), _f64(force_y)
    assert fx.size >= nstep and fy.size >= nstep
    ixr = np.ascontiguousarray(ix_rec, dtype=np.int32)
    iyr = np.ascontiguousarray(iy_rec, dtype=np.int32)
    sisvx = np.zeros((nrec, nstep))
    sisvy = np.zeros((nrec, nstep))
    energy = np.zeros(nstep)
    pvx = np.zeros((ny, nx)) if want_planes else None
    pvy = np.zeros((ny, nx)) if want_planes else None
    fields = np.zeros((9, nz, ny, nx)) if want_fields else None
    vnorm = C.c_double(0.0)
    rc = lib(kind).oracle_run_3d_iso(C.byref(cfg), *[_d(p) for p in px], *[_d(p) for p in py],
                                     *[_d(p) for p in pz], _d(fx), _d(fy), _i(ixr), _i(iyr),
                                     _d(sisvx), _d(sisvy), _d(energy), _d(pvx), _d(pvy),
                                     _d(fields), C.byref(vnorm))
    if rc != 0:
        raise RuntimeError(f"oracle_run_3d_iso failed rc={rc}")
    out = dict(sisvx=sisvx, sisvy=sisvy, total_energy=energy, vnorm=vnorm.value)
    if want_planes:
        out.update(plane_vx=pvx, plane_vy=pvy)
    if want_fields:
        out.update(dict(zip(("vx", "vy", "vz", "sigmaxx", "sigmayy", "sigmazz", "sigmaxy",
                             "sigmaxz", "sigmayz"), fields)))
    return out


def pml_profile_visco(n, delta, deltat, npoints_pml, use_min=True, use_max=True, *, cp, sqrt_taumax,
                      rcoef=0.0001, npower=2.0, k_max_pml=7.0, alpha_max_pml, clamp_alpha=False,
                      kind="golden"):
    """3D-visco :533-801 (d0 scaled by dsqrt(taumax), Rcoef = 1e-4, K_MAX_PML = 7)."""
    out = {k: np.zeros(n) for k in _PK}
    lib(kind).oracle_pml_profile_visco(n, delta, deltat, npoints_pml, int(use_min), int(use_max),
                                       cp, sqrt_taumax, rcoef, npower, k_max_pml, alpha_max_pml,
                                       0, int(clamp_alpha), *[_d(out[k]) for k in _PK])
    return out


def find_receivers_visco(nx, ny, deltax, deltay, xrec, yrec, kind="golden"):
    xr, yr = _f64(xrec), _f64(yrec)
    nrec = xr.size
    ix, iy = np.zeros(nrec, dtype=np.int32), np.zeros(nrec, dtype=np.int32)
    dist = np.zeros(nrec)
    lib(kind).oracle_find_receivers_visco(nx, ny, deltax, deltay, nrec, _d(xr), _d(yr), _i(ix), _i(iy), _d(dist))
    return ix, iy, dist


VISCO_FIELDS = ("vx", "vy", "vz", "sigmaxx", "sigmayy", "sigmazz", "sigmaxy", "sigmaxz", "sigmayz",
                "sigmaxx_R", "sigmayy_R", "sigmazz_R", "sigmaxy_R", "sigmaxz_R", "sigmayz_R")


def run_3d_visco(*, nx, ny, nz, nproc, deltax, deltay, deltaz, deltat, lam, mu, rho, nstep, npoints_pml,
                 isource, jsource, tau_epsilon_nu1, tau_sigma_nu1, tau_epsilon_nu2, tau_sigma_nu2,
                 prof_x, prof_y, prof_z, force_x, force_y, ix_rec, iy_rec, complete_halos=False, sigmazz_isotropic=False,
                 want_fields=False, kind="golden", **_ignored):
    nrec = len(ix_rec)
    A2 = C.c_double * 2
    cfg = OracleV3DConfig(nx, ny, nz, nproc, deltax, deltay, deltaz, deltat, lam, mu, rho, nstep,
                          npoints_pml, isource, jsource, nrec, A2(*tau_epsilon_nu1), A2(*tau_sigma_nu1),
                          A2(*tau_epsilon_nu2), A2(*tau_sigma_nu2), int(complete_halos), int(sigmazz_isotropic))
    px = [_f64(prof_x[k]) for k in _PK]
    py = [_f64(prof_y[k]) for k in _PK]
    pz = [_f64(prof_z[k]) for k in _PK]
    assert px[0].size == nx and py[0].size == ny and pz[0].size == nz
    fx, fy = _f64(force_x), _f64(force_y)
    assert fx.size >= nstep and fy.size >= nstep
    ixr = np.ascontiguousarray(ix_rec, dtype=np.int32)
    iyr = np.ascontiguousarray(iy_rec, dtype=np.int32)
    sisvx = np.zeros((nrec, nstep))
    sisvy = np.zeros((nrec, nstep))
    et, ek, ep = np.zeros(nstep), np.zeros(nstep), np.zeros(nstep)
    fields = np.zeros((15, nz, ny, nx)) if want_fields else None
    vnorm = C.c_double(0.0)
    rc = lib(kind).oracle_run_3d_visco(C.byref(cfg), *[_d(p) for p in px], *[_d(p) for p in py],
                                       *[_d(p) for p in pz], _d(fx), _d(fy), _i(ixr), _i(iyr),
                                       _d(sisvx), _d(sisvy), _d(et), _d(ek), _d(ep), _d(fields),
                                       C.byref(vnorm))
    if rc != 0:
        raise RuntimeError(f"oracle_run_3d_visco failed rc={rc}")
    out = dict(sisvx=sisvx, sisvy=sisvy, total_energy=et, energy_kinetic=ek, energy_potential=ep,
               vnorm=vnorm.value)
    if want_fields:
        out.update(dict(zip(VISCO_FIELDS, fields)))
    return out


def source_series_ricker(nstep, deltat, f0, t0, factor, angle_force_deg, deltax, deltay, kind="golden"):
    fx, fy = np.zeros(nstep), np.zeros(nstep)
    lib(kind).oracle_source_series_ricker(nstep, deltat, f0, t0, factor, angle_force_deg, deltax, deltay, _d(fx), _d(fy))
    return fx, fy


def run_2d_visco(*, order, nx, ny, deltax, deltay, deltat, nstep, npoints_pml, isource, jsource, lam, mu, rho,
                 tau_epsilon_nu1, tau_sigma_nu1, tau_epsilon_nu2, tau_sigma_nu2, prof_x, prof_y, force_x, force_y,
                 ix_rec, iy_rec, viscoelastic_attenuation=True, compute_energy=False, want_fields=False,
                 kind="golden", **_ignored):
    """lam/mu: UNRELAXED Lame parameters, nx*ny values, i fastest."""
    nrec = len(ix_rec)
    A3 = C.c_double * 3
    cfg = OracleV2DConfig(order, nx, ny, deltax, deltay, deltat, nstep, npoints_pml, isource, jsource, nrec,
                          int(viscoelastic_attenuation), int(compute_energy), A3(*tau_epsilon_nu1),
                          A3(*tau_sigma_nu1), A3(*tau_epsilon_nu2), A3(*tau_sigma_nu2))
    lam, mu, rho = _f64(lam).ravel(), _f64(mu).ravel(), _f64(rho).ravel()
    assert lam.size == nx * ny and mu.size == nx * ny and rho.size == nx * ny
    px = [_f64(prof_x[k]) for k in _PK]
    py = [_f64(prof_y[k]) for k in _PK]
    fx, fy = _f64(force_x), _f64(force_y)
    assert fx.size >= nstep and fy.size >= nstep
    ixr = np.ascontiguousarray(ix_rec, dtype=np.int32)
    iyr = np.ascontiguousarray(iy_rec, dtype=np.int32)
    sisvx, sisvy, sisp = np.zeros((nrec, nstep)), np.zeros((nrec, nstep)), np.zeros((nrec, nstep))
    ek, ep = np.zeros(nstep), np.zeros(nstep)
    fields = np.zeros((5, ny, nx)) if want_fields else None
    memvar = np.zeros((3, 3, ny, nx)) if want_fields else None
    vnorm = C.c_double(0.0)
    rc = lib(kind).oracle_run_2d_visco(C.byref(cfg), _d(lam), _d(mu), _d(rho), *[_d(p) for p in px],
                                       *[_d(p) for p in py], _d(fx), _d(fy), _i(ixr), _i(iyr),
                                       _d(sisvx), _d(sisvy), _d(sisp), _d(ek), _d(ep), _d(fields), _d(memvar),
                                       C.byref(vnorm))
    if rc != 0:
        raise RuntimeError(f"oracle_run_2d_visco failed rc={rc}")
    out = dict(sisvx=sisvx, sisvy=sisvy, sispressure=sisp, energy_kinetic=ek, energy_potential=ep,
               velocnorm=vnorm.value)
    if want_fields:
        out.update(dict(zip(("vx", "vy", "sigmaxx", "sigmayy", "sigmaxy"), fields)))
        out.update(e1=memvar[0], e11=memvar[1], e13=memvar[2])
    return out


def num_threads(kind="timed") -> int:
    return lib(kind).oracle_num_threads()


def set_num_threads(n: int, kind="timed") -> None:
    lib(kind).oracle_set_num_threads(int(n))


def set_ftz(on: bool, kind="timed") -> None:
    lib(kind).oracle_set_ftz(int(on))


def set_warmup_steps(w: int, kind="timed") -> None:
    """The next run_* call excludes its first w steps from last_loop_seconds()."""
    lib(kind).oracle_set_warmup_steps(int(w))


def last_loop_seconds(kind="timed") -> float:
    return lib(kind).oracle_last_loop_seconds()
