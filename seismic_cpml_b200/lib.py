"""ctypes binding of libcpml_b200.so (the C ABI of include/cpml_b200.h).

This is the Python stand-in for the Fortran ISO_C_BINDING module a maintainer of the
reference would add (drivers/fortran/cpml_b200_mod.f90): same entry points, same
argument meaning, same error codes.  There is no CPU fallback: if the CUDA library is
missing or no device is usable, calls raise CpmlError.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcpml_b200.so")

CPML_OK, CPML_EINVAL, CPML_ETOPOLOGY, CPML_ECFL, CPML_ECUDA, CPML_ESTATE, CPML_ENOMEM = range(7)
ERROR_NAMES = {0: "CPML_OK", 1: "CPML_EINVAL", 2: "CPML_ETOPOLOGY", 3: "CPML_ECFL", 4: "CPML_ECUDA",
               5: "CPML_ESTATE", 6: "CPML_ENOMEM"}
AXIS_X, AXIS_Y, AXIS_Z = 0, 1, 2
FIELDS_3D = ("vx", "vy", "vz", "sigmaxx", "sigmayy", "sigmazz", "sigmaxy", "sigmaxz", "sigmayz")
FIELDS_3D_VISCO = FIELDS_3D + ("sigmaxx_R", "sigmayy_R", "sigmazz_R", "sigmaxy_R", "sigmaxz_R", "sigmayz_R")
FIELDS_2D = ("vx", "vy", "sigmaxx", "sigmayy", "sigmaxy")
FIELDS_2D_VISCO = FIELDS_2D + tuple(f"{e}_{l}" for e in ("e1", "e11", "e13") for l in (1, 2, 3))
PROFILE_KEYS = ("a", "b", "K", "a_half", "b_half", "K_half")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class CpmlError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"{ERROR_NAMES.get(code, code)}: {message}")
        self.code = code


class CpmlConfig(C.Structure):
    """struct cpml_config of include/cpml_b200.h (field for field)."""
    _fields_ = [("ndim", C.c_int32), ("order", C.c_int32),
                ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32),
                ("nstep", C.c_int32), ("npoints_pml", C.c_int32), ("nrec", C.c_int32),
                ("isource", C.c_int32), ("jsource", C.c_int32), ("ksource", C.c_int32),
                ("nslabs", C.c_int32), ("slab_rank", C.c_int32), ("device", C.c_int32),
                ("energy_bug_compat", C.c_int32), ("rheology", C.c_int32),
                ("emulate_nproc", C.c_int32), ("compute_energy", C.c_int32), ("sigmazz_isotropic", C.c_int32),
                ("precision", C.c_int32),
                ("deltax", C.c_double), ("deltay", C.c_double), ("deltaz", C.c_double),
                ("deltat", C.c_double),
                ("lambda_", C.c_double), ("mu", C.c_double), ("lambdaplustwomu", C.c_double),
                ("rho", C.c_double), ("cp", C.c_double), ("reserved_d", C.c_double * 4)]


# every symbol include/cpml_b200.h declares: name -> (restype, argtypes)
_H = C.c_void_p
SYMBOLS = {
    "cpml_abi_version": (C.c_int32, []),
    "cpml_create": (C.c_int32, [C.POINTER(CpmlConfig), C.POINTER(_H)]),
    "cpml_destroy": (C.c_int32, [_H]),
    "cpml_last_error": (C.c_char_p, [_H]),
    "cpml_reset": (C.c_int32, [_H]),
    "cpml_set_stream": (C.c_int32, [_H, C.c_void_p]),
    "cpml_set_profiles": (C.c_int32, [_H, C.c_int32] + [_dp] * 6 + [C.c_int32]),
    "cpml_set_material_2d": (C.c_int32, [_H, _dp, _dp, _dp]),
    "cpml_set_attenuation": (C.c_int32, [_H, C.c_int32, _dp, _dp, _dp, _dp]),
    "cpml_set_source_series": (C.c_int32, [_H, _dp, _dp, C.c_int32]),
    "cpml_set_source_step": (C.c_int32, [_H, C.c_int32, C.c_double, C.c_double]),
    "cpml_fetch_step": (C.c_int32, [_H, C.c_int32]),
    "cpml_get_fetched_step": (C.c_int32, [_H, C.c_int32, _dp]),
    "cpml_set_receivers": (C.c_int32, [_H, _ip, _ip, C.c_int32]),
    "cpml_run": (C.c_int32, [_H, C.c_int32, C.c_int32]),
    "cpml_step_stress": (C.c_int32, [_H, C.c_int32]),
    "cpml_step_velocity": (C.c_int32, [_H, C.c_int32]),
    "cpml_step_finish": (C.c_int32, [_H, C.c_int32]),
    "cpml_synchronize": (C.c_int32, [_H]),
    "cpml_halo_plane": (C.c_int32, [_H, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "cpml_copy_plane": (C.c_int32, [_H, C.c_int32, _H, C.c_int32, C.c_int32]),
    "cpml_p2p_export": (C.c_int32, [_H, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
    "cpml_p2p_attach_ipc": (C.c_int32, [_H, C.c_int32, C.c_void_p, C.c_int64]),
    "cpml_p2p_attach_local": (C.c_int32, [_H, C.c_int32, _H]),
    "cpml_p2p_detach": (C.c_int32, [_H]),
    "cpml_get_launch_info": (C.c_int32, [_H, _ip, C.c_int32]),
    "cpml_get_seismograms": (C.c_int32, [_H, _dp, _dp]),
    "cpml_get_pressure_seismograms": (C.c_int32, [_H, _dp]),
    "cpml_get_seismograms_vz": (C.c_int32, [_H, _dp]),
    "cpml_get_energy": (C.c_int32, [_H, _dp, _dp, _dp]),
    "cpml_get_plane": (C.c_int32, [_H, C.c_int32, C.c_int32, _dp]),
    "cpml_snapshot_begin": (C.c_int32, [_H, C.c_int32, C.c_int32, C.c_int32]),
    "cpml_snapshot_end": (C.c_int32, [_H, C.c_int32, _dp, C.POINTER(_dp)]),
    "cpml_get_field": (C.c_int32, [_H, C.c_int32, _dp]),
    "cpml_get_maxnorm": (C.c_int32, [_H, _dp]),
    "cpml_get_kernel_times": (C.c_int32, [_H, _dp, _dp, C.POINTER(C.c_int64), C.c_int32]),
    "cpml_enable_kernel_timing": (C.c_int32, [_H, C.c_int32]),
    "cpml_algorithmic_bytes": (C.c_int32, [_H, _dp, _dp]),
    "cpml_multi_create": (C.c_int32, [C.POINTER(CpmlConfig), C.c_int32, _ip, C.POINTER(_H)]),
    "cpml_multi_destroy": (C.c_int32, [_H]),
    "cpml_multi_last_error": (C.c_char_p, [_H]),
    "cpml_multi_ngpus": (C.c_int32, [_H]),
    "cpml_multi_slab": (C.c_int32, [_H, C.c_int32, C.POINTER(_H)]),
    "cpml_multi_reset": (C.c_int32, [_H]),
    "cpml_multi_set_profiles": (C.c_int32, [_H, C.c_int32] + [_dp] * 6 + [C.c_int32]),
    "cpml_multi_set_attenuation": (C.c_int32, [_H, C.c_int32, _dp, _dp, _dp, _dp]),
    "cpml_multi_set_source_series": (C.c_int32, [_H, _dp, _dp, C.c_int32]),
    "cpml_multi_set_receivers": (C.c_int32, [_H, _ip, _ip, C.c_int32]),
    "cpml_multi_step": (C.c_int32, [_H, C.c_int32]),
    "cpml_multi_run": (C.c_int32, [_H, C.c_int32, C.c_int32]),
    "cpml_multi_synchronize": (C.c_int32, [_H]),
    "cpml_multi_get_seismograms": (C.c_int32, [_H, _dp, _dp]),
    "cpml_multi_get_seismograms_vz": (C.c_int32, [_H, _dp]),
    "cpml_multi_get_energy": (C.c_int32, [_H, _dp, _dp, _dp]),
    "cpml_multi_get_plane": (C.c_int32, [_H, C.c_int32, C.c_int32, _dp]),
    "cpml_multi_get_field": (C.c_int32, [_H, C.c_int32, _dp]),
    "cpml_multi_get_maxnorm": (C.c_int32, [_H, _dp]),
    "cpml_host_pml_profile": (C.c_int32, [C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_int32, C.c_int32,
                                          C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                          C.c_int32, C.c_int32] + [_dp] * 6),
    "cpml_host_pml_profile_visco": (C.c_int32, [C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_int32, C.c_int32,
                                                C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                                C.c_double, C.c_int32, C.c_int32] + [_dp] * 6),
    "cpml_host_find_receivers_at": (C.c_int32, [C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_int32,
                                                _dp, _dp, C.c_int32, _ip, _ip, _dp]),
    "cpml_host_source_series": (C.c_int32, [C.c_int32] + [C.c_double] * 5 + [_dp, _dp]),
    "cpml_host_find_receivers": (C.c_int32, [C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_int32,
                                             C.c_double, C.c_double, C.c_double, C.c_double, _ip, _ip, _dp]),
    "cpml_host_courant": (C.c_double, [C.c_double] * 5),
    "cpml_host_attenuation_fit": (C.c_int32, [C.c_int32] + [C.c_double] * 4 + [_dp, _dp, _dp]),
    "cpml_host_attenuation_fit_linear": (C.c_int32, [C.c_int32] + [C.c_double] * 3 + [_dp, _dp]),
    "cpml_host_write_seismograms": (C.c_int32, [C.c_char_p, _dp, _dp, C.c_int32, C.c_int32, C.c_double]),
    "cpml_host_write_seismograms_visco": (C.c_int32, [C.c_char_p, _dp, _dp, _dp, C.c_int32, C.c_int32, C.c_double, C.c_double]),
    "cpml_host_write_seismograms_vz": (C.c_int32, [C.c_char_p, _dp, C.c_int32, C.c_int32, C.c_double, C.c_double]),
    "cpml_host_write_timestamp": (C.c_int32, [C.c_char_p, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_double]),
    "cpml_host_write_energy_3d": (C.c_int32, [C.c_char_p, _dp, C.c_int32, C.c_double]),
    "cpml_host_write_energy_2d": (C.c_int32, [C.c_char_p, _dp, _dp, C.c_int32, C.c_double]),
    "cpml_host_format_real": (C.c_int32, [C.c_double, C.c_int32, C.c_char_p, C.c_int32]),
    "cpml_host_write_gnuplot_scripts": (C.c_int32, [C.c_char_p, C.c_int32]),
    "cpml_host_create_color_image": (C.c_int32, [C.c_char_p, _dp, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                                 C.c_int32, _ip, _ip, C.c_int32, C.c_int32, C.c_int32,
                                                 C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
}

_lib = None


def load() -> C.CDLL:
    """Loads libcpml_b200.so; fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CpmlError(CPML_ECUDA, f"{LIB_PATH} is missing: run `python -m seismic_cpml_b200.build` "
                            "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)   # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        if L.cpml_abi_version() != 1:
            raise CpmlError(CPML_EINVAL, "libcpml_b200.so ABI version mismatch")
        _lib = L
    return _lib


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


# ---------------------------------------------------------------- host helpers

def host_pml_profile(n, delta, deltat, npoints_pml, use_pml_min=True, use_pml_max=True, *, cp,
                     rcoef=0.001, npower=2.0, k_max_pml=1.0, alpha_max_pml,
                     origin_top_uses_n=False, clamp_alpha=False):
    out = {k: np.zeros(n) for k in PROFILE_KEYS}
    rc = load().cpml_host_pml_profile(n, delta, deltat, npoints_pml, int(use_pml_min), int(use_pml_max),
                                      cp, rcoef, npower, k_max_pml, alpha_max_pml,
                                      int(origin_top_uses_n), int(clamp_alpha),
                                      *[_d(out[k]) for k in PROFILE_KEYS])
    if rc:
        raise CpmlError(rc, "cpml_host_pml_profile")
    return out


def host_pml_profile_visco(n, delta, deltat, npoints_pml, use_pml_min=True, use_pml_max=True, *, cp,
                           sqrt_taumax, rcoef=0.0001, npower=2.0, k_max_pml=7.0, alpha_max_pml,
                           clamp_alpha=False):
    out = {k: np.zeros(n) for k in PROFILE_KEYS}
    rc = load().cpml_host_pml_profile_visco(n, delta, deltat, npoints_pml, int(use_pml_min), int(use_pml_max),
                                            cp, sqrt_taumax, rcoef, npower, k_max_pml, alpha_max_pml,
                                            0, int(clamp_alpha), *[_d(out[k]) for k in PROFILE_KEYS])
    if rc:
        raise CpmlError(rc, "cpml_host_pml_profile_visco")
    return out


def host_find_receivers_at(nx, ny, deltax, deltay, xrec, yrec, index_origin=1):
    xr, yr = _f64(xrec), _f64(yrec)
    nrec = xr.size
    ix, iy = np.zeros(nrec, dtype=np.int32), np.zeros(nrec, dtype=np.int32)
    dist = np.zeros(nrec)
    rc = load().cpml_host_find_receivers_at(nx, ny, deltax, deltay, nrec, _d(xr), _d(yr), index_origin,
                                            _i(ix), _i(iy), _d(dist))
    if rc:
        raise CpmlError(rc, "cpml_host_find_receivers_at")
    return ix, iy, dist


def host_source_series(nstep, deltat, f0, t0, factor, angle_force_deg):
    fx, fy = np.zeros(nstep), np.zeros(nstep)
    rc = load().cpml_host_source_series(nstep, deltat, f0, t0, factor, angle_force_deg, _d(fx), _d(fy))
    if rc:
        raise CpmlError(rc, "cpml_host_source_series")
    return fx, fy


def host_find_receivers(nx, ny, deltax, deltay, nrec, xdeb, ydeb, xfin, yfin):
    ix, iy = np.zeros(nrec, dtype=np.int32), np.zeros(nrec, dtype=np.int32)
    dist = np.zeros(nrec)
    rc = load().cpml_host_find_receivers(nx, ny, deltax, deltay, nrec, xdeb, ydeb, xfin, yfin,
                                         _i(ix), _i(iy), _d(dist))
    if rc:
        raise CpmlError(rc, "cpml_host_find_receivers")
    return ix, iy, dist


def host_courant(cp, deltat, deltax, deltay, deltaz=0.0):
    return load().cpml_host_courant(cp, deltat, deltax, deltay, deltaz)


def host_attenuation_band(f0):
    """f_min, f_max of the fit: f_max / f_min = 12 centred (in log) on f0 (2D-visco-4th :366-368,
    3D-visco :434-435)."""
    import math
    f_min = math.exp(math.log(f0) - math.log(12.0) / 2.0)
    return f_min, 12.0 * f_min


def host_attenuation_fit(n_sls, qref, f0, f_min=None, f_max=None, *, linear_only=False, return_info=False):
    """compute_attenuation_coeffs(N, Qref, f0, f_min, f_max, tau_epsilon, tau_sigma) of
    attenuation_model_with_SolvOpt.f90:122-169 -> (tau_epsilon, tau_sigma) tuples."""
    if f_min is None or f_max is None:
        f_min, f_max = host_attenuation_band(f0)
    te = np.zeros(n_sls)
    ts = np.zeros(n_sls)
    info = np.zeros(4)
    if linear_only:
        rc = load().cpml_host_attenuation_fit_linear(n_sls, qref, f_min, f_max, _d(te), _d(ts))
    else:
        rc = load().cpml_host_attenuation_fit(n_sls, qref, f0, f_min, f_max, _d(te), _d(ts), _d(info))
    if rc:
        raise CpmlError(rc, "cpml_host_attenuation_fit")
    out = (tuple(float(v) for v in te), tuple(float(v) for v in ts))
    return out + (info,) if return_info else out


def host_format_real(value, kind=4) -> str:
    """gfortran's list-directed text of a REAL(kind) item (cpml_host_format_real)."""
    buf = C.create_string_buffer(64)
    rc = load().cpml_host_format_real(float(value), kind, buf, 64)
    if rc:
        raise CpmlError(rc, "cpml_host_format_real")
    return buf.value.decode()


# ---------------------------------------------------------------- handle

class Solver:
    """One cpml_handle: a whole 2-D grid, a whole 3-D grid, or one z-slab of a 3-D grid."""

    def __init__(self, *, ndim, order=2, nx, ny, nz=1, nstep, npoints_pml, nrec, isource, jsource,
                 ksource=0, nslabs=1, slab_rank=0, device=-1, energy_bug_compat=True, sigmazz_isotropic=False,
                 deltax, deltay, deltaz=0.0, deltat, lam=0.0, mu=0.0, lambdaplustwomu=0.0, rho=0.0,
                 cp=0.0, rheology=0, emulate_nproc=0, compute_energy=False, precision=0):
        self._L = load()
        self.cfg = CpmlConfig(ndim=ndim, order=order, nx=nx, ny=ny, nz=nz, nstep=nstep,
                              npoints_pml=npoints_pml, nrec=nrec, isource=isource, jsource=jsource,
                              ksource=ksource, nslabs=nslabs, slab_rank=slab_rank, device=device,
                              energy_bug_compat=int(energy_bug_compat), sigmazz_isotropic=int(sigmazz_isotropic),
                              rheology=rheology, precision=precision,
                              emulate_nproc=emulate_nproc, compute_energy=int(compute_energy),
                              deltax=deltax, deltay=deltay, deltaz=deltaz, deltat=deltat,
                              lambda_=lam, mu=mu, lambdaplustwomu=lambdaplustwomu, rho=rho, cp=cp)
        self._h = _H()
        rc = self._L.cpml_create(C.byref(self.cfg), C.byref(self._h))
        if rc:
            msg = self._L.cpml_last_error(None)
            self._h = _H()
            raise CpmlError(rc, msg.decode() if msg else "cpml_create")
        self.nzl = nz // nslabs if ndim == 3 else 1
        self.koff = slab_rank * self.nzl

    # -- plumbing
    def _ck(self, rc):
        if rc:
            msg = self._L.cpml_last_error(self._h)
            raise CpmlError(rc, msg.decode() if msg else "")

    def close(self):
        if getattr(self, "_h", None):
            self._L.cpml_destroy(self._h)
            self._h = _H()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- inputs
    def set_stream(self, cuda_stream: int):
        self._ck(self._L.cpml_set_stream(self._h, C.c_void_p(cuda_stream)))

    def set_profiles(self, axis, prof):
        arrs = [_f64(prof[k]) for k in PROFILE_KEYS]
        self._ck(self._L.cpml_set_profiles(self._h, axis, *[_d(a) for a in arrs], arrs[0].size))

    def set_material_2d(self, lam, mu, rho):
        lam, mu, rho = _f64(lam).ravel(), _f64(mu).ravel(), _f64(rho).ravel()
        n = self.cfg.nx * self.cfg.ny
        if lam.size != n or mu.size != n or rho.size != n:
            raise CpmlError(CPML_EINVAL, "material arrays must hold NX*NY values")
        self._ck(self._L.cpml_set_material_2d(self._h, _d(lam), _d(mu), _d(rho)))

    def set_attenuation(self, tau_epsilon_nu1, tau_sigma_nu1, tau_epsilon_nu2, tau_sigma_nu2):
        """Relaxation times of the N_SLS = 2 mechanisms (3D-visco :439-443)."""
        arrs = [_f64(a) for a in (tau_epsilon_nu1, tau_sigma_nu1, tau_epsilon_nu2, tau_sigma_nu2)]
        if any(a.size != arrs[0].size for a in arrs):
            raise CpmlError(CPML_EINVAL, "relaxation-time arrays differ in length")
        self._ck(self._L.cpml_set_attenuation(self._h, arrs[0].size, *[_d(a) for a in arrs]))

    def set_source_series(self, force_x, force_y):
        fx, fy = _f64(force_x), _f64(force_y)
        if fx.size != fy.size:
            raise CpmlError(CPML_EINVAL, "force_x and force_y differ in length")
        self._ck(self._L.cpml_set_source_series(self._h, _d(fx), _d(fy), fx.size))

    def set_source_step(self, it, force_x, force_y):
        """force_x(it), force_y(it) of :1058-1071 for one step (pinned staging, asynchronous)."""
        self._ck(self._L.cpml_set_source_step(self._h, it, float(force_x), float(force_y)))

    def fetch_step(self, it):
        """Queue the device-to-host copy of step `it`'s energy and receiver-1 sample."""
        self._ck(self._L.cpml_fetch_step(self._h, it))

    def get_fetched_step(self, it):
        """(kinetic, potential, sisvx(it,1), sisvy(it,1)) of a fetched step; valid after synchronize()."""
        out = np.zeros(4)
        self._ck(self._L.cpml_get_fetched_step(self._h, it, _d(out)))
        return out

    def set_receivers(self, ix_rec, iy_rec):
        ix = np.ascontiguousarray(ix_rec, dtype=np.int32)
        iy = np.ascontiguousarray(iy_rec, dtype=np.int32)
        self._ck(self._L.cpml_set_receivers(self._h, _i(ix), _i(iy), ix.size))

    def reset(self):
        self._ck(self._L.cpml_reset(self._h))

    # -- the loop
    def run(self, it_begin, it_end):
        self._ck(self._L.cpml_run(self._h, it_begin, it_end))

    def step_stress(self, it):
        self._ck(self._L.cpml_step_stress(self._h, it))

    def step_velocity(self, it):
        self._ck(self._L.cpml_step_velocity(self._h, it))

    def step_finish(self, it):
        self._ck(self._L.cpml_step_finish(self._h, it))

    def synchronize(self):
        self._ck(self._L.cpml_synchronize(self._h))

    def halo_plane(self, field, klocal):
        ptr, nbytes = C.c_void_p(), C.c_int64()
        self._ck(self._L.cpml_halo_plane(self._h, field, klocal, C.byref(ptr), C.byref(nbytes)))
        return ptr.value, nbytes.value

    def copy_plane_from(self, klocal_dst, src: "Solver", klocal_src, field):
        """One MPI_SENDRECV of the reference: plane klocal_src of `src` -> plane klocal_dst here."""
        self._ck(self._L.cpml_copy_plane(self._h, klocal_dst, src._h, klocal_src, field))

    # -- direct slab-to-slab stores (replaces the MPI_SENDRECV calls of :811-823 / :951-963)
    def p2p_export(self) -> bytes:
        """The 64-byte CUDA IPC blob a neighbour process hands to p2p_attach_ipc."""
        buf = C.create_string_buffer(64)
        n = C.c_int64()
        self._ck(self._L.cpml_p2p_export(self._h, buf, 64, C.byref(n)))
        return buf.raw[:n.value]

    def p2p_attach_ipc(self, side: int, blob: bytes):
        buf = C.create_string_buffer(bytes(blob), len(blob))
        self._ck(self._L.cpml_p2p_attach_ipc(self._h, side, buf, len(blob)))

    def p2p_attach_local(self, side: int, neighbour: "Solver"):
        self._ck(self._L.cpml_p2p_attach_local(self._h, side, neighbour._h))

    def p2p_detach(self):
        self._ck(self._L.cpml_p2p_detach(self._h))

    def launch_info(self) -> dict:
        v = np.zeros(14, dtype=np.int32)
        self._ck(self._L.cpml_get_launch_info(self._h, _i(v), 14))
        keys = ("tma", "tile_x", "tile_y", "stages", "planes_per_item", "z_chunks", "items",
                "ctas_stress", "ctas_velocity", "peer_sides", "stress_tile_y", "stress_planes_per_item",
                "stress_z_chunks", "stress_items")
        return dict(zip(keys, (int(x) for x in v)))

    def kernel_names(self):
        """Names of the two update kernels of a 3-D isotropic handle (launch_info()["tma"]: 0 register-marching,
        1 TMA-staged, 2 TMA-staged with a producer warp)."""
        k = self.launch_info()["tma"]
        return {0: ("k_stress3d", "k_velocity3d"), 1: ("k_stress3d_tma", "k_velocity3d_tma"),
                2: ("k_stress3d_ws", "k_velocity3d_ws")}[k]

    # -- outputs
    def get_seismograms(self):
        """(sisvx, sisvy), each shaped (NREC, NSTEP): row r is the trace of receiver r+1
        (the reference's sisvx(:, irec))."""
        nrec, nstep = self.cfg.nrec, self.cfg.nstep
        sx, sy = np.zeros((nrec, nstep)), np.zeros((nrec, nstep))
        self._ck(self._L.cpml_get_seismograms(self._h, _d(sx), _d(sy)))
        return sx, sy

    def get_seismograms_vz(self):
        """sisvz (NREC, NSTEP): vz(ix_rec, iy_rec, NZ/2) -- an extension, the reference records Vx and Vy only
        (quirk B7).  3-D only."""
        sz = np.zeros((self.cfg.nrec, self.cfg.nstep))
        self._ck(self._L.cpml_get_seismograms_vz(self._h, _d(sz)))
        return sz

    def get_pressure_seismograms(self):
        """sispressure shaped (NREC, NSTEP) (2-D viscoelastic programs, 2D-visco-4th :1004-1035)."""
        sp = np.zeros((self.cfg.nrec, self.cfg.nstep))
        self._ck(self._L.cpml_get_pressure_seismograms(self._h, _d(sp)))
        return sp

    def get_energy(self):
        n = self.cfg.nstep
        tot, ek, ep = np.zeros(n), np.zeros(n), np.zeros(n)
        self._ck(self._L.cpml_get_energy(self._h, _d(tot), _d(ek), _d(ep)))
        return tot, ek, ep

    def get_plane(self, field, kglobal=0):
        out = np.zeros((self.cfg.ny, self.cfg.nx))
        self._ck(self._L.cpml_get_plane(self._h, field, kglobal, _d(out)))
        return out

    def snapshot_begin(self, slot, field, kglobal=0):
        """Start the asynchronous pull of one (NX,NY) plane (device-side copy + D2H into pinned memory on a side
        stream); the time loop is not stalled.  Collect with snapshot_end(slot)."""
        self._ck(self._L.cpml_snapshot_begin(self._h, slot, field, kglobal))

    def snapshot_end(self, slot, copy=True):
        """The plane started by snapshot_begin(slot): a copy (default) or a zero-copy view of the library's pinned
        buffer (valid until the next snapshot_begin on that slot)."""
        ny, nx = self.cfg.ny, self.cfg.nx
        if copy:
            out = np.zeros((ny, nx))
            self._ck(self._L.cpml_snapshot_end(self._h, slot, _d(out), None))
            return out
        ptr = _dp()
        self._ck(self._L.cpml_snapshot_end(self._h, slot, None, C.byref(ptr)))
        return np.ctypeslib.as_array(ptr, shape=(ny, nx))

    def get_field(self, field):
        shape = (self.nzl, self.cfg.ny, self.cfg.nx) if self.cfg.ndim == 3 else (self.cfg.ny, self.cfg.nx)
        out = np.zeros(shape)
        self._ck(self._L.cpml_get_field(self._h, field, _d(out)))
        return out

    def get_maxnorm(self):
        v = C.c_double()
        self._ck(self._L.cpml_get_maxnorm(self._h, C.byref(v)))
        return v.value

    def enable_kernel_timing(self, on=True):
        self._ck(self._L.cpml_enable_kernel_timing(self._h, int(on)))

    def get_kernel_times(self, reset=True):
        a, b, n = C.c_double(), C.c_double(), C.c_int64()
        self._ck(self._L.cpml_get_kernel_times(self._h, C.byref(a), C.byref(b), C.byref(n), int(reset)))
        return a.value, b.value, n.value

    def algorithmic_bytes(self):
        a, b = C.c_double(), C.c_double()
        self._ck(self._L.cpml_algorithmic_bytes(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value


class MultiSolver:
    """One cpml_multi handle: the whole 3-D grid decomposed into z-slabs over `ngpus` devices of this node, driven
    from this one process (the reference's MPI layout without MPI; include/cpml_b200.h "cpml_multi_*")."""

    def __init__(self, ngpus, devices=None, **cfg):
        self._L = load()
        nz = cfg["nz"]
        cfg.setdefault("ndim", 3)
        cfg.setdefault("order", 2)
        cfg.setdefault("energy_bug_compat", True)
        keys = dict(lam="lambda_")
        c = CpmlConfig()
        for k, v in cfg.items():
            setattr(c, keys.get(k, k), int(v) if isinstance(v, bool) else v)
        c.nslabs, c.slab_rank, c.device = 1, 0, -1
        self.cfg = c
        self.ngpus = ngpus
        self._m = _H()
        dev = None
        if devices is not None:
            dev = np.ascontiguousarray(devices, dtype=np.int32)
        rc = self._L.cpml_multi_create(C.byref(c), ngpus, _i(dev) if dev is not None else None, C.byref(self._m))
        if rc:
            msg = self._L.cpml_multi_last_error(None)
            self._m = _H()
            raise CpmlError(rc, msg.decode() if msg else "cpml_multi_create")
        self.nzl = nz // ngpus

    def _ck(self, rc):
        if rc:
            msg = self._L.cpml_multi_last_error(self._m)
            raise CpmlError(rc, msg.decode() if msg else "")

    def close(self):
        if getattr(self, "_m", None):
            self._L.cpml_multi_destroy(self._m)
            self._m = _H()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def slab_launch_info(self, rank):
        h = _H()
        self._ck(self._L.cpml_multi_slab(self._m, rank, C.byref(h)))
        v = np.zeros(14, dtype=np.int32)
        rc = self._L.cpml_get_launch_info(h, _i(v), 14)
        if rc:
            raise CpmlError(rc, "cpml_get_launch_info")
        return {"tma": int(v[0]), "peer_sides": int(v[9])}

    def set_profiles(self, axis, prof):
        arrs = [_f64(prof[k]) for k in PROFILE_KEYS]
        self._ck(self._L.cpml_multi_set_profiles(self._m, axis, *[_d(a) for a in arrs], arrs[0].size))

    def set_attenuation(self, tau_epsilon_nu1, tau_sigma_nu1, tau_epsilon_nu2, tau_sigma_nu2):
        arrs = [_f64(a) for a in (tau_epsilon_nu1, tau_sigma_nu1, tau_epsilon_nu2, tau_sigma_nu2)]
        self._ck(self._L.cpml_multi_set_attenuation(self._m, arrs[0].size, *[_d(a) for a in arrs]))

    def set_source_series(self, force_x, force_y):
        fx, fy = _f64(force_x), _f64(force_y)
        self._ck(self._L.cpml_multi_set_source_series(self._m, _d(fx), _d(fy), fx.size))

    def set_receivers(self, ix_rec, iy_rec):
        ix = np.ascontiguousarray(ix_rec, dtype=np.int32)
        iy = np.ascontiguousarray(iy_rec, dtype=np.int32)
        self._ck(self._L.cpml_multi_set_receivers(self._m, _i(ix), _i(iy), ix.size))

    def reset(self):
        self._ck(self._L.cpml_multi_reset(self._m))

    def step(self, it):
        self._ck(self._L.cpml_multi_step(self._m, it))

    def run(self, it_begin, it_end):
        self._ck(self._L.cpml_multi_run(self._m, it_begin, it_end))

    def synchronize(self):
        self._ck(self._L.cpml_multi_synchronize(self._m))

    def get_seismograms(self):
        nrec, nstep = self.cfg.nrec, self.cfg.nstep
        sx, sy = np.zeros((nrec, nstep)), np.zeros((nrec, nstep))
        self._ck(self._L.cpml_multi_get_seismograms(self._m, _d(sx), _d(sy)))
        return sx, sy

    def get_seismograms_vz(self):
        sz = np.zeros((self.cfg.nrec, self.cfg.nstep))
        self._ck(self._L.cpml_multi_get_seismograms_vz(self._m, _d(sz)))
        return sz

    def get_energy(self):
        n = self.cfg.nstep
        tot, ek, ep = np.zeros(n), np.zeros(n), np.zeros(n)
        self._ck(self._L.cpml_multi_get_energy(self._m, _d(tot), _d(ek), _d(ep)))
        return tot, ek, ep

    def get_plane(self, field, kglobal):
        out = np.zeros((self.cfg.ny, self.cfg.nx))
        self._ck(self._L.cpml_multi_get_plane(self._m, field, kglobal, _d(out)))
        return out

    def get_field(self, field):
        out = np.zeros((self.cfg.nz, self.cfg.ny, self.cfg.nx))
        self._ck(self._L.cpml_multi_get_field(self._m, field, _d(out)))
        return out

    def get_maxnorm(self):
        v = C.c_double()
        self._ck(self._L.cpml_multi_get_maxnorm(self._m, C.byref(v)))
        return v.value
