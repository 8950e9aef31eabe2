#!/usr/bin/env python
"""Golden vectors from the REFERENCE PROGRAMS THEMSELVES (run where /root/reference exists; the vectors travel).

There is no Fortran compiler in the image, so the reference cannot be built.  oracle/f90_exec.py executes the main
program of a reference `.f90` file from its own source text (statement-by-statement transliteration to Python, IEEE
double, source order, MPI ranks as threads).  This script runs all six programs of the hot path that way on reduced
grids -- `parameter` values overridden the way a user edits them -- and stores what the programs computed: C-PML
profiles, source and receiver indices, seismograms, energies and the final wavefields.

tests/test_reference_vectors.py then checks the C oracle (and, on a GPU box, the CUDA kernels) against these files, bit
for bit.  What is not the reference's: the relaxation times handed to the viscoelastic programs (their SolvOpt fit lives
in another file of the reference; it is pinned separately, tests/test_attenuation_fit.py) and, in the 3-D viscoelastic
program, the receiver offsets scaled by 0.04 so that they fall inside the reduced grid.

Usage: python tests/golden/make_reference_vectors.py [case ...]      (about ten minutes for all of them)
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import f90_exec as F      # noqa: E402
import refcfg                          # noqa: E402

REF = os.environ.get("CPML_REFERENCE_DIR", "/root/reference")
PROFILE_NAMES = {"a": "a_%s", "b": "b_%s", "K": "k_%s", "a_half": "a_%s_half", "b_half": "b_%s_half", "K_half": "k_%s_half"}

CASES = {
    "ref_3d_iso_np2": dict(kind="3d_iso", program="seismic_CPML_3D_isotropic_MPI_OpenMP.f90",
                           nx=20, ny=24, nz=12, npml=4, nstep=80, nproc=2, k_max=1.0),
    "ref_3d_iso_kmax3_np2": dict(kind="3d_iso", program="seismic_CPML_3D_isotropic_MPI_OpenMP.f90",
                                 nx=16, ny=18, nz=12, npml=3, nstep=40, nproc=2, k_max=3.0),
    "ref_3d_iso_np4": dict(kind="3d_iso", program="seismic_CPML_3D_isotropic_MPI_OpenMP.f90",
                           nx=16, ny=18, nz=16, npml=3, nstep=30, nproc=4, k_max=1.0),
    # the single-precision build the reference endorses (:114-116, "declare everything real"): every double precision
    # entity a 4-byte real, literals keep their kind -- the tolerance datum of cpml_config.precision = 1
    "ref_3d_iso_single_np2": dict(kind="3d_iso", program="seismic_CPML_3D_isotropic_MPI_OpenMP.f90",
                                  nx=20, ny=24, nz=12, npml=4, nstep=80, nproc=2, k_max=1.0, single=True),
    "ref_2d_second": dict(kind="2d_iso", program="seismic_CPML_2D_isotropic_second_order.f90", order=2,
                          nx=30, ny=44, npml=5, nstep=150, ydeb=300.0, yfin=80.0),
    "ref_2d_fourth": dict(kind="2d_iso", program="seismic_CPML_2D_isotropic_fourth_order.f90", order=4,
                          nx=30, ny=44, npml=5, nstep=150, ydeb=300.0, yfin=80.0),
    # the programs AT THEIR DEFAULT CONFIGURATION: the 2-D ones exactly as shipped (101 x 641, all 2000 / 4000 steps), the
    # 3-D one on its own x-y grid, source and receivers with NZ = 32 on two ranks for 1000 steps (both receivers have
    # seen the wave).  These use the vectorising mode of f90_exec (bit-identical to the scalar one, checked on every
    # array of the small cases) and take minutes to an hour: name them on the command line.  Fields: SHA-256.
    "ref_2d_second_default": dict(kind="2d_iso", program="seismic_CPML_2D_isotropic_second_order.f90", order=2,
                                  default=True, slow=True),
    "ref_2d_fourth_default": dict(kind="2d_iso", program="seismic_CPML_2D_isotropic_fourth_order.f90", order=4,
                                  default=True, slow=True),
    "ref_3d_iso_xy_default": dict(kind="3d_iso", program="seismic_CPML_3D_isotropic_MPI_OpenMP.f90", default=True, slow=True,
                                  nz=32, nproc=2, nstep=1000, k_max=1.0),
    # ... and for all 2500 steps of the program (two hours)
    "ref_3d_iso_xy_default_full": dict(kind="3d_iso", program="seismic_CPML_3D_isotropic_MPI_OpenMP.f90", default=True, slow=True,
                                       nz=32, nproc=2, nstep=2500, k_max=1.0),
    # mid-size viscoelastic runs (vectorising mode, fields as SHA-256): long enough for the wave to cross the receivers
    # and enter every shell
    "ref_3d_visco_mid_np2": dict(kind="3d_visco", program="seismic_CPML_3D_viscoelastic_MPI.f90", slow=True, fast=True, hash_fields=True,
                                 nx=64, ny=72, nz=24, npml=6, nstep=300, nproc=2, rec_scale=0.08),
    "ref_3d_visco_mid_np4": dict(kind="3d_visco", program="seismic_CPML_3D_viscoelastic_MPI.f90", slow=True, fast=True, hash_fields=True,
                                 nx=64, ny=72, nz=32, npml=6, nstep=300, nproc=4, rec_scale=0.08),
    "ref_2d_visco_second_mid": dict(kind="2d_visco", program="seismic_CPML_2D_velocity_and_stress_second_order_viscoelastic.f90",
                                    order=2, nx=201, ny=201, npml=10, nstep=1200, slow=True, fast=True, hash_fields=True),
    "ref_2d_visco_fourth_mid": dict(kind="2d_visco", program="seismic_CPML_2D_velocity_and_stress_fourth_order_viscoelastic.f90",
                                    order=4, nx=201, ny=201, npml=10, nstep=1200, slow=True, fast=True, hash_fields=True),
    "ref_3d_visco_np2": dict(kind="3d_visco", program="seismic_CPML_3D_viscoelastic_MPI.f90",
                             nx=32, ny=30, nz=12, npml=4, nstep=60, nproc=2, rec_scale=0.04),
    "ref_3d_visco_np4": dict(kind="3d_visco", program="seismic_CPML_3D_viscoelastic_MPI.f90",       # quirk B6 depends on NPROC
                             nx=32, ny=30, nz=16, npml=4, nstep=50, nproc=4, rec_scale=0.04),
    "ref_2d_visco_second": dict(kind="2d_visco", program="seismic_CPML_2D_velocity_and_stress_second_order_viscoelastic.f90",
                                order=2, nx=61, ny=71, npml=5, nstep=220),
    "ref_2d_visco_fourth": dict(kind="2d_visco", program="seismic_CPML_2D_velocity_and_stress_fourth_order_viscoelastic.f90",
                                order=4, nx=61, ny=71, npml=5, nstep=220),
}


def _interior(sp, name, n):
    """Elements 1..n of every axis of a program array, whatever its declared lower bounds."""
    lo = [int(eval(b[0], dict(sp))) for b in sp["_bounds"][name]]
    return sp[name][tuple(slice(1 - l, m + 1 - l) for l, m in zip(lo, n))]


def _profiles(sp, axes):
    return {f"prof_{ax}_{k}": np.array(sp[fmt % ax]) for ax in axes for k, fmt in PROFILE_NAMES.items()}


def _fit_from(tau):
    """Stands in for compute_attenuation_coeffs: first call of a rank -> the nu1 times, second -> nu2."""
    import threading
    count = {}

    def fit(n_sls, q, f0, fmin, fmax, tau_eps, tau_sig):
        me = threading.get_ident()
        count[me] = count.get(me, 0) + 1
        key = "nu1" if count[me] % 2 == 1 else "nu2"
        tau_eps[:] = tau["tau_epsilon_" + key]
        tau_sig[:] = tau["tau_sigma_" + key]
    return fit


def run_case(name):
    c = CASES[name]
    path = os.path.join(REF, c["program"])
    out = {}
    if c["kind"] == "3d_iso":
        if c.get("default"):          # the program's own x-y grid, source and receivers; only NZ, NPROC, NSTEP differ
            ov = {"NZ": c["nz"], "NPROC": c["nproc"], "NSTEP": c["nstep"]}
        else:
            ov = {"NX": c["nx"], "NY": c["ny"], "NZ": c["nz"], "NPROC": c["nproc"], "NSTEP": c["nstep"], "NPOINTS_PML": c["npml"],
                  "K_MAX_PML": f"{c['k_max']!r}d0", "ydeb": f"{(c['ny'] // 3) * 10}.d0", "yfin": "30.d0"}     # receivers as in refcfg.cfg3d
        sp = F.run_program(path, {k: str(v) for k, v in ov.items()}, nproc=c["nproc"], single=c.get("single", False),
                           vectorize=c.get("default", False))
        r = sp[sp[0]["rank_cut_plane"]]
        c = dict(c, nx=int(r["nx"]), ny=int(r["ny"]), npml=int(r["npoints_pml"]))
        nzl = c["nz"] // c["nproc"]
        out.update(_profiles(r, "xyz"))
        out.update(sisvx=r["sisvx"].T, sisvy=r["sisvy"].T, total_energy=r["total_energy"])
        for f in ("vx", "vy", "vz", "sigmaxx", "sigmayy", "sigmazz", "sigmaxy", "sigmaxz", "sigmayz"):
            out[f] = np.concatenate([_interior(q, f, (c["nx"], c["ny"], nzl)) for q in sp], axis=2).transpose(2, 1, 0)
    elif c["kind"] == "2d_iso":
        if c.get("default"):          # the program as shipped (at most a shorter NSTEP)
            ov = {"NSTEP": c["nstep"]} if "nstep" in c else {}
        else:
            ov = {"NX": c["nx"], "NY": c["ny"], "NSTEP": c["nstep"], "NPOINTS_PML": c["npml"],
                  "ydeb": f"{c['ydeb']!r}d0", "yfin": f"{c['yfin']!r}d0"}
        r = F.run_program(path, {k: str(v) for k, v in ov.items()}, vectorize=c.get("default", False))[0]
        c = dict(c, nx=int(r["nx"]), ny=int(r["ny"]), nstep=int(r["nstep"]), npml=int(r["npoints_pml"]))
        out.update(_profiles(r, "xy"))
        out.update(sisvx=r["sisvx"].T, sisvy=r["sisvy"].T, energy_kinetic=r["total_energy_kinetic"],
                   energy_potential=r["total_energy_potential"])
        for f, mine in (("vx", "vx"), ("vy", "vy"), ("sigmaxx", "sigmaxx"), ("sigmayy", "sigmayy"), ("sigmaxy", "sigmaxy")):
            src = f if f in r["_bounds"] else f.replace("sigma", "sigma_")
            out[mine] = _interior(r, src, (c["nx"], c["ny"])).T
    elif c["kind"] == "3d_visco":
        tau = refcfg.TAU_CARCIONE_1993
        ov = {"NX": c["nx"], "NY": c["ny"], "NZ": c["nz"], "NPROC": c["nproc"], "NSTEP": c["nstep"], "NPOINTS_PML": c["npml"]}
        sc = f"{c['rec_scale']!r}d0"
        edits = [(r"^(xrec|yrec)\((\d)\)\s*=\s*(\w+)\s*\+\s*(\d+\.d0)\s*$", r"\1(\2)=\3+\4*" + sc)]
        sp = F.run_program(path, {k: str(v) for k, v in ov.items()}, nproc=c["nproc"],
                           externals={"compute_attenuation_coeffs": _fit_from(tau)}, edits=edits, vectorize=c.get("fast", False))
        r = sp[sp[0]["rank_cut_plane"]]
        nzl = c["nz"] // c["nproc"]
        out.update(_profiles(r, "xyz"))
        out.update(sisvx=r["sisvx"].T, sisvy=r["sisvy"].T, total_energy=r["total_energy"],
                   energy_kinetic=r["total_energy_kinetic"], energy_potential=r["total_energy_potential"])
        for f in ("vx", "vy", "vz", "sigmaxx", "sigmayy", "sigmazz", "sigmaxy", "sigmaxz", "sigmayz"):
            out[f] = np.concatenate([_interior(q, f, (c["nx"], c["ny"], nzl)) for q in sp], axis=2).transpose(2, 1, 0)
    elif c["kind"] == "2d_visco":
        tau = refcfg.TAU_2D_VISCO
        isrc, jsrc = c["nx"] // 2 + 1, c["ny"] // 2 + 1
        ov = {"NX": c["nx"], "NY": c["ny"], "NSTEP": c["nstep"], "NPOINTS_PML": c["npml"],
              "xsource": f"{(isrc - 1) * 1.5!r}d0", "ysource": f"{(jsrc - 1) * 1.5!r}d0", "NREC": 2,
              "xdeb": "xsource + 20*deltax", "ydeb": "ysource + 20*deltax", "xfin": "xsource + 10*deltax",
              "yfin": "ysource - 25*deltax", "COMPUTE_ENERGY": ".true."}                                 # as in refcfg.cfgv2d
        r = F.run_program(path, {k: str(v) for k, v in ov.items()},
                          externals={"compute_attenuation_coeffs": _fit_from(tau)}, vectorize=c.get("fast", False))[0]
        out.update(_profiles(r, "xy"))
        out.update(sisvx=r["sisvx"].T, sisvy=r["sisvy"].T, sispressure=r["sispressure"].T,
                   energy_kinetic=r["total_energy_kinetic"], energy_potential=r["total_energy_potential"])
        for f, mine in (("sigma_xx", "sigmaxx"), ("sigma_yy", "sigmayy"), ("sigma_xy", "sigmaxy"), ("vx", "vx"), ("vy", "vy")):
            out[mine] = _interior(r, f, (c["nx"], c["ny"])).T
    out.update(isource=int(r["isource"]), jsource=int(r["jsource"]), ix_rec=np.array(r["ix_rec"]), iy_rec=np.array(r["iy_rec"]),
               deltat=float(r["deltat"]))
    if c.get("default") or c.get("hash_fields"):              # large fields: their SHA-256 instead of the arrays
        import hashlib
        for f in ("vx", "vy", "vz", "sigmaxx", "sigmayy", "sigmazz", "sigmaxy", "sigmaxz", "sigmayz"):
            if f in out:
                out["sha256_" + f] = hashlib.sha256(np.ascontiguousarray(out.pop(f), dtype=np.float64).tobytes()).hexdigest()
    out["meta"] = json.dumps(dict(case=name, **c))
    return out


def main(names):
    for name in names or [n for n, c in CASES.items() if not c.get("slow")]:
        t = time.time()
        out = run_case(name)
        fn = os.path.join(HERE, name + ".npz")
        np.savez_compressed(fn, **out)
        print(f"{name}: {time.time() - t:.0f} s, {os.path.getsize(fn) / 1024:.0f} KiB, max |sisvx| {np.abs(out['sisvx']).max():.3e}")


if __name__ == "__main__":
    main(sys.argv[1:])
