"""Host drivers over the C ABI: the Fortran ISO_C_BINDING module is checked against the
header symbol by symbol (it cannot be compiled here: no Fortran compiler), the C++ driver
is built, and -- on a GPU -- run end to end and compared with the oracle through its
output files."""
import os
import re
import subprocess

import numpy as np
import pytest

import refcfg
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRV = os.path.join(ROOT, "drivers")


def _header_functions():
    h = open(os.path.join(ROOT, "include", "cpml_b200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int32_t|double|const char \*)\s*(cpml_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", h, flags=re.S):
        args = [a.strip() for a in m.group(2).split(",") if a.strip() and a.strip() != "void"]
        out[m.group(1)] = len(args)
    return out


def test_fortran_module_binds_every_symbol_with_matching_arity():
    funcs = _header_functions()
    assert len(funcs) >= 33
    f90 = open(os.path.join(DRV, "fortran", "cpml_b200_mod.f90")).read()
    f90 = re.sub(r"&\s*\n\s*", " ", f90)          # join continuation lines
    bound = {}
    for m in re.finditer(r"function\s+(cpml_[a-z0-9_]+)\s*\(([^)]*)\)\s*bind\(C,\s*name='([a-z0-9_]+)'\)", f90):
        assert m.group(1) == m.group(3)
        bound[m.group(1)] = len([a for a in m.group(2).split(",") if a.strip()])
    assert set(bound) == set(funcs), set(bound) ^ set(funcs)
    for name, n in funcs.items():
        assert bound[name] == n, (name, bound[name], n)
    # the derived type mirrors struct cpml_config: 18 + 1 int32, then 9 + 4 doubles
    t = f90[f90.index("type, bind(C) :: cpml_config"):f90.index("end type cpml_config")]
    ints = re.findall(r"integer\(c_int32_t\)\s*::\s*(.*)", t)
    reals = re.findall(r"real\(c_double\)\s*::\s*(.*)", t)
    assert sum(len(x.split(",")) for x in ints) == 20      # 20 named (the last one: precision; it fills the old padding)
    assert sum(len(x.split(",")) for x in reals) == 10     # 9 named + reserved_d(4)


def test_fortran_driver_keeps_the_reference_parameter_surface():
    src = open(os.path.join(DRV, "fortran", "seismic_CPML_3D_isotropic_b200.f90")).read()
    for name in ("NX = 101", "NY = 641", "NZ = 640", "DELTAX = 10.d0", "NSTEP = 2500", "DELTAT = 1.6d-3",
                 "NPOINTS_PML = 10", "ISOURCE = NX - 2*NPOINTS_PML - 1", "JSOURCE = 2 * NY / 3 + 1",
                 "ANGLE_FORCE = 135.d0", "NREC = 2", "IT_DISPLAY = 100", "f0 = 7.d0", "factor = 1.d7"):
        assert name in src, name
    for call in ("cpml_create", "cpml_set_profiles", "cpml_set_source_series", "cpml_set_receivers", "cpml_run",
                 "cpml_get_seismograms", "cpml_get_energy", "cpml_get_plane", "cpml_get_maxnorm", "cpml_destroy"):
        assert call in src, call


@pytest.fixture(scope="module")
def driver_exe():
    subprocess.run(["make", "-C", DRV, "-s"], check=True)
    exe = os.path.join(DRV, "xseismic_cpml")
    assert os.path.exists(exe)
    return exe


def test_cpp_driver_builds_and_fails_loudly_without_gpu(driver_exe, tmp_path):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    r = subprocess.run([driver_exe, "--program", "2d_second", "NSTEP=4", "--out", str(tmp_path), "--no-images"],
                       capture_output=True, text=True)
    assert "Courant number is 0.933380951166243" in r.stdout
    if not has_gpu:
        assert r.returncode != 0 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("program,order", [("2d_second", 2), ("2d_fourth", 4)])
def test_cpp_driver_2d_output_files_match_oracle(driver_exe, tmp_path, program, order):
    nstep = 400
    r = subprocess.run([driver_exe, "--program", program, f"NSTEP={nstep}", "--out", str(tmp_path)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "End of the simulation" in r.stdout and f"Time step # 5 out of {nstep}" in r.stdout
    c = refcfg.cfg2d(order, nstep=nstep)
    o = O.run_2d(**c)
    for rec in (1, 2):
        for comp, key in (("Vx", "sisvx"), ("Vy", "sisvy")):
            d = np.loadtxt(tmp_path / f"{comp}_file_{rec:03d}.dat")
            assert d.shape == (nstep, 2)
            # the files hold single-precision values, like the reference's sngl(...)
            assert np.array_equal(d[:, 1].astype(np.float32), o[key][rec - 1].astype(np.float32))
    e = np.loadtxt(tmp_path / "energy.dat")
    assert e.shape == (nstep, 4)
    assert np.allclose(e[:, 1], o["energy_kinetic"], rtol=1e-6, atol=0)
    assert os.path.exists(tmp_path / f"image{(200 if order == 4 else 100):06d}_Vx.pnm")


@pytest.mark.gpu
def test_cpp_driver_3d_small_grid(driver_exe, tmp_path):
    r = subprocess.run([driver_exe, "--program", "3d_iso", "NX=48", "NY=60", "NZ=40", "NSTEP=120", "NPOINTS_PML=6",
                        "ydeb=300", "yfin=100", "IT_DISPLAY=50", "--out", str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    d = np.loadtxt(tmp_path / "Vx_file_001.dat")
    e = np.loadtxt(tmp_path / "energy.dat")
    assert d.shape == (120, 2) and e.shape == (120, 2) and np.abs(d[:, 1]).max() > 0 and np.all(np.isfinite(e))
    assert os.path.exists(tmp_path / "image000100_Vy.pnm")
    stamp = open(tmp_path / "timestamp000100").read()            # 3D-iso :1219-1229
    assert " Time step #          100\n" in stamp and "Total energy =" in stamp and "Elapsed time in hh:mm:ss" in stamp
    vz = np.loadtxt(tmp_path / "Vz_file_001.dat")               # extension (quirk B7)
    assert vz.shape == (120, 2) and np.all(np.isfinite(vz))


def test_cpp_driver_viscoelastic_setup_uses_the_solvopt_fit(driver_exe, tmp_path):
    """The viscoelastic programs of the C++ driver: set-up phase on the host (relaxation times from the
    SolvOpt fit, printed like the reference does at 3D-visco :445-455), then the same loud failure without
    a GPU."""
    r = subprocess.run([driver_exe, "--program", "2d_visco_fourth", "NSTEP=4", "--out", str(tmp_path), "--no-images"],
                       capture_output=True, text=True)
    # the reference's own constants (analytical program :124-128)
    assert "tau_epsilon_nu1 = 0.02408158185753685 0.004699608990861351 0.0009567997872435925" in r.stdout
    assert "tau_sigma_nu2 = 0.0225091977942949 0.004501388007338097 0.0008917332095369118" in r.stdout
    assert "in i,j = 1535 1535" in r.stdout and "Courant number is 0.293333333333333" in r.stdout
    r3 = subprocess.run([driver_exe, "--program", "3d_visco", "NX=40", "NY=50", "NZ=40", "NSTEP=4", "--out", str(tmp_path),
                         "--no-images"], capture_output=True, text=True)
    assert "tau_epsilon_nu1 = 0.03433147438440785 0.003631112527072353" in r3.stdout
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        assert r.returncode != 0 and "no CPU fallback" in r.stderr
        assert r3.returncode != 0


@pytest.mark.gpu
@pytest.mark.parametrize("program,order", [("2d_visco_second", 2), ("2d_visco_fourth", 4)])
def test_cpp_driver_2d_viscoelastic_output_files_match_oracle(driver_exe, tmp_path, program, order):
    from seismic_cpml_b200 import programs as P
    kw = dict(NX=121, NY=101, NSTEP=200, xsource=90.0, ysource=75.0, xdeb=120.0, ydeb=100.0, xfin=120.0, yfin=100.0)
    args = [f"{k}={v}" for k, v in kw.items()]
    r = subprocess.run([driver_exe, "--program", program, *args, "IT_DISPLAY=100", "--out", str(tmp_path)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "End of the simulation" in r.stdout
    p = P.Params2DVisco(order=order, **kw)
    s = P.setup_2d_visco(p)
    o = O.run_2d_visco(order=order, nx=p.NX, ny=p.NY, deltax=p.DELTAX, deltay=p.DELTAY, deltat=p.DELTAT, nstep=p.NSTEP,
                       npoints_pml=p.NPOINTS_PML, isource=p.ISOURCE, jsource=p.JSOURCE, lam=s.material[0], mu=s.material[1],
                       rho=s.material[2], tau_epsilon_nu1=p.tau_epsilon_nu1, tau_sigma_nu1=p.tau_sigma_nu1,
                       tau_epsilon_nu2=p.tau_epsilon_nu2, tau_sigma_nu2=p.tau_sigma_nu2, prof_x=s.prof_x, prof_y=s.prof_y,
                       force_x=s.force_x, force_y=s.force_y, ix_rec=s.ix_rec, iy_rec=s.iy_rec)
    for name, key, shift in (("Vx_file_001.dat", "sisvx", 0.0), ("Vy_file_half_a_grid_cell_away_from_Vx_001.dat", "sisvy", 0.0),
                             ("pressure_file_001.dat", "sispressure", 0.5 * p.DELTAT)):
        d = np.loadtxt(tmp_path / name)
        assert d.shape == (p.NSTEP, 2)
        assert np.abs(o[key][0]).max() > 0
        assert np.array_equal(d[:, 1].astype(np.float32), o[key][0].astype(np.float32)), name
        # time axis of write_seismograms (2D-visco-4th :1168,1178): (it-1) DELTAT - t0 (+ DELTAT/2 for the pressure)
        assert np.allclose(d[:, 0], np.arange(p.NSTEP) * p.DELTAT - p.t0 + shift, rtol=0, atol=1e-7)
    assert os.path.exists(tmp_path / "image000100_Vx.pnm")


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("program", ["3d_iso", "3d_visco"])
@pytest.mark.parametrize("ngpu,same", [(2, 1), (4, 1), (2, 0), (8, 0)])
def test_cpp_driver_ngpu_equals_single_gpu(driver_exe, tmp_path, program, ngpu, same):
    """NGPU= (the compiled driver's stand-in for the reference's NPROC MPI ranks: cpml_multi_*, one host thread, no
    MPI, no Python): the seismogram and energy FILES of an NGPU run equal those of the one-GPU run byte for byte
    (energy: to summation order).  same=1 puts every slab on device 0, so this also runs on a one-GPU box."""
    only = [int(x) for x in os.environ.get("CPML_TEST_WORLDS", "").split(",") if x.strip()]
    if only and (ngpu not in only or same):
        pytest.skip("CPML_TEST_WORLDS")
    if not same and _ngpu() < ngpu:
        pytest.skip(f"needs {ngpu} GPUs")
    common = ["NX=48", "NY=60", "NZ=48", "NSTEP=100", "NPOINTS_PML=5", "IT_DISPLAY=50", "--no-images"]
    if program == "3d_iso":
        common += ["ydeb=300", "yfin=100"]
    else:
        common += ["NPROC=4", "xrec1=120", "yrec1=100", "xrec2=100", "yrec2=120", "xrec3=120", "yrec3=120", "NSTEP=260"]
    outs = []
    for extra, sub in (([], "one"), ([f"NGPU={ngpu}", f"SAME_DEVICE={same}"], "multi")):
        d = tmp_path / sub
        d.mkdir()
        r = subprocess.run([driver_exe, "--program", program] + common + extra + ["--out", str(d)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr + r.stdout[-500:]
        outs.append(d)
        if extra:
            assert f"z-slab decomposition over {ngpu} GPU slabs" in r.stdout
    peak = 0.0
    for name in ("Vx_file_001.dat", "Vy_file_001.dat", "Vy_file_002.dat", "Vz_file_001.dat"):
        a, b = open(outs[0] / name).read(), open(outs[1] / name).read()
        assert a == b and len(a) > 1000, name
        peak = max(peak, np.abs(np.loadtxt(outs[0] / name)[:, 1]).max())
    assert peak > 0                    # (the viscoelastic source acts along y: Vx stays zero at receivers on its axis)
    e0, e1 = np.loadtxt(outs[0] / "energy.dat"), np.loadtxt(outs[1] / "energy.dat")
    assert np.allclose(e0, e1, rtol=1e-11, atol=0)


@pytest.mark.gpu
def test_cpp_driver_3d_viscoelastic_small_grid(driver_exe, tmp_path):
    r = subprocess.run([driver_exe, "--program", "3d_visco", "NX=60", "NY=80", "NZ=48", "NSTEP=60", "NPROC=2", "IT_DISPLAY=50",
                        "xrec1=200", "yrec1=200", "xrec2=160", "yrec2=240", "xrec3=200", "yrec3=240", "--out", str(tmp_path)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    d = np.loadtxt(tmp_path / "Vy_file_001.dat")
    e = np.loadtxt(tmp_path / "energy.dat")
    assert d.shape == (60, 2) and e.shape == (60, 4) and np.abs(d[:, 1]).max() > 0 and np.all(np.isfinite(e))
    assert d[0, 0] == pytest.approx(-1.2 / 18.0, rel=1e-6)          # time axis minus t0 (3D-visco :1603)
    assert np.loadtxt(tmp_path / "Vz_file_003.dat").shape == (60, 2)
    assert "Total energy =" in r.stdout


def test_fortran_viscoelastic_driver_keeps_the_reference_parameter_surface():
    src = open(os.path.join(DRV, "fortran", "seismic_CPML_2D_viscoelastic_b200.f90")).read()
    for name in ("NX = 2001", "NY = 2001", "DELTAX = 1.5d0", "DELTAT = 2.2d-4", "NSTEP = 5200", "NPOINTS_PML = 10",
                 "cp_unrelaxed = 2000.d0", "f0 = 35.d0", "factor = 1.d0", "xsource = 1500.d0", "ANGLE_FORCE = 0.d0",
                 "NREC = 1", "xdeb = 2301.d0", "IT_DISPLAY = 200", "N_SLS = 3", "Qp = 65.d0", "Qs = 55.d0",
                 "COMPUTE_ENERGY = .false."):
        assert name in src, name
    for call in ("cpml_host_attenuation_fit", "cpml_create", "cpml_set_profiles", "cpml_set_material_2d",
                 "cpml_set_attenuation", "cpml_set_source_series", "cpml_set_receivers", "cpml_run",
                 "cpml_get_seismograms", "cpml_get_pressure_seismograms", "cpml_host_write_seismograms_visco",
                 "cpml_get_maxnorm", "cpml_destroy"):
        assert call in src, call
    # every library function it calls is bound by the module with the arity used here
    f90 = open(os.path.join(DRV, "fortran", "cpml_b200_mod.f90")).read()
    for call in re.findall(r"\b(cpml_[a-z0-9_]+)\(", src):
        assert ("function " + call + "(") in f90 or ("subroutine " + call + "(") in f90, call


def _split_args(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        depth += ch == "("
        depth -= ch == ")"
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    out.append(cur)
    return [a for a in out if a.strip()]


def test_every_fortran_driver_calls_the_module_with_matching_arity():
    """The Fortran drivers cannot be compiled here; at least every call into the library must name a function
    the module binds and pass the number of arguments the binding declares."""
    mod = re.sub(r"&\s*\n\s*", " ", open(os.path.join(DRV, "fortran", "cpml_b200_mod.f90")).read())
    arity = {m.group(1): len([a for a in m.group(2).split(",") if a.strip()])
             for m in re.finditer(r"(?:function|subroutine)\s+(cpml_[a-z0-9_]+)\s*\(([^)]*)\)", mod)}
    drivers = sorted(f for f in os.listdir(os.path.join(DRV, "fortran")) if f.startswith("seismic_") and f.endswith(".f90"))
    assert len(drivers) >= 3
    for fn in drivers:
        src = open(os.path.join(DRV, "fortran", fn)).read()
        src = "\n".join(line for line in src.split("\n") if not line.lstrip().startswith("!"))
        src = re.sub(r"&\s*\n\s*", " ", src)
        calls = 0
        for m in re.finditer(r"\b(cpml_[a-z0-9_]+)\(", src):
            name, k, depth = m.group(1), m.end(), 1
            while depth > 0:
                depth += src[k] == "("
                depth -= src[k] == ")"
                k += 1
            assert name in arity, (fn, name)
            assert arity[name] == len(_split_args(src[m.end():k - 1])), (fn, name)
            calls += 1
        assert calls >= 15, fn
        assert src.count("program ") >= 2 and "use cpml_b200" in src and "implicit none" in src, fn
