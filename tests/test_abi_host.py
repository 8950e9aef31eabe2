"""CPU tests of the C-ABI library: it loads, exports every symbol include/cpml_b200.h
declares, its host-side helpers reproduce the oracle's set-up bit for bit, and
cpml_create validates like the reference's checks.  No device compute here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import refcfg
from oracle import oracle as O
from seismic_cpml_b200 import lib as L
from seismic_cpml_b200 import programs as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cuda_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "cpml_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(cpml_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 30
    lib = C.CDLL(L.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"libcpml_b200.so does not export {name}"
    assert declared == set(L.SYMBOLS), declared ^ set(L.SYMBOLS)
    assert L.load().cpml_abi_version() == 1


def test_config_struct_matches_header_layout():
    # 20 int32 (the last one, precision, fills what used to be padding) + 13 doubles
    assert C.sizeof(L.CpmlConfig) == 80 + 13 * 8
    assert L.CpmlConfig.deltax.offset == 80


def test_host_profiles_match_oracle_bitwise():
    p = P.Params3DIso()
    s = P.setup_3d(p)
    c = refcfg.cfg3d(nx=101, ny=641, nz=640, nstep=p.NSTEP, npml=10)
    for mine, ref in ((s.prof_x, c["prof_x"]), (s.prof_y, c["prof_y"]), (s.prof_z, c["prof_z"])):
        for k in L.PROFILE_KEYS:
            assert np.array_equal(mine[k], ref[k]), k
    assert np.array_equal(s.force_x, c["force_x"]) and np.array_equal(s.force_y, c["force_y"])
    assert s.courant == pytest.approx(0.914522826396367, rel=1e-14)
    assert (p.ISOURCE, p.JSOURCE) == (80, 428)
    assert list(s.ix_rec) == [70, 80] and list(s.iy_rec) == [231, 31] and np.all(s.dist_rec == 0.0)


@pytest.mark.parametrize("order", [2, 4])
def test_host_setup_2d_matches_oracle_bitwise(order):
    p = P.Params2DIso(order=order)
    s = P.setup_2d(p)
    c = refcfg.cfg2d(order)
    assert p.NSTEP == (2000 if order == 2 else 4000) and p.DELTAT == (2e-3 if order == 2 else 1e-3)
    for mine, ref in ((s.prof_x, c["prof_x"]), (s.prof_y, c["prof_y"])):
        for k in L.PROFILE_KEYS:
            assert np.array_equal(mine[k], ref[k]), k
    assert np.array_equal(s.force_x, c["force_x"])
    assert list(s.ix_rec) == list(c["ix_rec"]) and list(s.iy_rec) == list(c["iy_rec"])
    assert s.courant == pytest.approx(0.933380951166243 if order == 2 else 0.933380951166243 / 2, rel=1e-14)
    for mine, ref in zip(s.material, (c["lam"], c["mu"], c["rho"])):
        assert np.array_equal(mine, ref)


def test_host_profile_kmax_and_one_sided():
    a = L.host_pml_profile(64, 5.0, 1e-3, 6, True, False, cp=2000.0, alpha_max_pml=30.0, k_max_pml=7.0)
    b = O.pml_profile(64, 5.0, 1e-3, 6, True, False, cp=2000.0, alpha_max_pml=30.0, k_max_pml=7.0)
    for k in L.PROFILE_KEYS:
        assert np.array_equal(a[k], b[k])
    assert np.all(a["a"][7:] == 0.0) and a["K"][0] == 7.0


def _bad(**over):
    base = dict(ndim=3, order=2, nx=32, ny=32, nz=32, nstep=10, npoints_pml=4, nrec=1, isource=10,
                jsource=10, deltax=10.0, deltay=10.0, deltaz=10.0, deltat=1e-3, lam=1e10, mu=1e10,
                lambdaplustwomu=3e10, rho=2800.0, cp=3300.0)
    base.update(over)
    return base


@pytest.mark.parametrize("over,code", [
    (dict(ndim=4), L.CPML_EINVAL),
    (dict(order=4), L.CPML_EINVAL),                       # 3-D iso is second order
    (dict(nslabs=3, slab_rank=0), L.CPML_ETOPOLOGY),      # NZ % nb_procs (3D-iso :391), odd (:388)
    (dict(nslabs=16, slab_rank=0), L.CPML_ETOPOLOGY),     # NZ_LOCAL < NPOINTS_PML (:394)
    (dict(nslabs=2, slab_rank=2), L.CPML_ETOPOLOGY),
    (dict(deltat=3e-3), L.CPML_ECFL),                     # Courant > 1 (:717)
    (dict(isource=40), L.CPML_EINVAL),
    (dict(ndim=2, nslabs=2), L.CPML_ETOPOLOGY),
    (dict(order=4, rheology=1, nz=30, emulate_nproc=3), L.CPML_ETOPOLOGY),   # 'nb_procs must be even' (3D-visco :522-523)
    (dict(order=4, rheology=1, emulate_nproc=5), L.CPML_ETOPOLOGY),
])
def test_create_validation(over, code):
    with pytest.raises(L.CpmlError) as e:
        L.Solver(**_bad(**over))
    assert e.value.code == code


@pytest.mark.skipif(_cuda_available(), reason="box has a GPU")
def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to run (it never routes to the oracle)."""
    with pytest.raises(L.CpmlError) as e:
        L.Solver(**_bad())
    assert e.value.code == L.CPML_ECUDA and "no CPU fallback" in str(e.value)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing in the product package references it."""
    pkg = os.path.join(ROOT, "seismic_cpml_b200")
    pat = re.compile(r"import\s+oracle|from\s+oracle|liboracle|cpml_oracle|np_restatement")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                assert not pat.search(open(os.path.join(dirpath, f)).read()), f


def test_output_writers(tmp_path):
    nt, nrec = 50, 2
    sx = np.linspace(-1, 1, nt * nrec).reshape(nrec, nt)
    sy = sx[::-1].copy()
    lib = L.load()
    assert lib.cpml_host_write_seismograms(str(tmp_path).encode(), L._d(sx), L._d(sy), nt, nrec, 2e-3) == 0
    for name in ("Vx_file_001.dat", "Vx_file_002.dat", "Vy_file_001.dat", "Vy_file_002.dat"):
        d = np.loadtxt(tmp_path / name)
        assert d.shape == (nt, 2)
        assert d[1, 0] == pytest.approx(2e-3, rel=1e-6)
    assert np.allclose(np.loadtxt(tmp_path / "Vx_file_002.dat")[:, 1], sx[1], rtol=1e-6, atol=1e-7)
    e = np.abs(sx[0]) + 1.0
    assert lib.cpml_host_write_energy_3d(str(tmp_path / "energy.dat").encode(), L._d(e), nt, 2e-3) == 0
    assert np.allclose(np.loadtxt(tmp_path / "energy.dat")[:, 1], e, rtol=1e-15)
    img = np.zeros((40, 30))
    img[20, 15] = 1.0
    img[5, 5] = -0.5
    ix = np.array([10], dtype=np.int32)
    iy = np.array([30], dtype=np.int32)
    assert lib.cpml_host_create_color_image(str(tmp_path).encode(), L._d(img), 30, 40, 5, 20, 10, L._i(ix),
                                            L._i(iy), 1, 4, 1, 1, 1, 1, 1) == 0
    toks = open(tmp_path / "image000005_Vx.pnm").read().split()
    assert toks[:4] == ["P3", "30", "40", "255"] and len(toks) == 4 + 3 * 30 * 40
    px = np.array(toks[4:], dtype=int).reshape(40, 30, 3)[::-1]   # row 0 = iy 1
    assert tuple(px[20, 15]) == (255, 0, 0)                        # positive maximum: red
    assert px[5, 5, 2] > 0 and px[5, 5, 0] == 0                    # negative: blue
    assert tuple(px[9, 19]) == (255, 157, 0)                       # source cross
    assert tuple(px[29, 9]) == (30, 180, 60)                       # receiver square
    assert tuple(px[0, 0]) == (0, 0, 0)                            # frame
