"""CPU tests of the oracle: closed-form set-up constants of the reference (SURVEY.md
Appendix C.1), golden vectors, and the independent numpy restatement."""
import math
import os

import numpy as np
import pytest

import refcfg
from oracle import oracle as O
from oracle.np_restatement import run_2d_np, run_3d_iso_np

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_setup_constants_2d_second_order():
    """Values the reference prints before the loop (closed forms of 2D-2nd :296,:513,:517-521)."""
    c = refcfg.cfg2d(2, nstep=100)
    px = c["prof_x"]
    assert px["a"][0] == pytest.approx(-0.495338702436472, rel=1e-14)
    assert px["b"][0] == pytest.approx(0.504661297563528, rel=1e-14)
    assert px["a"][9] == pytest.approx(-0.00668237069136469, rel=1e-13)
    assert px["b"][9] == pytest.approx(0.95463830814965, rel=1e-14)
    assert px["a"][10] == 0.0 and px["b"][10] == pytest.approx(0.956970898435118, rel=1e-14)
    assert px["a_half"][0] == pytest.approx(-0.460087856454662, rel=1e-14)
    assert px["b_half"][0] == pytest.approx(0.538272802416877, rel=1e-14)
    assert px["a_half"][90] == pytest.approx(-0.00167302343165511, rel=1e-12)
    assert px["a_half"][100] == pytest.approx(-0.529502305141351, rel=1e-14)
    # shell extents: a_x on [1,10] U [92,101]; a_x_half on [1,10] U [91,101]
    assert list(np.nonzero(px["a"])[0] + 1) == list(range(1, 11)) + list(range(92, 102))
    assert list(np.nonzero(px["a_half"])[0] + 1) == list(range(1, 11)) + list(range(91, 102))
    assert np.all(px["K"] == 1.0) and np.all(px["K_half"] == 1.0)
    # source position and receivers (2D-2nd :171-183, :486-509)
    assert (c["isource"], c["jsource"]) == (80, 428)
    assert list(c["ix_rec"]) == [70, 80] and list(c["iy_rec"]) == [231, 31]
    # source time function at it = 1 and it = 100 (:644-657)
    assert c["force_x"][0] == pytest.approx(788.498403142922, rel=1e-13)
    assert c["force_y"][0] == pytest.approx(-788.498403142922, rel=1e-13)
    term100 = c["force_x"][99] / math.sin(135.0 * refcfg.PI / 180.0)
    assert term100 == pytest.approx(-182663334.290646, rel=1e-13)


def test_setup_constants_3d():
    c = refcfg.cfg3d(nx=101, ny=641, nz=640, nstep=100, npml=10)
    px = c["prof_x"]
    assert px["a"][0] == pytest.approx(-0.421371261182048, rel=1e-14)
    assert px["b"][0] == pytest.approx(0.578628738817952, rel=1e-14)
    assert px["a"][9] == pytest.approx(-0.00537059775760104, rel=1e-13)
    assert px["b"][9] == pytest.approx(0.963542968239206, rel=1e-14)
    assert c["mu"] == pytest.approx(1.01645963229843e10, rel=1e-14)
    assert c["lam"] == pytest.approx(1.01628073540314e10, rel=1e-14)
    term100 = c["force_x"][99] / math.sin(135.0 * refcfg.PI / 180.0)
    assert term100 == pytest.approx(116083756.572156, rel=1e-13)
    # z profiles are the x formulas on NZ points, without the alpha clamp
    pz = c["prof_z"]
    assert list(np.nonzero(pz["a"])[0] + 1) == list(range(1, 11)) + list(range(631, 641))
    assert list(np.nonzero(pz["a_half"])[0] + 1) == list(range(1, 11)) + list(range(630, 641))


def test_fourth_order_top_pml_quirk_b4():
    """2D-4th :401 puts yorigintop at NY*DELTAY - L: a_y != 0 on [633,641], not [632,641]."""
    c2, c4 = refcfg.cfg2d(2, nstep=10), refcfg.cfg2d(4, nstep=10)
    nz2 = np.nonzero(c2["prof_y"]["a"])[0] + 1
    nz4 = np.nonzero(c4["prof_y"]["a"])[0] + 1
    assert nz2[10] == 632 and nz4[10] == 633 and nz2[-1] == nz4[-1] == 641


@pytest.mark.parametrize("name,order", [("cpml2d_second_default", 2), ("cpml2d_fourth_default", 4)])
def test_oracle_2d_matches_golden_prefix(name, order):
    """The first 400 steps of the shipped configurations reproduce the committed vectors."""
    g = np.load(os.path.join(GOLD, name + ".npz"))
    nfull = 2000 if order == 2 else 4000
    c = refcfg.cfg2d(order, nstep=nfull)
    c["nstep"] = 400      # same source series, shorter run
    o = O.run_2d(**c)
    for k in ("sisvx", "sisvy"):
        assert np.array_equal(o[k], g[k][:, :400]), k
    for k in ("energy_kinetic", "energy_potential"):
        assert np.array_equal(o[k], g[k][:400]), k


def test_golden_2d_second_order_physics():
    """Qualitative behaviour the reference documents through its energy plot: the energy
    rises while the source acts, then decays by orders of magnitude; no blow-up."""
    g = np.load(os.path.join(GOLD, "cpml2d_second_default.npz"))
    e = g["energy_kinetic"] + g["energy_potential"]
    assert e.argmax() + 1 == 108 and e.max() == pytest.approx(1.7148e8, rel=1e-4)
    assert e[-1] < 1e-7 * e.max()
    assert np.abs(g["sisvx"][0]).max() == pytest.approx(2.0544, rel=1e-4)
    assert np.abs(g["sisvy"][1]).max() == pytest.approx(0.61216, rel=1e-4)
    assert np.all(np.isfinite(g["sisvx"])) and np.all(np.isfinite(g["sisvy"]))


@pytest.mark.parametrize("order", [2, 4])
def test_numpy_restatement_2d_bit_identical(order):
    c = refcfg.cfg2d(order, nx=83, ny=131, nstep=600, npml=8, material="layered", ydeb=600.0, yfin=200.0)
    g = np.load(os.path.join(GOLD, f"cpml2d_layered_order{order}.npz"))
    o = O.run_2d(**c, want_fields=True)
    n = run_2d_np(**c)
    for k in ("sisvx", "sisvy", "vx", "vy", "sigmaxx", "sigmayy", "sigmaxy"):
        assert np.array_equal(o[k], n[k]), k
    for k in ("sisvx", "sisvy", "energy_kinetic", "energy_potential"):
        assert np.array_equal(o[k], g[k]), k
    for k in ("energy_kinetic", "energy_potential"):
        assert refcfg.rel_l2(n[k], o[k]) < 1e-13


def test_numpy_restatement_3d_bit_identical_and_slab_invariance():
    """3-D: numpy (single address space) == C oracle for 1, 2 and 4 emulated MPI slabs."""
    c = refcfg.cfg3d()
    g = np.load(os.path.join(GOLD, "cpml3d_iso_small.npz"))
    n = run_3d_iso_np(**c)
    for nproc in (1, 2, 4):
        o = O.run_3d_iso(**c, nproc=nproc, want_fields=True, want_planes=True)
        for k in ("sisvx", "sisvy", "vx", "vy", "vz", "sigmaxx", "sigmayy", "sigmazz", "sigmaxy",
                  "sigmaxz", "sigmayz"):
            assert np.array_equal(o[k], n[k]), (nproc, k)
        assert refcfg.rel_l2(o["total_energy"], n["total_energy"]) < 1e-13
        assert np.array_equal(o["sisvx"], g["sisvx"]) and np.array_equal(o["sisvy"], g["sisvy"])
        if nproc == 2:
            assert np.array_equal(o["total_energy"], g["total_energy"])
            assert np.array_equal(o["plane_vx"], g["plane_vx"])
    assert np.abs(g["sisvx"]).max() > 0.1


def test_oracle_3d_kmax_golden():
    c = refcfg.cfg3d(nx=30, ny=34, nz=32, nstep=120, npml=5, k_max=3.0)
    g = np.load(os.path.join(GOLD, "cpml3d_iso_kmax3.npz"))
    assert np.any(c["prof_x"]["K"] != 1.0)
    o = O.run_3d_iso(**c, nproc=2)
    assert np.array_equal(o["sisvx"], g["sisvx"]) and np.array_equal(o["total_energy"], g["total_energy"])
    n = run_3d_iso_np(**c)
    assert np.array_equal(o["sisvx"], n["sisvx"]) and np.array_equal(o["sisvy"], n["sisvy"])


def test_oracle_3d_energy_bug_flag():
    """Quirk B2: the reference potential energy counts yy twice and omits zz."""
    c = refcfg.cfg3d(nstep=60)
    a = O.run_3d_iso(**c, nproc=2, energy_bug_compat=True)
    b = O.run_3d_iso(**c, nproc=2, energy_bug_compat=False)
    assert np.array_equal(a["sisvx"], b["sisvx"])
    assert refcfg.rel_l2(a["total_energy"], b["total_energy"]) > 1e-3


def test_oracle_3d_topology_checks():
    c = refcfg.cfg3d(nstep=2)
    for nproc in (3, 6):          # odd, or NZ not a multiple (3D-iso :388-391)
        with pytest.raises(RuntimeError):
            O.run_3d_iso(**c, nproc=nproc)
    with pytest.raises(RuntimeError):   # NZ_LOCAL < NPOINTS_PML (:394)
        O.run_3d_iso(**c, nproc=10)
