"""CPU tests of the 3-D viscoelastic oracle (oracle/cpml_oracle_visco.c): closed-form set-up
constants of seismic_CPML_3D_viscoelastic_MPI.f90, the independent numpy restatement (which also
checks the analysis of the reference's incomplete halo exchange), slab-count
behaviour, and the golden vector."""
import math
import os

import numpy as np
import pytest

import refcfg
from oracle import oracle as O
from oracle.np_restatement_visco import run_3d_visco_np

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_setup_constants_visco():
    """Closed forms of 3D-visco :450-477, :547, :856 with the Carcione (1993) relaxation times."""
    tau = refcfg.TAU_CARCIONE_1993
    taumax, taumin = refcfg.visco_taumax_taumin(tau)
    assert taumax == pytest.approx(0.0352 / 0.0287, rel=1e-15)        # shear mode, first mechanism
    assert taumin == pytest.approx(0.0334 / 0.0303, rel=1e-15)
    c = refcfg.cfgv3d(nx=210, ny=800, nz=220, nstep=10, npml=10, rec_scale=1.0)   # the shipped grid
    # d0 = -(NPOWER+1) cp sqrt(taumax) ln(Rcoef) / (2 L), L = 40 m  (:547)
    d0 = 3.0 * 3000.0 * math.sqrt(taumax) * math.log(1e4) / 80.0
    px = c["prof_x"]
    # profile at i = 1: abscissa_normalized = 1 -> d = d0, K = K_MAX_PML = 7, alpha = 0
    assert px["K"][0] == 7.0
    assert px["b"][0] == pytest.approx(math.exp(-(d0 / 7.0) * 4e-4), rel=1e-14)
    assert px["a"][0] == pytest.approx(d0 * (px["b"][0] - 1.0) / (7.0 * d0), rel=1e-14)
    assert np.all(px["K"][10:200] == 1.0) and np.all(px["a"][10:200] == 0.0)
    # Courant number (:856) and the source / receiver positions (:205-208, :832-853)
    cn = 3000.0 * math.sqrt(taumax) * 4e-4 * math.sqrt(3.0 / 16.0)
    assert cn == pytest.approx(0.57546, rel=1e-4) and cn < 1.0
    assert (c["isource"], c["jsource"]) == (30, 161)
    # targets (xs+500, ys+500), (xs, ys+2260), (xs+500, ys+2260); grid abscissa = DELTAX * i
    assert list(c["ix_rec"]) == [155, 30, 155] and list(c["iy_rec"]) == [286, 726, 726]
    assert c["force_x"][0] == 0.0                                      # ANGLE_FORCE = 0 (:212)


@pytest.mark.parametrize("nproc", [1, 4])
def test_numpy_restatement_matches_c_oracle_bit_for_bit(nproc):
    c = refcfg.cfgv3d(nstep=40)
    r = O.run_3d_visco(**c, nproc=nproc, want_fields=True)
    n = run_3d_visco_np(**c, emulate_nproc=nproc)
    assert np.abs(r["sisvy"]).max() > 0
    for f in O.VISCO_FIELDS:
        assert np.array_equal(r[f], n[f]), f
    assert np.array_equal(r["sisvx"], n["sisvx"]) and np.array_equal(r["sisvy"], n["sisvy"])
    for e in ("total_energy", "energy_kinetic", "energy_potential"):
        assert refcfg.rel_l2(n[e], r[e]) <= 1e-12


def test_slab_count_behaviour():
    """With every halo plane exchanged the result does not depend on the number of slabs; with the
    reference's exchange it does (quirk B6), and one slab has no interface to lose taps at."""
    c = refcfg.cfgv3d(nstep=60)
    c1 = O.run_3d_visco(**c, nproc=1, complete_halos=True, want_fields=True)
    c4 = O.run_3d_visco(**c, nproc=4, complete_halos=True, want_fields=True)
    r1 = O.run_3d_visco(**c, nproc=1, want_fields=True)
    r4 = O.run_3d_visco(**c, nproc=4, want_fields=True)
    for f in O.VISCO_FIELDS:
        assert np.array_equal(c1[f], c4[f]), f
        assert np.array_equal(c1[f], r1[f]), f
    assert refcfg.rel_l2(c4["total_energy"], c1["total_energy"]) <= 1e-12
    assert not np.array_equal(r4["vy"], r1["vy"])
    assert 1e-6 < refcfg.rel_l2(r4["vy"], r1["vy"]) < 0.2


def test_golden_visco():
    g = np.load(os.path.join(GOLD, "cpml3d_visco_small.npz"))
    o = O.run_3d_visco(**refcfg.cfgv3d(), nproc=4)
    assert np.array_equal(o["sisvx"], g["sisvx"]) and np.array_equal(o["sisvy"], g["sisvy"])
    for e in ("total_energy", "energy_kinetic", "energy_potential"):
        assert np.array_equal(o[e], g[e])
    # qualitative behaviour: the source injects energy, nothing blows up
    assert o["total_energy"][-1] > o["total_energy"][5] > 0 and np.isfinite(o["total_energy"]).all()
