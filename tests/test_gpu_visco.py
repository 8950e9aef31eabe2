"""GPU parity tests of the 3-D viscoelastic solver (fourth order, N_SLS = 2): the CUDA path through
the C ABI against the CPU oracle (oracle/cpml_oracle_visco.c) on the same inputs.

Tolerance: north_star asks for relative L2 <= 1e-5 on seismograms and the energy trace.  The
kernels keep the reference's operation order and every division (-fmad=false), so all 15 fields
and the seismograms are bit-identical to the oracle; energies are sums in a different order.
`emulate_nproc = n` must reproduce the reference run on n MPI ranks (incomplete halo exchange,
SURVEY.md quirk B6) on one GPU.
"""
import os

import numpy as np
import pytest

import refcfg
from oracle import oracle as O
from seismic_cpml_b200 import lib as L

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-5
TOL_ENERGY = 1e-11
FV = L.FIELDS_3D_VISCO


def solver_visco(c, emulate_nproc=0, nslabs=1, slab_rank=0, device=-1, **kw):
    s = L.Solver(ndim=3, order=4, rheology=1, emulate_nproc=emulate_nproc, nx=c["nx"], ny=c["ny"], nz=c["nz"], **kw,
                 nstep=c["nstep"], npoints_pml=c["npoints_pml"], nrec=len(c["ix_rec"]), isource=c["isource"],
                 jsource=c["jsource"], nslabs=nslabs, slab_rank=slab_rank, device=device, deltax=c["deltax"],
                 deltay=c["deltay"], deltaz=c["deltaz"], deltat=c["deltat"], lam=c["lam"], mu=c["mu"],
                 rho=c["rho"], cp=c["cp_eff"])
    s.set_profiles(L.AXIS_X, c["prof_x"])
    s.set_profiles(L.AXIS_Y, c["prof_y"])
    s.set_profiles(L.AXIS_Z, c["prof_z"])
    s.set_attenuation(c["tau_epsilon_nu1"], c["tau_sigma_nu1"], c["tau_epsilon_nu2"], c["tau_sigma_nu2"])
    s.set_source_series(c["force_x"], c["force_y"])
    s.set_receivers(c["ix_rec"], c["iy_rec"])
    return s


def check_against(s, o, c):
    sx, sy = s.get_seismograms()
    assert np.abs(o["sisvy"]).max() > 0
    for g, r in ((sx, o["sisvx"]), (sy, o["sisvy"])):
        assert np.all(np.isfinite(g))
        assert refcfg.rel_l2(g, r) <= TOL
        assert np.array_equal(g, r), f"max abs diff {np.abs(g - r).max()}"
    for f, name in enumerate(FV):
        got = s.get_field(f)
        assert np.array_equal(got, o[name]), (name, np.abs(got - o[name]).max())
    tot, ek, ep = s.get_energy()
    assert refcfg.rel_l2(tot, o["total_energy"]) <= TOL_ENERGY
    assert refcfg.rel_l2(ek, o["energy_kinetic"]) <= TOL_ENERGY
    assert refcfg.rel_l2(ep, o["energy_potential"]) <= TOL_ENERGY
    assert s.get_maxnorm() == pytest.approx(o["vnorm"], rel=1e-15)


@pytest.mark.parametrize("vkernel", ["ws", "reg"])
@pytest.mark.parametrize("shape", [(38, 46, 40, 6), (64, 33, 48, 5), (97, 40, 32, 8)])
@pytest.mark.parametrize("nproc", [1, 4])
def test_visco_matches_oracle(shape, nproc, vkernel, monkeypatch):
    """All 15 fields, seismograms and the three energy traces after 120 steps on ragged grids, for
    the single-rank semantics and for the reference's default NPROC = 4; with the TMA-staged velocity kernel
    (ws, the default) and the register-marching one (CPML_VKERNEL=reg)."""
    monkeypatch.setenv("CPML_VKERNEL", vkernel)
    nx, ny, nz, npml = shape
    c = refcfg.cfgv3d(nx=nx, ny=ny, nz=nz, npml=npml, nstep=120)
    o = O.run_3d_visco(**c, nproc=nproc, want_fields=True)
    with solver_visco(c, emulate_nproc=nproc) as s:
        assert s.launch_info()["tma"] == (2 if vkernel == "ws" else 0)
        s.run(1, c["nstep"])
        check_against(s, o, c)


@pytest.mark.parametrize("tx", [64, 104, 108])
@pytest.mark.parametrize("stages", [1, 2, 3])
def test_visco_ws_velocity_tiles(tx, stages, monkeypatch):
    """Every tile width / ring depth of the TMA-staged velocity kernel: two x tiles per row (the second one ragged),
    three z chunks, K_MAX_PML = 7 shells on all six faces; bitwise."""
    monkeypatch.setenv("CPML_VWS_TX", str(tx))
    monkeypatch.setenv("CPML_VWS_STAGES", str(stages))
    monkeypatch.setenv("CPML_VWS_ZCHUNKS", "3")
    c = refcfg.cfgv3d(nx=120, ny=37, nz=36, npml=5, nstep=70)
    o = O.run_3d_visco(**c, nproc=2, want_fields=True)
    with solver_visco(c, emulate_nproc=2) as s:
        s.run(1, c["nstep"])
        check_against(s, o, c)


def test_visco_golden():
    g = np.load(os.path.join(GOLD, "cpml3d_visco_small.npz"))
    c = refcfg.cfgv3d()
    with solver_visco(c, emulate_nproc=4) as s:
        s.run(1, c["nstep"])
        sx, sy = s.get_seismograms()
        tot, ek, ep = s.get_energy()
    assert np.array_equal(sx, g["sisvx"]) and np.array_equal(sy, g["sisvy"])
    assert refcfg.rel_l2(tot, g["total_energy"]) <= TOL_ENERGY
    assert refcfg.rel_l2(ek, g["energy_kinetic"]) <= TOL_ENERGY
    assert refcfg.rel_l2(ep, g["energy_potential"]) <= TOL_ENERGY


@pytest.mark.parametrize("tile", ["64x4", "16x16"])
def test_visco_tile_and_chunk_independent(tile, monkeypatch):
    """Results must not depend on the launch geometry (z chunk of 5 planes, other thread tiles)."""
    import subprocess
    import sys
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import refcfg; from oracle import oracle as O; import test_gpu_visco as T\n"
        "c = refcfg.cfgv3d(nx=70, ny=41, nz=36, npml=5, nstep=60)\n"
        "o = O.run_3d_visco(**c, nproc=2, want_fields=True)\n"
        "s = T.solver_visco(c, emulate_nproc=2); s.run(1, c['nstep']); T.check_against(s, o, c); s.close(); print('ok')\n"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    tx, ty = tile.split("x")
    env = dict(os.environ, CPML_VTX=tx, CPML_VTY=ty, CPML_VKCHUNK="5")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


def test_visco_slabs_in_one_process_match_whole_grid():
    """Two slab handles on one GPU exchanging planes with cpml_copy_plane (complete halos: two
    planes one way, one the other, per field) == the whole grid, for emulate_nproc = 4."""
    c = refcfg.cfgv3d(nx=40, ny=37, nz=48, npml=5, nstep=80)
    o = O.run_3d_visco(**c, nproc=4, want_fields=True)
    slabs = [solver_visco(c, emulate_nproc=4, nslabs=2, slab_rank=r) for r in range(2)]
    nzl = slabs[0].nzl
    lo, hi = slabs
    VX, VY, VZ, SZZ, SXZ, SYZ = 0, 1, 2, 5, 7, 8
    for it in range(1, c["nstep"] + 1):
        for f in (VX, VY):                       # 3D-visco :962-970 + the plane the reference never sends
            lo.copy_plane_from(nzl + 1, hi, 1, f); lo.copy_plane_from(nzl + 2, hi, 2, f)
            hi.copy_plane_from(0, lo, nzl, f)
        hi.copy_plane_from(-1, lo, nzl - 1, VZ); hi.copy_plane_from(0, lo, nzl, VZ)      # :972-975
        lo.copy_plane_from(nzl + 1, hi, 1, VZ)
        for s in slabs:
            s.step_stress(it)
        lo.copy_plane_from(nzl + 1, hi, 1, SZZ); lo.copy_plane_from(nzl + 2, hi, 2, SZZ)  # :1229-1232
        hi.copy_plane_from(0, lo, nzl, SZZ)
        for f in (SYZ, SXZ):                                                              # :1234-1242
            hi.copy_plane_from(-1, lo, nzl - 1, f); hi.copy_plane_from(0, lo, nzl, f)
            lo.copy_plane_from(nzl + 1, hi, 1, f)
        for s in slabs:
            s.step_velocity(it)
            s.step_finish(it)
    for s in slabs:
        s.synchronize()
    for f, name in enumerate(FV):
        got = np.concatenate([s.get_field(f) for s in slabs], axis=0)
        assert np.array_equal(got, o[name]), name
    e = slabs[0].get_energy()[0] + slabs[1].get_energy()[0]
    assert refcfg.rel_l2(e, o["total_energy"]) <= TOL_ENERGY
    sx, sy = slabs[0].get_seismograms()          # plane NZ/2 is the last plane of slab 0
    assert np.array_equal(sx, o["sisvx"]) and np.array_equal(sy, o["sisvy"])
    for s in slabs:
        s.close()


def test_visco_default_grid_window_equals_small_oracle():
    """BASELINE config 5 at the reference's own size (210 x 800 x 220, NPROC = 4), 8 steps: the
    window around the source must equal, bit for bit, the oracle run on a small grid that shares
    the source neighbourhood -- same x-min PML, same slab interface right above the source plane.
    Numerical domain of dependence: the fourth-order leapfrog moves information 4 cells per step,
    so a feature the two grids do not share at distance D from the source cannot change a point at
    distance w before step (2 D - w) / 4; the nearest one (the small grid's y-min PML, D = 35)
    leaves w = 20 untouched for 8 steps.  Also: Dirichlet faces stay zero at full size."""
    from seismic_cpml_b200 import programs as P
    steps, w = 8, 20
    p = P.Params3DVisco(NSTEP=steps, **refcfg.TAU_CARCIONE_1993)   # defaults of 3D-visco :152-244, relaxation times of :402-413
    prog = P.Program3DVisco(p)
    prog.solver.run(1, steps)
    big = {name: prog.solver.get_field(f) for f, name in enumerate(FV) if name in ("vx", "vy", "vz", "sigmaxy", "sigmazz", "sigmayz_R")}
    sx_big, _ = prog.solver.get_seismograms()
    prog.solver.close()
    assert p.ISOURCE == 30 and p.JSOURCE == 161 and p.NZ // 2 == 110
    # small grid: source at (30, 45, 48), slab interface between planes 48 and 49 like 110 | 111
    c = refcfg.cfgv3d(nx=80, ny=220, nz=96, npml=10, nstep=steps)
    assert c["isource"] == 30 and c["jsource"] == 45
    o = O.run_3d_visco(**c, nproc=2, want_fields=True)
    for name, a in big.items():
        wb = a[110 - 1 - w:110 + w, 161 - 1 - w:161 + w, 0:30 + w]
        ws = o[name][48 - 1 - w:48 + w, 45 - 1 - w:45 + w, 0:30 + w]
        assert np.abs(ws).max() > 0, name
        assert np.array_equal(wb, ws), (name, np.abs(wb - ws).max())
    for a in (big["vx"], big["vy"], big["vz"]):           # Dirichlet, two planes per face (:1337-1371)
        assert not a[:, :, :1].any() and not a[:, :, -1:].any() and not a[:, :1, :].any() and not a[:, -1:, :].any()
        assert not a[:1].any() and not a[-1:].any()
    assert np.isfinite(sx_big).all()


def test_visco_vz_seismograms_extension():
    """cpml_get_seismograms_vz for the viscoelastic program (extension, quirk B7): the last sample equals the
    solver's own vz plane NZ/2 at the receivers, and the oracle's final vz field."""
    c = refcfg.cfgv3d()
    k = c["nz"] // 2
    with solver_visco(c, emulate_nproc=4) as s:
        s.run(1, c["nstep"])
        sz = s.get_seismograms_vz()
        plane = s.get_plane(2, k)
    o = O.run_3d_visco(**c, nproc=4, want_fields=True)
    assert np.abs(sz).max() > 0
    for r, (ix, iy) in enumerate(zip(c["ix_rec"], c["iy_rec"])):
        assert sz[r, -1] == plane[iy - 1, ix - 1] == o["vz"][k - 1, iy - 1, ix - 1]


def test_visco_sigmazz_isotropic_option_matches_the_oracle_variant():
    """cfg.sigmazz_isotropic = 1 (NOT the reference: the isotropic memory-variable term in sigmazz, quirk B14):
    bit-identical to the oracle run with the same flag, and different from the reference run."""
    c = refcfg.cfgv3d()
    with solver_visco(c, emulate_nproc=4, sigmazz_isotropic=True) as s:
        s.run(1, c["nstep"])
        sx, sy = s.get_seismograms()
        szz = s.get_field(5)
    o = O.run_3d_visco(**c, nproc=4, sigmazz_isotropic=True, want_fields=True)
    r = O.run_3d_visco(**c, nproc=4)
    assert np.array_equal(sx, o["sisvx"]) and np.array_equal(sy, o["sisvy"])
    assert np.array_equal(szz, o["sigmazz"])
    assert not np.array_equal(sx, r["sisvx"])
