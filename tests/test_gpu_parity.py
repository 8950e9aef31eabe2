"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same
inputs, against the committed golden vectors, and -- at full size -- through
size-independent properties.

Tolerance: BASELINE.json's north_star asks for relative L2 <= 1e-5 on seismograms and
the energy trace (TOL).  The kernels are built with -fmad=false and keep the
reference's operation order, so velocities and stresses are in fact bit-identical to
the oracle; that stronger statement is asserted wherever it holds (BITWISE).  Energies
are sums whose order differs (quirk B11): TOL_ENERGY.
"""
import os

import numpy as np
import pytest

import refcfg
from oracle import oracle as O
from seismic_cpml_b200 import lib as L
from seismic_cpml_b200 import programs as P

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-5            # north_star tolerance, relative L2
TOL_ENERGY = 1e-11    # what we actually hold the energy traces to
F3 = L.FIELDS_3D
F2 = L.FIELDS_2D


def solver3d(c, nslabs=1, slab_rank=0, **kw):
    s = L.Solver(ndim=3, order=2, nx=c["nx"], ny=c["ny"], nz=c["nz"], nstep=c["nstep"],
                 npoints_pml=c["npoints_pml"], nrec=len(c["ix_rec"]), isource=c["isource"],
                 jsource=c["jsource"], nslabs=nslabs, slab_rank=slab_rank, deltax=c["deltax"],
                 deltay=c["deltay"], deltaz=c["deltaz"], deltat=c["deltat"], lam=c["lam"], mu=c["mu"],
                 lambdaplustwomu=c["lambdaplustwomu"], rho=c["rho"], cp=3300.0, **kw)
    s.set_profiles(L.AXIS_X, c["prof_x"])
    s.set_profiles(L.AXIS_Y, c["prof_y"])
    s.set_profiles(L.AXIS_Z, c["prof_z"])
    s.set_source_series(c["force_x"], c["force_y"])
    s.set_receivers(c["ix_rec"], c["iy_rec"])
    return s


def solver2d(c):
    s = L.Solver(ndim=2, order=c["order"], nx=c["nx"], ny=c["ny"], nstep=c["nstep"],
                 npoints_pml=c["npoints_pml"], nrec=len(c["ix_rec"]), isource=c["isource"],
                 jsource=c["jsource"], deltax=c["deltax"], deltay=c["deltay"], deltat=c["deltat"], cp=3300.0)
    s.set_profiles(L.AXIS_X, c["prof_x"])
    s.set_profiles(L.AXIS_Y, c["prof_y"])
    s.set_material_2d(c["lam"], c["mu"], c["rho"])
    s.set_source_series(c["force_x"], c["force_y"])
    s.set_receivers(c["ix_rec"], c["iy_rec"])
    return s


def check_traces(got, ref, bitwise=True):
    for g, r in zip(got, ref):
        assert np.all(np.isfinite(g))
        assert refcfg.rel_l2(g, r) <= TOL
        if bitwise:
            assert np.array_equal(g, r), f"max abs diff {np.abs(g - r).max()}"


# ------------------------------------------------------------------ 3-D

@pytest.mark.parametrize("shape", [(37, 45, 40, 6), (64, 33, 48, 5), (129, 40, 32, 8), (33, 130, 36, 4)])
def test_3d_iso_matches_oracle(shape):
    """Fields, seismograms and energy after 150 steps on ragged grids (NX, NY not multiples
    of the tile; NX = 64 exercises pitch == NX)."""
    nx, ny, nz, npml = shape
    c = refcfg.cfg3d(nx=nx, ny=ny, nz=nz, npml=npml, nstep=150)
    o = O.run_3d_iso(**c, nproc=2, want_fields=True, want_planes=True)
    with solver3d(c) as s:
        s.run(1, c["nstep"])
        sx, sy = s.get_seismograms()
        check_traces((sx, sy), (o["sisvx"], o["sisvy"]))
        for f, name in enumerate(F3):
            got = s.get_field(f)
            assert np.array_equal(got, o[name]), (name, np.abs(got - o[name]).max())
        tot, ek, ep = s.get_energy()
        assert refcfg.rel_l2(tot, o["total_energy"]) <= TOL_ENERGY
        assert np.array_equal(s.get_plane(0, nz // 2), o["plane_vx"])
        assert s.get_maxnorm() == pytest.approx(o["vnorm"], rel=1e-15)
    assert np.abs(o["sisvx"]).max() > 1e-3       # the wave did reach the receivers


def test_3d_iso_golden_vectors():
    g = np.load(os.path.join(GOLD, "cpml3d_iso_small.npz"))
    c = refcfg.cfg3d()
    with solver3d(c) as s:
        s.run(1, c["nstep"])
        check_traces(s.get_seismograms(), (g["sisvx"], g["sisvy"]))
        assert refcfg.rel_l2(s.get_energy()[0], g["total_energy"]) <= TOL_ENERGY
        assert np.array_equal(s.get_plane(1, c["nz"] // 2), g["plane_vy"])


def test_3d_iso_kmax_pml():
    """K_MAX_PML != 1 exercises the value/K path of the recursion (:849-851)."""
    g = np.load(os.path.join(GOLD, "cpml3d_iso_kmax3.npz"))
    c = refcfg.cfg3d(nx=30, ny=34, nz=32, nstep=120, npml=5, k_max=3.0)
    with solver3d(c) as s:
        s.run(1, c["nstep"])
        check_traces(s.get_seismograms(), (g["sisvx"], g["sisvy"]))
        assert refcfg.rel_l2(s.get_energy()[0], g["total_energy"]) <= TOL_ENERGY


def test_3d_energy_bug_flag_and_reset():
    c = refcfg.cfg3d(nstep=60)
    a = O.run_3d_iso(**c, nproc=2, energy_bug_compat=False)
    with solver3d(c, energy_bug_compat=False) as s:
        s.run(1, 60)
        e1 = s.get_energy()[0]
        assert refcfg.rel_l2(e1, a["total_energy"]) <= TOL_ENERGY
        sx1, _ = s.get_seismograms()
        s.reset()                                  # == the zeroing of :720-756
        assert s.get_maxnorm() == 0.0 and not s.get_seismograms()[0].any()
        s.run(1, 60)
        assert np.array_equal(s.get_seismograms()[0], sx1) and np.array_equal(s.get_energy()[0], e1)


def test_3d_partial_runs_and_zero_tail():
    """cpml_run in pieces == one run; traces beyond the last executed step stay zero
    (the reference writes partial seismogram files, :1234)."""
    c = refcfg.cfg3d(nstep=90)
    with solver3d(c) as s1, solver3d(c) as s2:
        s1.run(1, 60)
        for a, b in ((1, 5), (6, 6), (7, 60)):
            s2.run(a, b)
        for x, y in zip(s1.get_seismograms(), s2.get_seismograms()):
            assert np.array_equal(x, y) and not x[:, 60:].any() and x[:, :60].any()
        assert np.array_equal(s1.get_energy()[0], s2.get_energy()[0])


def exchange_on_host(slabs, phase):
    """The plane exchange of 3D-iso :811-823 (phase 'v') / :951-963 (phase 's') between slab
    handles owned by one process (cpml_copy_plane)."""
    nzl = slabs[0].nzl
    moves = {"v": [(0, "left"), (1, "left"), (2, "right")], "s": [(5, "left"), (8, "right"), (7, "right")]}[phase]
    for r in range(len(slabs) - 1):
        lo, hi = slabs[r], slabs[r + 1]
        for f, direction in moves:
            if direction == "left":    # plane 1 of the upper slab -> halo NZ_LOCAL+1 of the lower
                lo.copy_plane_from(nzl + 1, hi, 1, f)
            else:                      # plane NZ_LOCAL of the lower slab -> halo 0 of the upper
                hi.copy_plane_from(0, lo, nzl, f)
    for s in slabs:
        s.synchronize()


@pytest.mark.parametrize("nslabs", [2, 4])
def test_3d_slab_decomposition_matches_single_slab(nslabs):
    """N slab handles + plane exchange (the reference's MPI layout) == one whole-grid handle,
    bit for bit; the slab energies add up to the total (MPI_REDUCE, :1179)."""
    c = refcfg.cfg3d(nx=40, ny=37, nz=48, npml=5, nstep=80)
    with solver3d(c) as whole:
        whole.run(1, c["nstep"])
        ref_sx, ref_sy = whole.get_seismograms()
        ref_e = whole.get_energy()[0]
        ref_fields = [whole.get_field(f) for f in range(9)]
    slabs = [solver3d(c, nslabs=nslabs, slab_rank=r) for r in range(nslabs)]
    try:
        with pytest.raises(L.CpmlError):
            slabs[0].run(1, 2)                     # cpml_run is for whole grids
        for it in range(1, c["nstep"] + 1):
            exchange_on_host(slabs, "v")
            for s in slabs:
                s.step_stress(it)
            for s in slabs:
                s.synchronize()
            exchange_on_host(slabs, "s")
            for s in slabs:
                s.step_velocity(it)
                s.step_finish(it)
            for s in slabs:
                s.synchronize()
        owner = nslabs // 2 - 1                    # rank_cut_plane, :346
        sx, sy = slabs[owner].get_seismograms()
        assert np.array_equal(sx, ref_sx) and np.array_equal(sy, ref_sy)
        for r, s in enumerate(slabs):
            if r != owner:
                assert not s.get_seismograms()[0].any()
        e = sum(s.get_energy()[0] for s in slabs)
        assert refcfg.rel_l2(e, ref_e) <= TOL_ENERGY
        for f in range(9):
            got = np.concatenate([s.get_field(f) for s in slabs], axis=0)
            assert np.array_equal(got, ref_fields[f]), F3[f]
    finally:
        for s in slabs:
            s.close()


@pytest.mark.parametrize("nslabs", [2, 4])
def test_3d_slabs_with_peer_stores_match_single_slab(nslabs):
    """The slab handles attached to each other (cpml_p2p_attach_local): the update kernels
    write the six boundary planes of :811-823 / :951-963 straight into the neighbour's halo
    planes and device-side flags order the half steps -- no exchange call, no host
    synchronisation inside the loop.  Result == whole grid, bit for bit."""
    c = refcfg.cfg3d(nx=40, ny=37, nz=48, npml=5, nstep=80)
    with solver3d(c) as whole:
        whole.run(1, c["nstep"])
        ref_sx, ref_sy = whole.get_seismograms()
        ref_e = whole.get_energy()[0]
        ref_fields = [whole.get_field(f) for f in range(9)]
    slabs = [solver3d(c, nslabs=nslabs, slab_rank=r) for r in range(nslabs)]
    try:
        with pytest.raises(L.CpmlError):
            slabs[0].p2p_attach_local(0, slabs[1])          # rank 0 has no lower neighbour
        for r in range(nslabs - 1):
            slabs[r].p2p_attach_local(1, slabs[r + 1])
            slabs[r + 1].p2p_attach_local(0, slabs[r])
        assert slabs[0].launch_info()["peer_sides"] == 2 and slabs[-1].launch_info()["peer_sides"] == 1
        for it in range(1, c["nstep"] + 1):
            for s in slabs:
                s.step_stress(it)
            for s in slabs:
                s.step_velocity(it)
                s.step_finish(it)
        for s in slabs:
            s.synchronize()
        owner = nslabs // 2 - 1
        sx, sy = slabs[owner].get_seismograms()
        assert np.array_equal(sx, ref_sx) and np.array_equal(sy, ref_sy)
        e = sum(s.get_energy()[0] for s in slabs)
        assert refcfg.rel_l2(e, ref_e) <= TOL_ENERGY
        for f in range(9):
            got = np.concatenate([s.get_field(f) for s in slabs], axis=0)
            assert np.array_equal(got, ref_fields[f]), F3[f]
    finally:
        for s in slabs:
            s.close()


TMA_TILES = [("tma", t) for t in [(64, 4), (64, 8), (104, 7), (104, 8), (128, 4), (128, 6), (128, 8)]] + \
            [("ws", t) for t in [(64, 4), (64, 8), (104, 7), (104, 8), (128, 6), (128, 7)]]


@pytest.mark.parametrize("kernel,tile", TMA_TILES)
@pytest.mark.parametrize("stages", [1, 3])
def test_3d_tma_tiles_and_ring_depths(kernel, tile, stages, monkeypatch):
    """Every TMA box shape / shared-memory ring depth of both TMA-staged kernel families (CPML_KERNEL=tma: thread 0
    issues the loads; ws, the default: a producer warp does) gives the same bits (ragged grid: NX, NY not multiples
    of any tile; several z chunks, so the boundary-chunks-first item order of the ws kernels is exercised)."""
    monkeypatch.setenv("CPML_KERNEL", kernel)
    monkeypatch.setenv("CPML_TX", str(tile[0]))
    monkeypatch.setenv("CPML_TY", str(tile[1]))
    monkeypatch.setenv("CPML_TY_STRESS", str(tile[1]))
    monkeypatch.setenv("CPML_STAGES", str(stages))
    monkeypatch.setenv("CPML_ZCHUNKS", "3")
    c = refcfg.cfg3d(nx=70, ny=45, nz=40, npml=6, nstep=60)
    o = O.run_3d_iso(**c, nproc=2, want_fields=True)
    with solver3d(c) as s:
        info = s.launch_info()
        assert info["tma"] == (2 if kernel == "ws" else 1)
        assert (info["tile_x"], info["tile_y"]) == tile and info["z_chunks"] == 3
        s.run(1, c["nstep"])
        check_traces(s.get_seismograms(), (o["sisvx"], o["sisvy"]))
        for f, name in enumerate(F3):
            assert np.array_equal(s.get_field(f), o[name]), name
        assert refcfg.rel_l2(s.get_energy()[0], o["total_energy"]) <= TOL_ENERGY


@pytest.mark.parametrize("zchunks,split", [(4, 2), (3, 4), (5, 3), (4, 1)])
@pytest.mark.parametrize("sched", ["queue", "static"])
def test_3d_work_queue_and_finer_tail(zchunks, split, sched, monkeypatch):
    """The persistent kernels' work distribution must not show in the results: items claimed from the queue or in
    static shares, the last coarse items split into 2..4 finer ones (CPML_TAIL_SPLIT; the default only splits chunks
    of >= 16 planes, which the small test grids never have), ragged chunk lengths -- bit-identical to the oracle
    either way, energies to 1e-11 (the partial sums are per work item)."""
    c = refcfg.cfg3d(nx=37, ny=45, nz=42, npml=6, nstep=90)
    o = O.run_3d_iso(**c, nproc=2, want_fields=True)
    monkeypatch.setenv("CPML_ZCHUNKS", str(zchunks))
    monkeypatch.setenv("CPML_ZCHUNKS_STRESS", str(zchunks + 1))
    monkeypatch.setenv("CPML_TAIL_SPLIT", str(split))
    if sched == "static":
        monkeypatch.setenv("CPML_SCHED", "static")
    with solver3d(c) as s:
        assert s.launch_info()["tma"] == 2
        s.run(1, c["nstep"])
        sx, sy = s.get_seismograms()
        assert np.abs(o["sisvx"]).max() > 1e-4
        assert np.array_equal(sx, o["sisvx"]) and np.array_equal(sy, o["sisvy"])
        for f, name in enumerate(F3):
            assert np.array_equal(s.get_field(f), o[name]), name
        assert refcfg.rel_l2(s.get_energy()[0], o["total_energy"]) <= 1e-11


def test_3d_register_kernels_still_match(monkeypatch):
    """CPML_KERNEL=reg: the register-marching kernels kept for A/B measurements."""
    monkeypatch.setenv("CPML_KERNEL", "reg")
    c = refcfg.cfg3d(nx=37, ny=45, nz=40, npml=6, nstep=60)
    o = O.run_3d_iso(**c, nproc=2, want_fields=True)
    with solver3d(c) as s:
        assert s.launch_info()["tma"] == 0 and s.kernel_names() == ("k_stress3d", "k_velocity3d")
        s.run(1, c["nstep"])
        check_traces(s.get_seismograms(), (o["sisvx"], o["sisvy"]))
        for f, name in enumerate(F3):
            assert np.array_equal(s.get_field(f), o[name]), name


def test_3d_default_grid_full_size_vs_timed_oracle():
    """BASELINE config 3 (101 x 641 x 640, 41.4 M points) for 12 steps against the OpenMP
    build of the oracle (FMA-contracted, so not bit-identical: TOL applies), plus
    properties: Dirichlet faces are zero, fields are finite, energy is positive."""
    p = P.Params3DIso(NSTEP=12)
    prog = P.Program3DIso(p)
    res = prog.run()
    s = prog.s
    o = O.run_3d_iso(nx=p.NX, ny=p.NY, nz=p.NZ, nproc=2, deltax=p.DELTAX, deltay=p.DELTAY, deltaz=p.DELTAZ,
                     deltat=p.DELTAT, lam=p.lam, mu=p.mu, lambdaplustwomu=p.lambdaplustwomu, rho=p.rho,
                     nstep=p.NSTEP, npoints_pml=p.NPOINTS_PML, isource=p.ISOURCE, jsource=p.JSOURCE,
                     prof_x=s.prof_x, prof_y=s.prof_y, prof_z=s.prof_z, force_x=s.force_x, force_y=s.force_y,
                     ix_rec=s.ix_rec, iy_rec=s.iy_rec, want_planes=True, kind="timed")
    assert refcfg.rel_l2(res["total_energy"], o["total_energy"]) <= TOL
    pvx = prog.solver.get_plane(0, p.NZ // 2)
    assert refcfg.rel_l2(pvx, o["plane_vx"]) <= TOL
    assert pvx[p.JSOURCE - 1, p.ISOURCE - 1] != 0.0
    assert not pvx[0].any() and not pvx[-1].any() and not pvx[:, 0].any() and not pvx[:, -1].any()
    assert not prog.solver.get_plane(2, 1).any() and not prog.solver.get_plane(2, p.NZ).any()
    assert np.all(res["total_energy"][1:] > 0)
    assert prog.solver.get_maxnorm() == pytest.approx(o["vnorm"], rel=1e-9)
    prog.solver.close()


# ------------------------------------------------------------------ 2-D

@pytest.mark.parametrize("kernel", ["ws", "pair"])
@pytest.mark.parametrize("order", [2, 4])
def test_2d_layered_matches_oracle_and_golden(order, kernel, monkeypatch):
    """Both 2-D kernel families: the TMA-staged y-marching kernels (ws, the default) and the pair kernels."""
    monkeypatch.setenv("CPML_2D_KERNEL", kernel)
    c = refcfg.cfg2d(order, nx=83, ny=131, nstep=600, npml=8, material="layered", ydeb=600.0, yfin=200.0)
    g = np.load(os.path.join(GOLD, f"cpml2d_layered_order{order}.npz"))
    o = O.run_2d(**c, want_fields=True)
    with solver2d(c) as s:
        s.run(1, c["nstep"])
        check_traces(s.get_seismograms(), (g["sisvx"], g["sisvy"]))
        for f, name in enumerate(F2):
            assert np.array_equal(s.get_field(f), o[name]), name
        _, ek, ep = s.get_energy()
        assert refcfg.rel_l2(ek, g["energy_kinetic"]) <= TOL_ENERGY
        assert refcfg.rel_l2(ep, g["energy_potential"]) <= TOL_ENERGY
        assert s.get_maxnorm() == pytest.approx(o["velocnorm"], rel=1e-15)


@pytest.mark.parametrize("order", [2, 4])
def test_2d_shipped_configuration_full_run(order):
    """BASELINE config 1 (second order as shipped, all 2000 steps) and its fourth-order
    twin (4000 steps) against the committed golden vectors, through the driver mirror."""
    g = np.load(os.path.join(GOLD, "cpml2d_second_default.npz" if order == 2 else "cpml2d_fourth_default.npz"))
    prog = P.Program2DIso(P.Params2DIso(order=order))
    res = prog.run()
    check_traces((res["sisvx"], res["sisvy"]), (g["sisvx"], g["sisvy"]))
    assert refcfg.rel_l2(res["energy_kinetic"], g["energy_kinetic"]) <= TOL_ENERGY
    assert refcfg.rel_l2(res["energy_potential"], g["energy_potential"]) <= TOL_ENERGY
    # the driver displayed at it = 5 and every IT_DISPLAY steps (:716), never unstable
    its = [d[0] for d in res["display_log"]]
    assert its[0] == 5 and its[1] == prog.p.IT_DISPLAY and its[-1] == prog.p.NSTEP
    e = res["energy_kinetic"] + res["energy_potential"]
    assert e[-1] < 1e-6 * e.max()
    prog.solver.close()


@pytest.mark.parametrize("order", [2, 4])
@pytest.mark.parametrize("chunks,minb", [(1, 3), (3, 3), (5, 2)])
def test_2d_ws_strips_and_chunks(order, chunks, minb, monkeypatch):
    """The y-marching kernels on a grid with three x strips (the last one ragged), NY not a multiple of the row block,
    heterogeneous medium, K_MAX_PML = 2 shells, for 1 / 3 / 5 y chunks (the chunk seams fall inside the wavefield):
    fields bitwise, energies to summation order."""
    monkeypatch.setenv("CPML_2D_CHUNKS", str(chunks))
    monkeypatch.setenv("CPML_2D_WS_MINB", str(minb))
    c = refcfg.cfg2d(order, nx=150, ny=203, nstep=500, npml=8, material="layered", ydeb=900.0, yfin=300.0, k_max=2.0)
    o = O.run_2d(**c, want_fields=True)
    with solver2d(c) as s:
        s.run(1, c["nstep"])
        check_traces(s.get_seismograms(), (o["sisvx"], o["sisvy"]))
        for f, name in enumerate(F2):
            assert np.array_equal(s.get_field(f), o[name]), name
        _, ek, ep = s.get_energy()
        assert refcfg.rel_l2(ek, o["energy_kinetic"]) <= TOL_ENERGY and refcfg.rel_l2(ep, o["energy_potential"]) <= TOL_ENERGY
    assert np.abs(o["sisvx"]).max() > 1e-3


def test_2d_kmax_quirk_b3_fourth_order():
    """With K_MAX_PML != 1 the fourth-order program's K_y(j) (2D-4th :596) matters."""
    c = refcfg.cfg2d(4, nx=70, ny=90, nstep=300, npml=8, k_max=2.5, ydeb=500.0, yfin=200.0)
    o = O.run_2d(**c)
    with solver2d(c) as s:
        s.run(1, c["nstep"])
        check_traces(s.get_seismograms(), (o["sisvx"], o["sisvy"]))


def test_2d_4096_fourth_order_properties():
    """BASELINE config 2 (4096 x 4096, fourth order), 40 steps: a 256 x 256 window around
    the source must equal the oracle run on a smaller grid that shares that window
    (finite propagation speed: nothing outside the window has reached it yet)."""
    n, steps = 4096, 40
    p = P.Params2DIso(order=4, NX=n, NY=n, NSTEP=steps, ISOURCE=n - 100, JSOURCE=n - 120)
    prog = P.Program2DIso(p)
    prog.run()
    big_vx = prog.solver.get_field(0)
    prog.solver.close()
    # small grid: same spacing/time step/source law, source at the same distance from the
    # upper-right corner so that the PML/edges seen by the window are identical
    m = 512
    q = P.Params2DIso(order=4, NX=m, NY=m, NSTEP=steps, ISOURCE=m - (n - p.ISOURCE), JSOURCE=m - (n - p.JSOURCE))
    s = P.setup_2d(q)
    o = O.run_2d(order=4, nx=m, ny=m, deltax=q.DELTAX, deltay=q.DELTAY, deltat=q.DELTAT, nstep=steps,
                 npoints_pml=q.NPOINTS_PML, isource=q.ISOURCE, jsource=q.JSOURCE, lam=s.material[0],
                 mu=s.material[1], rho=s.material[2], prof_x=s.prof_x, prof_y=s.prof_y, force_x=s.force_x,
                 force_y=s.force_y, ix_rec=s.ix_rec, iy_rec=s.iy_rec, want_fields=True)
    w = 200
    a = big_vx[n - w:, n - w:]
    b = o["vx"][m - w:, m - w:]
    assert np.abs(b).max() > 0
    assert np.array_equal(a, b)
    # far from the source nothing has moved yet
    assert not big_vx[: n // 2, : n // 2].any()


def test_3d_vz_seismograms_extension():
    """Vz seismograms (cpml_get_seismograms_vz): an extension -- the reference records Vx and Vy only although
    its plot script reads Vz files (quirk B7).  sisvz(it, irec) must be vz(ix_rec, iy_rec, NZ/2) after step it:
    checked against the oracle's final vz field for several run lengths, and against the solver's own plane."""
    c0 = refcfg.cfg3d(nx=37, ny=45, nz=40, nstep=120, npml=6)
    k = c0["nz"] // 2
    with solver3d(c0) as s:
        s.run(1, c0["nstep"])
        sz = s.get_seismograms_vz()
        sx, _ = s.get_seismograms()
        plane = s.get_plane(2, k)
    assert sz.shape == sx.shape and np.abs(sz).max() > 0
    for r, (ix, iy) in enumerate(zip(c0["ix_rec"], c0["iy_rec"])):
        assert sz[r, -1] == plane[iy - 1, ix - 1]
    for n in (30, 75, 120):
        c = refcfg.cfg3d(nx=37, ny=45, nz=40, nstep=n, npml=6)
        o = O.run_3d_iso(**c, nproc=2, want_fields=True)
        for r, (ix, iy) in enumerate(zip(c["ix_rec"], c["iy_rec"])):
            assert sz[r, n - 1] == o["vz"][k - 1, iy - 1, ix - 1], (n, r)
    with solver2d(refcfg.cfg2d(2, nx=60, ny=70, nstep=10, npml=6)) as s2:
        with pytest.raises(L.CpmlError):
            s2.get_seismograms_vz()


# ------------------------------------------------------------------ sizes the bench numbers are quoted on

def _iso3d_workload(nx, ny, nz, nstep, **kw):
    """The 3-D isotropic program's own parameter block (:124-218) on an nx x ny x nz grid, oracle-side set-up."""
    from oracle import workloads as WL
    return WL.iso3d(nx, ny, nz, nstep, **kw)


@pytest.mark.parametrize("tx", [64, 104, 128])
def test_3d_interior_x_tiles(tx, monkeypatch):
    """NX = 300: several x tiles per row for every tile width -- tiles that touch the low x shell, interior
    tiles (no x shell: the staged x-shell rows are skipped) and tiles that touch the high shell; bitwise."""
    monkeypatch.setenv("CPML_TX", str(tx))
    c = refcfg.cfg3d(nx=300, ny=40, nz=32, npml=6, nstep=70)
    o = O.run_3d_iso(**c, nproc=2, want_fields=True)
    with solver3d(c) as s:
        info = s.launch_info()
        assert info["tile_x"] == tx and (300 + tx - 1) // tx >= 3
        s.run(1, c["nstep"])
        check_traces(s.get_seismograms(), (o["sisvx"], o["sisvy"]))
        for f, name in enumerate(F3):
            assert np.array_equal(s.get_field(f), o[name]), name
        assert refcfg.rel_l2(s.get_energy()[0], o["total_energy"]) <= TOL_ENERGY


def test_3d_cfg4_slab_window_vs_oracle():
    """BASELINE config 4's per-GPU grid (1024 x 1024 x 128, default tiles: several x tiles per row, 512-thread
    CTAs), 30 steps: a window around the source -- it includes the x-max C-PML shell, 21 cells from the source --
    must equal, bit for bit, the oracle run on an 80 x 126 x 84 grid that shares that window (finite numerical
    propagation speed: one cell per step, so nothing outside the window has reached it and the small grid's
    other shells are still at rest)."""
    nx, ny, nz, steps = 1024, 1024, 128, 30
    big = _iso3d_workload(nx, ny, nz, steps)
    mx, my, mz = 80, 126, 84
    small = _iso3d_workload(mx, my, mz, steps)
    isb, jsb, ksb = big["isource"], big["jsource"], nz // 2
    iss, jss, kss = small["isource"], small["jsource"], mz // 2
    assert nx - isb == mx - iss == 21                       # same distance to the x-max edge
    assert min(iss - 1, jss - 1, my - jss, kss - 1, mz - kss) >= steps + 10 + 1
    o = O.run_3d_iso(**small, nproc=2, want_fields=True)
    with solver3d(big) as s:
        info = s.launch_info()
        assert (nx + info["tile_x"] - 1) // info["tile_x"] >= 2
        s.run(1, steps)
        x0, y0 = isb - iss, jsb - jss                       # big index = small index + offset
        for f in (0, 1, 2, 3, 5, 6, 7, 8):
            name = F3[f]
            for dk in (-25, -7, -1, 0, 1, 6, 20):
                got = s.get_plane(f, ksb + dk)[y0:y0 + my, x0:x0 + mx]
                ref = o[name][kss + dk - 1]
                assert np.array_equal(got, ref), (name, dk, np.abs(got - ref).max())
            assert np.abs(o[name][kss - 1]).max() > 0
        # the wave has not left the window: planes far from the source are still at rest
        assert not s.get_plane(0, 5).any() and not s.get_plane(2, nz - 4).any()
        pv = s.get_plane(0, ksb)
        assert not pv[: y0 - 2].any() and not pv[:, : x0 - 2].any()
        assert refcfg.rel_l2(s.get_energy()[0], o["total_energy"]) <= TOL_ENERGY


@pytest.mark.slow
def test_3d_default_xy_geometry_full_run_seismograms():
    """The shipped x-y geometry (101 x 641, source (80,428), receivers (70,231) and (80,31)) with NZ = 64, ALL 2500
    time steps: seismograms at the real receivers through first arrival, PML contact and decay, and the energy
    trace, against the OpenMP build of the oracle (FMA-contracted: TOL applies) -- north_star's <= 1e-5."""
    c = _iso3d_workload(101, 641, 64, 2500)
    assert (list(c["ix_rec"]), list(c["iy_rec"])) == ([70, 80], [231, 31])
    o = O.run_3d_iso(**c, nproc=2, kind="timed")
    with solver3d(c) as s:
        s.run(1, c["nstep"])
        sx, sy = s.get_seismograms()
        e = s.get_energy()[0]
    for g, r in ((sx, o["sisvx"]), (sy, o["sisvy"])):
        for rec in range(2):
            assert np.abs(r[rec]).max() > 1e-3
            assert refcfg.rel_l2(g[rec], r[rec]) <= TOL, (rec, refcfg.rel_l2(g[rec], r[rec]))
    assert refcfg.rel_l2(e, o["total_energy"]) <= TOL
    assert e[-1] < 1e-3 * e.max()                           # the shells did absorb the wavefield


@pytest.mark.slow
def test_3d_default_grid_450_steps_vs_timed_oracle():
    """BASELINE config 3 at full size (101 x 641 x 640) for 450 steps: past the first arrival at the nearest
    receiver (197 cells from the source), with the wavefront inside the x shells; seismograms, the energy
    trace and the receiver plane against the OpenMP build of the oracle (TOL)."""
    c = _iso3d_workload(101, 641, 640, 450)
    o = O.run_3d_iso(**c, nproc=2, want_planes=True, kind="timed")
    with solver3d(c) as s:
        s.run(1, c["nstep"])
        sx, sy = s.get_seismograms()
        e = s.get_energy()[0]
        pvx, pvy = s.get_plane(0, 320), s.get_plane(1, 320)
    assert np.abs(o["sisvx"][0]).max() > 1e-6               # receiver 1 has seen the wave
    assert refcfg.rel_l2(sx[0], o["sisvx"][0]) <= TOL and refcfg.rel_l2(sy[0], o["sisvy"][0]) <= TOL
    assert refcfg.rel_l2(e, o["total_energy"]) <= TOL
    assert refcfg.rel_l2(pvx, o["plane_vx"]) <= TOL and refcfg.rel_l2(pvy, o["plane_vy"]) <= TOL


def test_per_step_source_upload_and_result_fetch():
    """The per-step host API a driver's `do it` body uses (bench.py's e2e leg): cpml_set_source_step uploads the source
    term of step `it` (one strided copy from pinned memory), cpml_fetch_step queues the read of that step's kinetic /
    potential energy and first-receiver sample (one 32-byte copy of what k_post3d left side by side).  Same bits as
    the whole-series path, 3-D and 2-D."""
    c = refcfg.cfg3d(nx=37, ny=45, nz=40, npml=6, nstep=90)
    with solver3d(c) as whole:
        whole.run(1, c["nstep"])
        wx, wy = whole.get_seismograms()
        we = whole.get_energy()
    s = L.Solver(ndim=3, order=2, nx=c["nx"], ny=c["ny"], nz=c["nz"], nstep=c["nstep"], npoints_pml=c["npoints_pml"],
                 nrec=len(c["ix_rec"]), isource=c["isource"], jsource=c["jsource"], deltax=c["deltax"], deltay=c["deltay"],
                 deltaz=c["deltaz"], deltat=c["deltat"], lam=c["lam"], mu=c["mu"], lambdaplustwomu=c["lambdaplustwomu"],
                 rho=c["rho"], cp=3300.0)
    with s:
        s.set_profiles(L.AXIS_X, c["prof_x"]); s.set_profiles(L.AXIS_Y, c["prof_y"]); s.set_profiles(L.AXIS_Z, c["prof_z"])
        s.set_receivers(c["ix_rec"], c["iy_rec"])
        for it in range(1, c["nstep"] + 1):
            s.set_source_step(it, c["force_x"][it - 1], c["force_y"][it - 1])
            s.step_stress(it); s.step_velocity(it); s.step_finish(it)
            s.fetch_step(it)
        sx, sy = s.get_seismograms()
        e = s.get_energy()
        assert np.abs(wx).max() > 1e-5
        assert np.array_equal(sx, wx) and np.array_equal(sy, wy)
        assert np.array_equal(e[1], we[1]) and np.array_equal(e[2], we[2])
        for it in (1, 17, c["nstep"]):
            f = s.get_fetched_step(it)
            assert f[0] == e[1][it - 1] and f[1] == e[2][it - 1] and f[2] == sx[0, it - 1] and f[3] == sy[0, it - 1], (it, f)

    c2 = refcfg.cfg2d(4, nx=90, ny=70, nstep=50, npml=6)
    o2 = O.run_2d(**c2)
    p2 = solver2d(c2)
    with p2:
        for it in range(1, c2["nstep"] + 1):
            p2.step_stress(it); p2.step_velocity(it); p2.step_finish(it)
            p2.fetch_step(it)
        sx2, sy2 = p2.get_seismograms()
        assert np.array_equal(sx2, o2["sisvx"])
        f = p2.get_fetched_step(33)
        assert f[2] == sx2[0, 32] and f[3] == sy2[0, 32]


def test_asynchronous_snapshot_planes():
    """cpml_snapshot_begin / _end: the plane is the one of the step at which the pull was STARTED, however far the loop
    has moved on when it is collected (device-side copy first); slots are independent; misuse is refused."""
    c = refcfg.cfg3d(nx=37, ny=45, nz=40, nstep=90, npml=6)
    k = c["nz"] // 2
    with solver3d(c) as s:
        s.run(1, 50)
        ref_vx, ref_vy = s.get_plane(0, k), s.get_plane(1, k)
        s.snapshot_begin(0, 0, k)
        s.snapshot_begin(1, 1, k)
        with pytest.raises(L.CpmlError):
            s.snapshot_begin(0, 0, k)                  # slot 0 has not been collected yet
        for it in range(51, 91):                       # the loop goes on while the planes travel
            s.step_stress(it); s.step_velocity(it); s.step_finish(it)
        a = s.snapshot_end(0)
        b = s.snapshot_end(1, copy=False).copy()
        assert np.array_equal(a, ref_vx) and np.array_equal(b, ref_vy) and np.abs(a).max() > 0
        assert not np.array_equal(s.get_plane(0, k), ref_vx)
        with pytest.raises(L.CpmlError):
            s.snapshot_end(0)
        s.snapshot_begin(0, 2, k)                      # the slot is free again
        assert np.array_equal(s.snapshot_end(0), s.get_plane(2, k))
    c2 = refcfg.cfg2d(4, nx=70, ny=90, nstep=40, npml=8, ydeb=500.0, yfin=200.0)
    with solver2d(c2) as s2:
        s2.run(1, 30)
        ref = s2.get_plane(1)
        s2.snapshot_begin(0, 1)
        s2.run(31, 40)
        assert np.array_equal(s2.snapshot_end(0), ref)
