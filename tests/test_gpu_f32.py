"""Single-precision mode of the 3-D isotropic solver (cpml_config.precision = 1).

The reference endorses a single-precision build ("significantly faster", seismic_CPML_3D_isotropic_MPI_OpenMP.f90:114-116:
declare everything `real`).  Two statements are tested:
  * the CUDA kernels in FP32 are BIT-IDENTICAL to an independent IEEE-single restatement of the loop (the numpy
    restatement run with dtype=float32: every operation rounded to single, no FMA, same operation order) -- so the FP32
    path has the same kind of oracle as the FP64 one;
  * FP32 against the FP64 oracle: the stated tolerance.  Measured on B200 (profiles/r02_f32_tolerance.txt) the
    seismograms differ by ~1e-6 relative L2 over a few hundred steps; TOL_F32 = 1e-4 is asserted, north_star's 1e-5
    is reported.
"""
import numpy as np
import pytest

import refcfg
from oracle import oracle as O
from oracle.np_restatement import run_3d_iso_np
from seismic_cpml_b200 import lib as L

pytestmark = pytest.mark.gpu
TOL_F32 = 1e-4
F3 = L.FIELDS_3D


def solver3d(c, precision=1, **kw):
    s = L.Solver(ndim=3, order=2, nx=c["nx"], ny=c["ny"], nz=c["nz"], nstep=c["nstep"], npoints_pml=c["npoints_pml"],
                 nrec=len(c["ix_rec"]), isource=c["isource"], jsource=c["jsource"], deltax=c["deltax"], deltay=c["deltay"],
                 deltaz=c["deltaz"], deltat=c["deltat"], lam=c["lam"], mu=c["mu"], lambdaplustwomu=c["lambdaplustwomu"],
                 rho=c["rho"], cp=3300.0, precision=precision, **kw)
    s.set_profiles(L.AXIS_X, c["prof_x"])
    s.set_profiles(L.AXIS_Y, c["prof_y"])
    s.set_profiles(L.AXIS_Z, c["prof_z"])
    s.set_source_series(c["force_x"], c["force_y"])
    s.set_receivers(c["ix_rec"], c["iy_rec"])
    return s


@pytest.mark.parametrize("shape,k_max", [((37, 45, 40, 6), 1.0), ((70, 33, 36, 5), 1.0), ((130, 30, 24, 6), 1.0), ((30, 34, 32, 5), 3.0)])
def test_f32_kernels_equal_the_ieee_single_restatement(shape, k_max):
    """Fields and seismograms of the FP32 kernels == the numpy restatement in float32, bit for bit (ragged grids, one
    and several x tiles, K_MAX_PML != 1 exercises the single-precision division); energy (summed in double) to 1e-10."""
    nx, ny, nz, npml = shape
    c = refcfg.cfg3d(nx=nx, ny=ny, nz=nz, npml=npml, nstep=100, k_max=k_max)
    o = run_3d_iso_np(**c, dtype=np.float32)
    with solver3d(c) as s:
        assert s.launch_info()["tma"] == 2
        s.run(1, c["nstep"])
        sx, sy = s.get_seismograms()
        assert np.abs(o["sisvx"]).max() > 1e-4
        assert np.array_equal(sx, o["sisvx"].astype(np.float64)) and np.array_equal(sy, o["sisvy"].astype(np.float64))
        for f, name in enumerate(F3):
            got = s.get_field(f)
            assert np.array_equal(got, o[name].astype(np.float64)), (name, np.abs(got - o[name]).max())
        assert refcfg.rel_l2(s.get_energy()[0], o["total_energy"]) <= 1e-10
        k = nz // 2
        assert np.array_equal(s.get_plane(1, k), o["vy"][k - 1].astype(np.float64))
        s.snapshot_begin(0, 2, k)
        assert np.array_equal(s.snapshot_end(0), o["vz"][k - 1].astype(np.float64))
        vn = np.sqrt(o["vx"].astype(np.float64) ** 2 + o["vy"].astype(np.float64) ** 2 + o["vz"].astype(np.float64) ** 2).max()
        assert s.get_maxnorm() == pytest.approx(vn, rel=1e-14)


def test_f32_against_the_double_precision_oracle_tolerance():
    """The tolerance study: 64 x 120 x 64, 500 steps (the wave crosses the receivers and enters every shell)."""
    c = refcfg.cfg3d(nx=64, ny=120, nz=64, npml=8, nstep=500)
    o = O.run_3d_iso(**c, nproc=2, kind="timed")
    with solver3d(c) as s:
        s.run(1, c["nstep"])
        sx, sy = s.get_seismograms()
        e = s.get_energy()[0]
    with solver3d(c, precision=0) as d:
        d.run(1, c["nstep"])
        dx, _ = d.get_seismograms()
    errs = [refcfg.rel_l2(sx[r], o["sisvx"][r]) for r in range(2)] + [refcfg.rel_l2(sy[r], o["sisvy"][r]) for r in range(2)]
    ee = refcfg.rel_l2(e, o["total_energy"])
    print(f"FP32 vs FP64 oracle: seismogram rel-L2 {max(errs):.3e} (per trace {['%.2e' % x for x in errs]}), energy rel-L2 {ee:.3e}; "
          f"FP64 kernels vs the same oracle {refcfg.rel_l2(dx[0], o['sisvx'][0]):.3e}")
    assert np.abs(o["sisvx"]).max() > 1e-3
    assert max(errs) <= TOL_F32 and ee <= TOL_F32


def test_f32_is_refused_where_it_is_not_implemented():
    c = refcfg.cfg3d()
    with pytest.raises(L.CpmlError):
        solver3d(c, nslabs=2, slab_rank=0)
    c2 = refcfg.cfg2d(2, nx=60, ny=70, nstep=10, npml=6)
    with pytest.raises(L.CpmlError):
        L.Solver(ndim=2, order=2, nx=60, ny=70, nstep=10, npoints_pml=6, nrec=2, isource=c2["isource"], jsource=c2["jsource"],
                 deltax=10.0, deltay=10.0, deltat=2e-3, cp=3300.0, precision=1)
