"""GPU parity tests of the 2-D viscoelastic solvers (second and fourth order, N_SLS = 3): the CUDA
path through the C ABI against the CPU oracle.  Tolerance of north_star: relative L2 <= 1e-5; the
kernels keep the reference's operation order (-fmad=false), so the five fields, the nine memory
variables and the three seismograms are bit-identical."""
import os

import numpy as np
import pytest

import refcfg
from oracle import oracle as O
from seismic_cpml_b200 import lib as L
from seismic_cpml_b200 import programs as P

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-5
TOL_ENERGY = 1e-11


def solver_v2d(c, compute_energy=False):
    s = L.Solver(ndim=2, order=c["order"], rheology=1, compute_energy=compute_energy, nx=c["nx"], ny=c["ny"],
                 nstep=c["nstep"], npoints_pml=c["npoints_pml"], nrec=len(c["ix_rec"]), isource=c["isource"],
                 jsource=c["jsource"], deltax=c["deltax"], deltay=c["deltay"], deltat=c["deltat"], cp=0.0)
    s.set_profiles(L.AXIS_X, c["prof_x"])
    s.set_profiles(L.AXIS_Y, c["prof_y"])
    s.set_material_2d(c["lam"], c["mu"], c["rho"])
    s.set_attenuation(c["tau_epsilon_nu1"], c["tau_sigma_nu1"], c["tau_epsilon_nu2"], c["tau_sigma_nu2"])
    s.set_source_series(c["force_x"], c["force_y"])
    s.set_receivers(c["ix_rec"], c["iy_rec"])
    return s


@pytest.mark.parametrize("order", [2, 4])
@pytest.mark.parametrize("shape", [(81, 97, 8, 1.0), (130, 67, 6, 1.0), (64, 75, 7, 2.5)])
def test_visco2d_matches_oracle(order, shape):
    nx, ny, npml, kmax = shape
    c = refcfg.cfgv2d(order=order, nx=nx, ny=ny, npml=npml, nstep=250, material="layered", k_max=kmax)
    o = O.run_2d_visco(**c, want_fields=True, compute_energy=True)
    with solver_v2d(c, compute_energy=True) as s:
        s.run(1, c["nstep"])
        sx, sy = s.get_seismograms()
        sp = s.get_pressure_seismograms()
        for g, r in ((sx, o["sisvx"]), (sy, o["sisvy"]), (sp, o["sispressure"])):
            assert np.abs(r).max() > 0 and refcfg.rel_l2(g, r) <= TOL
            assert np.array_equal(g, r), np.abs(g - r).max()
        for f, name in enumerate(("vx", "vy", "sigmaxx", "sigmayy", "sigmaxy")):
            assert np.array_equal(s.get_field(f), o[name]), name
        for e, name in enumerate(("e1", "e11", "e13")):
            for l in range(3):
                assert np.array_equal(s.get_field(5 + 3 * e + l), o[name][l]), (name, l)
        tot, ek, ep = s.get_energy()
        assert refcfg.rel_l2(ek, o["energy_kinetic"]) <= TOL_ENERGY
        assert refcfg.rel_l2(ep, o["energy_potential"]) <= TOL_ENERGY
        assert s.get_maxnorm() == pytest.approx(o["velocnorm"], rel=1e-15)


def test_visco2d_energy_off_by_default_and_golden():
    for order in (2, 4):
        g = np.load(os.path.join(GOLD, f"cpml2d_visco_order{order}.npz"))
        c = refcfg.cfgv2d(order=order, material="layered", nstep=300)
        with solver_v2d(c) as s:
            s.run(1, c["nstep"])
            sx, sy = s.get_seismograms()
            sp = s.get_pressure_seismograms()
            tot, ek, ep = s.get_energy()
        assert np.array_equal(sx, g["sisvx"]) and np.array_equal(sy, g["sisvy"]) and np.array_equal(sp, g["sispressure"])
        assert not tot.any()                   # COMPUTE_ENERGY = .false. in the reference (:201)


def test_program2dvisco_mirror_and_elastic_branch():
    """The driver mirror on a reduced grid; VISCOELASTIC_ATTENUATION = False must equal the oracle's
    elastic branch (2D-visco-4th :713-760)."""
    for visco in (True, False):
        # relaxation times: the program's own SolvOpt fit to Qp = 65, Qs = 55 (2D-visco-4th :366-376)
        p = P.Params2DVisco(order=4, NX=121, NY=101, NSTEP=200, xsource=90.0, ysource=75.0, xdeb=120.0, ydeb=100.0,
                            xfin=120.0, yfin=100.0, VISCOELASTIC_ATTENUATION=visco)
        prog = P.Program2DVisco(p)
        res = prog.run()
        s = prog.s
        o = O.run_2d_visco(order=4, nx=p.NX, ny=p.NY, deltax=p.DELTAX, deltay=p.DELTAY, deltat=p.DELTAT, nstep=p.NSTEP,
                           npoints_pml=p.NPOINTS_PML, isource=p.ISOURCE, jsource=p.JSOURCE, lam=s.material[0],
                           mu=s.material[1], rho=s.material[2], tau_epsilon_nu1=p.tau_epsilon_nu1,
                           tau_sigma_nu1=p.tau_sigma_nu1, tau_epsilon_nu2=p.tau_epsilon_nu2,
                           tau_sigma_nu2=p.tau_sigma_nu2, prof_x=s.prof_x, prof_y=s.prof_y,
                           force_x=s.force_x, force_y=s.force_y, ix_rec=s.ix_rec, iy_rec=s.iy_rec,
                           viscoelastic_attenuation=visco)
        prog.solver.close()
        assert np.abs(o["sisvy"]).max() > 0
        assert np.array_equal(res["sisvx"], o["sisvx"]) and np.array_equal(res["sisvy"], o["sisvy"])
        assert np.array_equal(res["sispressure"], o["sispressure"])
