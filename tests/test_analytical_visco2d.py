"""The 2-D viscoelastic solvers against the analytical solution the reference validates them with
(analytical_solution_viscoelastic_2D_plane_strain_Carcione_correct_with_1_over_L.f90, overlaid on the
solver's seismograms by plotall_fit_is_perfect_for_viscoelastic_fourth_order.gnu).  This is the one
accuracy check the reference itself holds for this path; `oracle/analytical_visco2d.py` restates it.

Geometry of the reference comparison (2D-visco-4th :187-199, analytical :137-143): the vertical force
acts on vy, i.e. at (ISOURCE + 1/2, JSOURCE + 1/2); vx of a receiver m cells away sits (m - 1/2) cells
from it, vy m cells -- hence the analytical program's `801 - 1.5/2` for Vx and `801` for Vz.
Clock: the reference plots solver sample `it` at (it-1) DELTAT (:1178); the leapfrog velocity after step
`it` is centred half a step later, which the eye cannot see on the reference's plot but the L2 norm can.
Both clocks are tested."""
import math

import numpy as np
import pytest

import refcfg
from oracle import analytical_visco2d as A
from oracle import oracle as O

DX, DT, F0 = 1.5, 2.2e-4, 35.0
MEDIUM = dict(vp=2000.0, vs=2000.0 / 1.732, rho=2000.0, f0=F0, t0=1.2 / F0, **refcfg.TAU_2D_VISCO)


def _rel(a, b):
    return math.sqrt(float(np.sum((a - b) ** 2)) / float(np.sum(b ** 2)))


def _analytic(nstep, mx, my, shift, **kw):
    t = (np.arange(nstep) + shift) * DT
    vx, _ = A.velocity_seismograms(t, (mx - 0.5) * DX, (my - 0.5) * DX, **{**MEDIUM, **kw})
    _, vy = A.velocity_seismograms(t, mx * DX, my * DX, **{**MEDIUM, **kw})
    return vx, vy


def _oracle(order, n, nstep, mx, my):
    c = refcfg.cfgv2d(order=order, nx=n, ny=n, nstep=nstep, npml=10)
    src = (n - max(mx, my)) // 2
    c["isource"], c["jsource"] = src, src
    c["ix_rec"] = np.array([src + mx], dtype=np.int32)
    c["iy_rec"] = np.array([src + my], dtype=np.int32)
    o = O.run_2d_visco(**c)
    return o["sisvx"][0], o["sisvy"][0]


def test_complex_velocities_tend_to_the_unrelaxed_ones():
    v1, v2 = A.complex_velocities(np.array([2.0 * math.pi * 1e7]), **{k: MEDIUM[k] for k in MEDIUM if k not in ("f0", "t0")})
    assert abs(v1[0]) == pytest.approx(2000.0, rel=1e-6) and abs(v2[0]) == pytest.approx(2000.0 / 1.732, rel=1e-6)
    # waves slow down and attenuate inside the band: Q = Re(V^2) / Im(V^2) ~ 55 for S at f0 (Qs of 2D-visco-4th :317)
    _, v2 = A.complex_velocities(np.array([2.0 * math.pi * F0]), **{k: MEDIUM[k] for k in MEDIUM if k not in ("f0", "t0")})
    assert (v2[0] ** 2).real / (v2[0] ** 2).imag == pytest.approx(55.0, rel=0.01)
    assert abs(v2[0]) < 2000.0 / 1.732


def test_fourth_order_oracle_fits_the_analytical_solution():
    nstep, mx, my = 1100, 70, 55
    sx, sy = _oracle(4, 301, nstep, mx, my)
    ax, ay = _analytic(nstep, mx, my, 0.0)
    bx, by = _analytic(nstep, mx, my, 0.5)
    ex, ey = _analytic(nstep, mx, my, 0.5, attenuation=False)
    assert abs(np.abs(sx).max() / np.abs(ax).max() - 1.0) < 5e-3          # peak amplitudes (measured 1.1e-3)
    assert abs(np.abs(sy).max() / np.abs(ay).max() - 1.0) < 5e-3
    assert _rel(sx, ax) < 0.05 and _rel(sy, ay) < 0.05                    # the reference's clock (measured 3.2 %)
    assert _rel(sx, bx) < 0.01 and _rel(sy, by) < 0.01                    # leapfrog clock (measured 0.43 %)
    assert _rel(sx, ex) > 0.3 and _rel(sy, ey) > 0.3                      # the elastic solution does not fit


def test_second_order_oracle_fits_the_analytical_solution_up_to_its_dispersion():
    nstep, mx, my = 1100, 70, 55
    sx, sy = _oracle(2, 301, nstep, mx, my)
    ax, ay = _analytic(nstep, mx, my, 0.0)
    assert abs(np.abs(sx).max() / np.abs(ax).max() - 1.0) < 5e-3
    assert abs(np.abs(sy).max() / np.abs(ay).max() - 1.0) < 5e-3
    assert _rel(sx, ax) < 0.12 and _rel(sy, ay) < 0.12                    # measured 8.0 % / 8.8 %


@pytest.mark.gpu
def test_reference_configuration_on_the_gpu_fits_the_analytical_solution():
    """seismic_CPML_2D_velocity_and_stress_fourth_order_viscoelastic.f90 exactly as shipped (2001 x 2001,
    5200 steps, source (1500, 1500), receiver (2301, 2301), relaxation times from the SolvOpt fit) through
    the program mirror on the GPU, against the analytical program's receiver (801 - 0.75 / 801 m)."""
    from seismic_cpml_b200 import programs as P
    p = P.Params2DVisco()
    assert (p.NX, p.NY, p.NSTEP, p.ISOURCE, p.JSOURCE) == (2001, 2001, 5200, 1001, 1001)
    prog = P.Program2DVisco(p)
    res = prog.run()
    assert list(prog.s.ix_rec) == [1535] and list(prog.s.iy_rec) == [1535]
    prog.solver.close()
    sx, sy = res["sisvx"][0], res["sisvy"][0]
    m = 1535 - 1001
    ax, ay = _analytic(p.NSTEP, m, m, 0.0)
    bx, by = _analytic(p.NSTEP, m, m, 0.5)
    ex, ey = _analytic(p.NSTEP, m, m, 0.5, attenuation=False)
    print("rel L2 (reference clock)", _rel(sx, ax), _rel(sy, ay), "(leapfrog clock)", _rel(sx, bx), _rel(sy, by))
    # measured on B200 (and, identically to 12 digits, with the CPU oracle in 25 minutes): 4.03 % on the
    # reference's clock, 1.91 % on the leapfrog clock after 34 S wavelengths, peak amplitudes within 0.04 %
    assert abs(np.abs(sx).max() / np.abs(ax).max() - 1.0) < 0.005
    assert abs(np.abs(sy).max() / np.abs(ay).max() - 1.0) < 0.005
    assert _rel(sx, ax) < 0.05 and _rel(sy, ay) < 0.05
    assert _rel(sx, bx) < 0.025 and _rel(sy, by) < 0.025
    assert _rel(sx, ex) > 0.5 and _rel(sy, ey) > 0.5
