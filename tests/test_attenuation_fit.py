"""cpml_host_attenuation_fit (the SolvOpt fit of attenuation_model_with_SolvOpt.f90) against the one
known-answer the reference itself holds for it: the relaxation times hard-coded in its
analytical-solution program for the medium of the 2-D viscoelastic programs
(analytical_solution_viscoelastic_2D_plane_strain_Carcione_correct_with_1_over_L.f90:124-128,
N_SLS = 3, f0 = 35 Hz; Qp = 65, Qs = 55 at 2D-visco-4th :316-317).  Host code only, no device."""
import math

import numpy as np
import pytest

from seismic_cpml_b200 import lib as L
from seismic_cpml_b200 import programs as P

# analytical_solution_..._with_1_over_L.f90:124-128, digits exactly as printed there
REF_QP65 = dict(tau_epsilon=("2.408158185753685e-002", "4.699608990861351e-003", "9.567997872435925e-004"),
                tau_sigma=("2.256014638636808e-002", "4.508471279712252e-003", "8.937876403768840e-004"))
REF_QS55 = dict(tau_epsilon=("2.430544480527216e-002", "4.728107829226396e-003", "9.667252695863502e-004"),
                tau_sigma=("2.250919779429490e-002", "4.501388007338097e-003", "8.917332095369118e-004"))


def _q_of_model(tau_epsilon, tau_sigma, freq):
    """Quality factor of N Zener solids in parallel (with the 1/N factor of the reference's formulation,
    attenuation_model_with_SolvOpt.f90:22-31): Q = Re M / Im M, M = 1/N sum (1 + i w te) / (1 + i w ts)."""
    w = 2.0 * math.pi * np.asarray(freq)
    m = sum((1.0 + 1j * w * te) / (1.0 + 1j * w * ts) for te, ts in zip(tau_epsilon, tau_sigma)) / len(tau_sigma)
    return m.real / m.imag


@pytest.mark.parametrize("q, ref", [(65.0, REF_QP65), (55.0, REF_QS55)])
def test_fit_reproduces_the_reference_constants_to_the_last_printed_digit(q, ref):
    f_min, f_max = L.host_attenuation_band(35.0)                    # 2D-visco-4th :366-368
    assert f_max / f_min == pytest.approx(12.0) and math.sqrt(f_min * f_max) == pytest.approx(35.0)
    te, ts, info = L.host_attenuation_fit(3, q, 35.0, f_min, f_max, return_info=True)
    assert info[0] > 0                                              # SolvOpt: normal termination, iterations
    for mine, theirs in ((te, ref["tau_epsilon"]), (ts, ref["tau_sigma"])):
        for a, b in zip(mine, theirs):
            # 16 significant digits, the precision of the literals in the reference source
            assert "%.15e" % a == "%.15e" % float(b), (a, b)


@pytest.mark.parametrize("n_sls, q, f0", [(3, 65.0, 35.0), (3, 55.0, 35.0), (2, 20.0, 16.0), (2, 10.0, 16.0),
                                           (4, 100.0, 10.0), (5, 30.0, 1.0)])
def test_fitted_model_has_the_requested_q_over_the_band(n_sls, q, f0):
    te, ts, info = L.host_attenuation_fit(n_sls, q, f0, return_info=True)
    assert info[0] > 0
    assert all(e > s > 0.0 for e, s in zip(te, ts))                 # positivity (Blanc et al. 2015)
    if n_sls <= 3:
        assert list(ts) == sorted(ts, reverse=True)               # mechanisms stay in first-guess order
    f_min, f_max = L.host_attenuation_band(f0)
    freq = np.geomspace(f_min, f_max, 200)
    err = np.abs(_q_of_model(te, ts, freq) / q - 1.0).max()
    # two mechanisms cannot be flat over a decade: 6 %; three and more: well under 1 %
    assert err < (0.06 if n_sls == 2 else 0.01), err
    # the misfit SolvOpt reports: sum over 4N log-spaced frequencies of (Q_ref Im M - Re M)^2 (:1777-1824)
    k = 4 * n_sls
    fk = [f_min * (f_max / f_min) ** (float(np.float32(i - 1.0) / np.float32(k - 1.0))) for i in range(1, k + 1)]
    w = 2.0 * math.pi * np.asarray(fk)
    m = sum((1.0 + 1j * w * e) / (1.0 + 1j * w * s) for e, s in zip(te, ts)) / n_sls
    misfit = float(np.sum((q * m.imag - m.real) ** 2))
    assert misfit == pytest.approx(info[1], rel=1e-6)


def test_nonlinear_fit_improves_on_its_linear_first_guess():
    f_min, f_max = L.host_attenuation_band(35.0)
    te0, ts0 = L.host_attenuation_fit(3, 65.0, 35.0, linear_only=True)
    te1, ts1 = L.host_attenuation_fit(3, 65.0, 35.0)
    # first guess: relaxation frequencies log-spaced over the band (remplit_point :421-444)
    assert 1.0 / ts0[0] == pytest.approx(2.0 * math.pi * f_min, rel=1e-14)
    assert 1.0 / ts0[2] == pytest.approx(2.0 * math.pi * f_max, rel=1e-14)
    freq = np.geomspace(f_min, f_max, 200)
    e0 = np.abs(_q_of_model(te0, ts0, freq) / 65.0 - 1.0).max()
    e1 = np.abs(_q_of_model(te1, ts1, freq) / 65.0 - 1.0).max()
    assert e1 < 0.3 * e0, (e0, e1)


def test_bad_arguments_are_rejected():
    load = L.load()
    out = np.zeros(4)
    d = L._d
    # N = 1: the reference's first guess evaluates 0./0. (:469) -- rejected instead of looping on NaN
    assert load.cpml_host_attenuation_fit(1, 65.0, 35.0, 10.0, 120.0, d(out), d(out), None) == L.CPML_EINVAL
    assert load.cpml_host_attenuation_fit(3, -1.0, 35.0, 10.0, 120.0, d(out), d(out), None) == L.CPML_EINVAL
    assert load.cpml_host_attenuation_fit(3, 65.0, 35.0, 120.0, 10.0, d(out), d(out), None) == L.CPML_EINVAL
    assert load.cpml_host_attenuation_fit(3, 65.0, 35.0, 10.0, 120.0, None, d(out), None) == L.CPML_EINVAL
    with pytest.raises(L.CpmlError):
        L.host_attenuation_fit(1, 65.0, 35.0)


def test_program_parameter_blocks_fit_their_relaxation_times_like_the_reference():
    p2 = P.Params2DVisco()                                           # Qp = 65, Qs = 55, f0 = 35, N_SLS = 3
    for k in ("tau_epsilon_nu1", "tau_sigma_nu1", "tau_epsilon_nu2", "tau_sigma_nu2"):
        assert np.allclose(getattr(p2, k), P.TAU_2D_VISCO[k], rtol=1e-15, atol=0.0), k
    p3 = P.Params3DVisco()                                           # QKappa = 20, QMu = 10, f0_attenuation = 16 (:191-192)
    assert p3.tau_epsilon_nu1 == L.host_attenuation_fit(2, 20.0, 16.0)[0]
    assert p3.tau_sigma_nu2 == L.host_attenuation_fit(2, 10.0, 16.0)[1]
    # the same order of magnitude as the Carcione (1993) values the reference quotes for these Q (:402-413)
    for k, v in P.TAU_CARCIONE_1993.items():
        assert np.allclose(getattr(p3, k), v, rtol=0.5), k
    # explicit relaxation times bypass the fit; half a pair is an error
    p3c = P.Params3DVisco(**P.TAU_CARCIONE_1993)
    assert p3c.tau_sigma_nu1 == (0.0303, 0.0025)
    with pytest.raises(ValueError):
        P.Params3DVisco(tau_epsilon_nu1=(0.03, 0.003))
    # VISCOELASTIC_ATTENUATION = .false.: the dummy values of 2D-visco-4th :374-380, no fit
    assert P.Params2DVisco(VISCOELASTIC_ATTENUATION=False).tau_sigma_nu2 == (1.0, 1.0, 1.0)


def test_single_correction_division_tool(tmp_path):
    """tools/div_small_check.c: the one-correction division by 24 and 3 of the viscoelastic kernels equals a / c
    (a short run here; the full 11.2e9-case run is quoted in DESIGN.md)."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "div_small_check")
    subprocess.run(["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-o", exe, os.path.join(root, "tools", "div_small_check.c"), "-lm"],
                   check=True)
    r = subprocess.run([exe, "2"], capture_output=True, text=True)
    assert r.returncode == 0 and " bad 0" in r.stdout, r.stdout
