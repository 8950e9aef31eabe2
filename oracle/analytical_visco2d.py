"""TEST INFRASTRUCTURE (never imported by the product): the analytical velocity seismograms of a vertical
point force in a homogeneous 2-D plane-strain viscoelastic medium, restated from the reference's
analytical_solution_viscoelastic_2D_plane_strain_Carcione_correct_with_1_over_L.f90 (Carcione, Kosloff &
Kosloff 1988, GJI 95, Appendix B, with the two typos the reference fixes).  The reference overlays its
output on the seismograms of seismic_CPML_2D_velocity_and_stress_fourth_order_viscoelastic.f90
(plotall_fit_is_perfect_for_viscoelastic_fourth_order.gnu): it is the only accuracy check the reference
holds for the viscoelastic solver, so tests use it to pin the oracle and the CUDA path physically.

What is restated (file:line of the analytical program):
  source spectrum, Ricker centred on t0, times i*omega for velocity   :179-205
  complex moduli of the N Zener solids, unrelaxed reference            :218-244
  complex P and S velocities                                           :246-255
  Green's function u1 (horizontal), u2 (vertical), G1, G2              :428-536
  Hankel functions of the second kind, orders 0 and 1                  :540-582
  synthesis in time                                                    :275-345

What is deliberately different: the reference evaluates the Hankel functions in single precision (NAG
S17DLE) and synthesises 33.5 million time samples with a single-precision FFT; here scipy.special.hankel2
in double precision and a direct Fourier sum at the requested times.  The frequency step is chosen from the
length of the requested window (period >= 8 x the window) instead of the reference's fixed 400 / 524288 Hz:
the signal is causal and decays, so the aliased copies are empty.
"""
from __future__ import annotations

import numpy as np

PI = 3.141592653589793


def complex_velocities(omega, *, vp, vs, rho, tau_epsilon_nu1, tau_sigma_nu1, tau_epsilon_nu2, tau_sigma_nu2,
                       attenuation=True):
    """:218-255 -- V1 (P) and V2 (S) at angular frequencies omega; vp, vs are the UNRELAXED velocities."""
    omega = np.asarray(omega, dtype=np.float64)
    m2_unrelaxed = vs ** 2 * 2.0 * rho                                           # :61
    m1_unrelaxed = 2.0 * vp ** 2 * rho - m2_unrelaxed                            # :62
    te1, ts1 = np.asarray(tau_epsilon_nu1), np.asarray(tau_sigma_nu1)
    te2, ts2 = np.asarray(tau_epsilon_nu2), np.asarray(tau_sigma_nu2)
    if attenuation:
        temp = sum((1.0 + 1j * omega * e) / (1.0 + 1j * omega * s) for e, s in zip(te1, ts1))
        m1c = (m1_unrelaxed / np.sum(te1 / ts1)) * temp                          # :223
        temp = sum((1.0 + 1j * omega * e) / (1.0 + 1j * omega * s) for e, s in zip(te2, ts2))
        m2c = (m2_unrelaxed / np.sum(te2 / ts2)) * temp                          # :230
    else:                                                                        # TURN_ATTENUATION_OFF :232-239
        m1c = np.full(omega.shape, m1_unrelaxed, dtype=np.complex128)
        m2c = np.full(omega.shape, m2_unrelaxed, dtype=np.complex128)
    e = (m1c + m2c) / 2.0
    return np.sqrt(e / rho), np.sqrt(m2c / (2.0 * rho))


def _green(omega, v1, v2, x1, x2, rho, force):
    """u1, u2 of :428-482 with G1, G2 of :486-536 (full solution, near field included)."""
    from scipy.special import hankel2
    r = np.sqrt(x1 ** 2 + x2 ** 2)
    h0_1, h0_2 = hankel2(0, omega * r / v1), hankel2(0, omega * r / v2)
    h1_1, h1_2 = hankel2(1, omega * r / v1), hankel2(1, omega * r / v2)
    g1 = (h0_1 / v1 ** 2 + h1_2 / (omega * r * v2) - h1_1 / (omega * r * v1)) * (-0.5j * PI)
    g2 = (h0_2 / v2 ** 2 - h1_2 / (omega * r * v2) + h1_1 / (omega * r * v1)) * (+0.5j * PI)
    den = 2.0 * PI * rho * r ** 2
    u1 = force * x1 * x2 * (g1 + g2) / den
    u2 = force * (x2 * x2 * g1 - x1 * x1 * g2) / den
    return u1, u2


def velocity_seismograms(times, x1, x2, *, vp, vs, rho, f0, t0, force=1.0, tau_epsilon_nu1, tau_sigma_nu1,
                         tau_epsilon_nu2, tau_sigma_nu2, attenuation=True, freqmax=400.0, chunk=512):
    """Horizontal and vertical velocity at offset (x1, x2) from a vertical force with a Ricker time
    function of dominant frequency f0 centred on t0, at absolute times `times` (the same clock as the
    finite-difference source: t = (it - 1) * DELTAT)."""
    times = np.asarray(times, dtype=np.float64)
    period = 8.0 * max(float(times.max()), 4.0 * t0)
    nfreq = int(np.ceil(freqmax * period))
    deltafreq = freqmax / nfreq
    freq = deltafreq * np.arange(1, nfreq - 1)          # the reference fills ifreq = 1 .. nfreq-2 (:288-291)
    omega = 2.0 * PI * freq
    a = PI ** 2 * f0 ** 2
    # Ricker spectrum centred on t0 (:198), times i*omega: velocity instead of displacement (:200)
    fomega = (np.sqrt(PI) * np.exp(-1j * omega * t0) * omega ** 2 * np.exp(-omega ** 2 / (4.0 * a))
              / (2.0 * np.sqrt(a ** 3))) * (1j * omega)
    v1, v2 = complex_velocities(omega, vp=vp, vs=vs, rho=rho, tau_epsilon_nu1=tau_epsilon_nu1,
                                tau_sigma_nu1=tau_sigma_nu1, tau_epsilon_nu2=tau_epsilon_nu2,
                                tau_sigma_nu2=tau_sigma_nu2, attenuation=attenuation)
    u1, u2 = _green(omega, v1, v2, float(x1), float(x2), rho, force)
    phi1, phi2 = u1 * fomega, u2 * fomega
    # the zero-frequency term carries omega**3 = 0; negative frequencies are the conjugates (:259-263), so
    # v(t) = deltafreq * 2 Re sum_k phi_k exp(+i omega_k t)   (inverse DFT / (nt * deltat), :311-317)
    keep = freq <= 8.0 * f0                              # exp(-(f/f0)^2) < 2e-28 beyond
    omega, phi1, phi2 = omega[keep], phi1[keep], phi2[keep]
    vx, vz = np.empty(times.size), np.empty(times.size)
    for s in range(0, times.size, chunk):
        e = np.exp(1j * np.outer(times[s:s + chunk], omega))
        vx[s:s + chunk] = 2.0 * deltafreq * (e @ phi1).real
        vz[s:s + chunk] = 2.0 * deltafreq * (e @ phi2).real
    return vx, vz
