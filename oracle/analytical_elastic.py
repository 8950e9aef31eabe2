"""TEST INFRASTRUCTURE (never imported by the product): closed-form velocity of a point force in a
homogeneous, unbounded, isotropic elastic medium, used to pin the oracle of the ISOTROPIC solvers
physically.  The reference holds no analytical solution for its elastic programs, so
this is not a restatement of reference code: 3-D is the textbook full-space Green's function (Aki &
Richards, Quantitative Seismology, eq. 4.23), 2-D is the elastic branch (`TURN_ATTENUATION_OFF`) of the
Green's function restated in oracle/analytical_visco2d.py.

Source of the isotropic programs (seismic_CPML_3D_isotropic_MPI_OpenMP.f90:1055-1083, 2D-2nd :656-686):
  source_term(t) = -factor * 2 a (t - t0) exp(-a (t - t0)^2) = factor * g'(t),  g(t) = exp(-a (t - t0)^2),
  a = pi^2 f0^2; the programs add force * DELTAT / rho to the velocity of ONE grid point, i.e. a body force
  density `force` over one cell: a point force of force * cell volume.
  force_x acts on vx at node (i, j, k); force_y on vy, which the staggered grid holds at
  (i + 1/2, j + 1/2, k) -- the two components of the "point" force sit half a cell apart (A.1 of SURVEY.md).
"""
from __future__ import annotations

import numpy as np

PI = 3.141592653589793


def _g(t, a, t0, order):
    """g, g', g'' of the Gaussian exp(-a (t - t0)^2)."""
    u = t - t0
    g = np.exp(-a * u * u)
    if order == 0:
        return g
    if order == 1:
        return -2.0 * a * u * g
    return (-2.0 * a + 4.0 * a * a * u * u) * g


def velocity_3d(times, offset, i, j, *, cp, cs, rho, f0, t0, amplitude):
    """Component i of the particle velocity at `offset` (3-vector, receiver minus source) for a force
    amplitude * g'(t) along axis j.  Aki & Richards (4.23) with X0 = amplitude * g', differentiated once
    in time; the near-field integral int_{r/cp}^{r/cs} tau X0'(t - tau) dtau is integrated by parts."""
    times = np.asarray(times, dtype=np.float64)
    offset = np.asarray(offset, dtype=np.float64)
    r = float(np.sqrt(np.sum(offset * offset)))
    gam = offset / r
    a = PI * PI * f0 * f0
    dij = 1.0 if i == j else 0.0
    tp, ts = r / cp, r / cs
    near = (-ts * _g(times - ts, a, t0, 1) + tp * _g(times - tp, a, t0, 1)
            - _g(times - ts, a, t0, 0) + _g(times - tp, a, t0, 0))
    v = ((3.0 * gam[i] * gam[j] - dij) / r ** 3 * near
         + gam[i] * gam[j] * _g(times - tp, a, t0, 2) / (cp * cp * r)
         - (gam[i] * gam[j] - dij) * _g(times - ts, a, t0, 2) / (cs * cs * r))
    return amplitude / (4.0 * PI * rho) * v


def receiver_velocities_3d(times, mx, my, *, delta, cp, cs, rho, f0, t0, factor, angle_force_deg):
    """vx and vy the 3-D isotropic program records at a receiver (mx, my) cells from the source, in the
    source plane: superposition of the x force at the vx node and the y force at the vy node."""
    rad = angle_force_deg * PI / 180.0
    vol = delta ** 3
    fx, fy = np.sin(rad) * factor * vol, np.cos(rad) * factor * vol
    kw = dict(cp=cp, cs=cs, rho=rho, f0=f0, t0=t0)
    # positions in cells: force_x (0, 0), force_y (1/2, 1/2), vx receiver (mx, my), vy receiver (mx + 1/2, my + 1/2)
    d = delta
    vx = (velocity_3d(times, (mx * d, my * d, 0.0), 0, 0, amplitude=fx, **kw)
          + velocity_3d(times, ((mx - 0.5) * d, (my - 0.5) * d, 0.0), 0, 1, amplitude=fy, **kw))
    vy = (velocity_3d(times, ((mx + 0.5) * d, (my + 0.5) * d, 0.0), 1, 0, amplitude=fx, **kw)
          + velocity_3d(times, (mx * d, my * d, 0.0), 1, 1, amplitude=fy, **kw))
    return vx, vy


def receiver_velocities_2d(times, mx, my, *, delta, cp, cs, rho, f0, t0, factor, angle_force_deg, freqmax=80.0):
    """vx and vy the 2-D isotropic programs record at a receiver (mx, my) cells from the source: the elastic
    branch of the 2-D Green's function of oracle/analytical_visco2d.py (line force, plane strain), source
    time function factor * g'(t), x force at the vx node and y force at the vy node (half a cell apart)."""
    from . import analytical_visco2d as A
    times = np.asarray(times, dtype=np.float64)
    period = 8.0 * max(float(times.max()), 4.0 * t0)
    nfreq = int(np.ceil(freqmax * period))
    deltafreq = freqmax / nfreq
    freq = deltafreq * np.arange(1, nfreq)
    omega = 2.0 * PI * freq
    a = PI * PI * f0 * f0
    # spectrum of the VELOCITY response to the force factor * g'(t): (i omega)^2 * factor * G(omega),
    # G = sqrt(pi / a) exp(-omega^2 / 4a) exp(-i omega t0)
    spec = -(omega ** 2) * factor * np.sqrt(PI / a) * np.exp(-omega ** 2 / (4.0 * a)) * np.exp(-1j * omega * t0)
    one = (1.0,)
    v1, v2 = A.complex_velocities(omega, vp=cp, vs=cs, rho=rho, tau_epsilon_nu1=one, tau_sigma_nu1=one,
                                  tau_epsilon_nu2=one, tau_sigma_nu2=one, attenuation=False)
    rad = angle_force_deg * PI / 180.0
    area = delta * delta
    fx, fy = np.sin(rad) * area, np.cos(rad) * area
    d = delta

    def green(x1, x2):          # (u1, u2) for a unit force along axis 2 (vertical in the reference's wording)
        return A._green(omega, v1, v2, float(x1), float(x2), rho, 1.0)

    # force along y: as restated; force along x: the same function with the axes swapped
    ux_fy, _ = green((mx - 0.5) * d, (my - 0.5) * d)        # vx receiver seen from the y force
    _, uy_fy = green(mx * d, my * d)                        # vy receiver seen from the y force
    _, ux_fx = green(my * d, mx * d)                        # vx receiver seen from the x force (axes swapped)
    uy_fx, _ = green((my + 0.5) * d, (mx + 0.5) * d)        # vy receiver seen from the x force (axes swapped)
    phi_x = (fx * ux_fx + fy * ux_fy) * spec
    phi_y = (fx * uy_fx + fy * uy_fy) * spec
    vx, vy = np.empty(times.size), np.empty(times.size)
    for s in range(0, times.size, 512):
        e = np.exp(1j * np.outer(times[s:s + 512], omega))
        vx[s:s + 512] = 2.0 * deltafreq * (e @ phi_x).real
        vy[s:s + 512] = 2.0 * deltafreq * (e @ phi_y).real
    return vx, vy


def velocity_3d_visco(times, offset, i, j, *, lam_relaxed, mu_relaxed, rho, f0, t0, amplitude,
                      tau_epsilon_nu1, tau_sigma_nu1, tau_epsilon_nu2, tau_sigma_nu2, freqmax=None):
    """The same Green's function (Aki & Richards 4.23) in the frequency domain with the complex moduli of the
    3-D viscoelastic program (correspondence principle): from its memory-variable equations
    (seismic_CPML_3D_viscoelastic_MPI.f90:458-477, 1003-1077)
        K(w)  = K_relaxed  * (1 - L + sum_l (1 + i w tau_eps1_l) / (1 + i w tau_sig1_l)),  K_relaxed = lambda + 2/3 mu,
        mu(w) = mu_relaxed * (1 - L + sum_l (1 + i w tau_eps2_l) / (1 + i w tau_sig2_l)),
    i.e. the classical form WITHOUT the 1/L factor, relaxed moduli as reference.  Force amplitude * g'(t) along
    axis j, velocity component i, time dependence exp(+i w t)."""
    times = np.asarray(times, dtype=np.float64)
    offset = np.asarray(offset, dtype=np.float64)
    r = float(np.sqrt(np.sum(offset * offset)))
    gam = offset / r
    dij = 1.0 if i == j else 0.0
    a = PI * PI * f0 * f0
    if freqmax is None:
        freqmax = 8.0 * f0
    period = 8.0 * max(float(times.max()), 4.0 * t0)
    nfreq = int(np.ceil(freqmax * period))
    deltafreq = freqmax / nfreq
    w = 2.0 * PI * deltafreq * np.arange(1, nfreq)
    L = len(tau_sigma_nu1)
    m1 = 1.0 - L + sum((1.0 + 1j * w * e) / (1.0 + 1j * w * s) for e, s in zip(tau_epsilon_nu1, tau_sigma_nu1))
    m2 = 1.0 - L + sum((1.0 + 1j * w * e) / (1.0 + 1j * w * s) for e, s in zip(tau_epsilon_nu2, tau_sigma_nu2))
    kappa = (lam_relaxed + 2.0 / 3.0 * mu_relaxed) * m1
    mu = mu_relaxed * m2
    alpha = np.sqrt((kappa + 4.0 / 3.0 * mu) / rho)
    beta = np.sqrt(mu / rho)

    def prim(tau):          # antiderivative of tau exp(-i w tau)
        return np.exp(-1j * w * tau) * (1j * w * tau + 1.0) / (w * w)

    near = prim(r / beta) - prim(r / alpha)
    green = ((3.0 * gam[i] * gam[j] - dij) / r ** 3 * near
             + gam[i] * gam[j] * np.exp(-1j * w * r / alpha) / (alpha ** 2 * r)
             - (gam[i] * gam[j] - dij) * np.exp(-1j * w * r / beta) / (beta ** 2 * r)) / (4.0 * PI * rho)
    # velocity spectrum for the force amplitude * g'(t): (i w)^2 * amplitude * G(w) * green
    spec = -(w ** 2) * amplitude * np.sqrt(PI / a) * np.exp(-w ** 2 / (4.0 * a)) * np.exp(-1j * w * t0) * green
    v = np.empty(times.size)
    for s0 in range(0, times.size, 512):
        v[s0:s0 + 512] = 2.0 * deltafreq * (np.exp(1j * np.outer(times[s0:s0 + 512], w)) @ spec).real
    return v
