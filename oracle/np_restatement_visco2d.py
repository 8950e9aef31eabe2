"""Independent numpy restatement of the 2-D viscoelastic time loops (TEST INFRASTRUCTURE).

Second, independently written restatement of
  seismic_CPML_2D_velocity_and_stress_fourth_order_viscoelastic.f90:705-1066
  seismic_CPML_2D_velocity_and_stress_second_order_viscoelastic.f90 (same loop, second-order operator)
used to cross-check oracle/cpml_oracle_visco2d.c: whole-array slice arithmetic, elementwise IEEE
double operations in the order of the Fortran expressions, so velocities, stresses and memory
variables must agree with the C oracle bit for bit.  Arrays are indexed [i, j] with the Fortran
indices (0 and N+1 are the zero ghost ring of the fourth-order file).
"""
from __future__ import annotations

import numpy as np


def _p1(a, n):
    out = np.zeros(n + 2)
    out[1:n + 1] = np.asarray(a, dtype=np.float64)
    return out


def run_2d_visco_np(*, order, nx, ny, deltax, deltay, deltat, nstep, npoints_pml, isource, jsource, lam, mu, rho,
                    tau_epsilon_nu1, tau_sigma_nu1, tau_epsilon_nu2, tau_sigma_nu2, prof_x, prof_y, force_x,
                    force_y, ix_rec, iy_rec, **_ignored):
    NX, NY, DT = nx, ny, deltat
    sh = (NX + 2, NY + 2)
    fourth = order == 4
    vx, vy, sxx, syy, sxy = (np.zeros(sh) for _ in range(5))
    mem = {n: np.zeros(sh) for n in ("dvx_dx", "dvx_dy", "dvy_dx", "dvy_dy", "dsxx_dx", "dsyy_dy", "dsxy_dx", "dsxy_dy")}
    e1, e11, e13 = ([np.zeros(sh) for _ in range(3)] for _ in range(3))
    L, M, R = np.zeros(sh), np.zeros(sh), np.zeros(sh)
    L[1:NX + 1, 1:NY + 1] = np.asarray(lam, dtype=np.float64).reshape(NY, NX).T
    M[1:NX + 1, 1:NY + 1] = np.asarray(mu, dtype=np.float64).reshape(NY, NX).T
    R[1:NX + 1, 1:NY + 1] = np.asarray(rho, dtype=np.float64).reshape(NY, NX).T
    X = {k: _p1(prof_x[k], NX)[:, None] for k in prof_x}
    Y = {k: _p1(prof_y[k], NY)[None, :] for k in prof_y}
    nrec = len(ix_rec)
    sisvx, sisvy, sisp = np.zeros((nrec, nstep)), np.zeros((nrec, nstep)), np.zeros((nrec, nstep))

    c98x, c98y = 9.0 / (8.0 * deltax), 9.0 / (8.0 * deltay)
    c24x, c24y = 1.0 / (24.0 * deltax), 1.0 / (24.0 * deltay)
    odx, ody = 1.0 / deltax, 1.0 / deltay
    te1, ts1 = np.asarray(tau_epsilon_nu1, float), np.asarray(tau_sigma_nu1, float)
    te2, ts2 = np.asarray(tau_epsilon_nu2, float), np.asarray(tau_sigma_nu2, float)
    s1 = s2 = 0.0
    for l in range(3):
        s1 = s1 + te1[l] / ts1[l]
        s2 = s2 + te2[l] / ts2[l]
    half1 = [0.5 * DT / ts1[l] for l in range(3)]
    half2 = [0.5 * DT / ts2[l] for l in range(3)]
    mul1 = [1.0 / (1.0 + 0.5 * DT * (1.0 / ts1[l])) for l in range(3)]
    mul2 = [1.0 / (1.0 + 0.5 * DT * (1.0 / ts2[l])) for l in range(3)]
    phi1 = [DT * (1.0 - te1[l] / ts1[l]) / ts1[l] / s1 for l in range(3)]
    phi2 = [DT * (1.0 - te2[l] / ts2[l]) / ts2[l] / s2 for l in range(3)]

    def rng(I, J):
        def s(di=0, dj=0):
            return (slice(I[0] + di, I[1] + 1 + di), slice(J[0] + dj, J[1] + 1 + dj))
        return s

    def dfwd(f, s, axis):
        p1 = s(1, 0) if axis == 0 else s(0, 1)
        if not fourth:
            return (f[p1] - f[s()]) * (odx if axis == 0 else ody)
        m1 = s(-1, 0) if axis == 0 else s(0, -1)
        p2 = s(2, 0) if axis == 0 else s(0, 2)
        return (f[p1] - f[s()]) * (c98x if axis == 0 else c98y) + (f[m1] - f[p2]) * (c24x if axis == 0 else c24y)

    def dbwd(f, s, axis):
        m1 = s(-1, 0) if axis == 0 else s(0, -1)
        if not fourth:
            return (f[s()] - f[m1]) * (odx if axis == 0 else ody)
        m2 = s(-2, 0) if axis == 0 else s(0, -2)
        p1 = s(1, 0) if axis == 0 else s(0, 1)
        return (f[s()] - f[m1]) * (c98x if axis == 0 else c98y) + (f[m2] - f[p1]) * (c24x if axis == 0 else c24y)

    def cpml(name, value, C, half, sl):
        m = mem[name]
        m[sl] = C["b" + half] * m[sl] + C["a" + half] * value
        return value / C["K" + half] + m[sl]

    def cx(I):
        return {k: v[I[0]:I[1] + 1, :] for k, v in X.items()}

    def cy(J):
        return {k: v[:, J[0]:J[1] + 1] for k, v in Y.items()}

    for it in range(1, nstep + 1):
        # sigma_xx, sigma_yy, e1, e11 : j=2..NY, i=1..NX-1
        I, J = (1, NX - 1), (2, NY)
        s = rng(I, J)
        q = s()
        lhx = 0.5 * (L[s(1, 0)] + L[q])
        mhx = 0.5 * (M[s(1, 0)] + M[q])
        lpm = lhx + mhx
        l2m = lhx + 2.0 * mhx
        dxx = cpml("dvx_dx", dfwd(vx, s, 0), cx(I), "_half", q)
        dyy = cpml("dvy_dy", dbwd(vy, s, 1), cy(J), "", q)
        sum1 = 0.0
        sum11 = 0.0
        for l in range(3):
            old1 = e1[l][q].copy()
            old11 = e11[l][q].copy()
            e1[l][q] = (old1 + (dxx + dyy) * phi1[l] - old1 * half1[l]) * mul1[l]
            e11[l][q] = (old11 + 0.5 * (dxx - dyy) * phi2[l] - old11 * half2[l]) * mul2[l]
            sum1 = sum1 + e1[l][q] + old1
            sum11 = sum11 + e11[l][q] + old11
        sxx[q] = sxx[q] + (l2m * dxx + lhx * dyy + (0.5 * lpm * sum1 + mhx * sum11)) * DT
        syy[q] = syy[q] + (lhx * dxx + l2m * dyy + (0.5 * lpm * sum1 - mhx * sum11)) * DT

        # sigma_xy, e13 : j=1..NY-1, i=2..NX
        I, J = (2, NX), (1, NY - 1)
        s = rng(I, J)
        q = s()
        mhy = 0.5 * (M[s(0, 1)] + M[q])
        dyx = cpml("dvy_dx", dbwd(vy, s, 0), cx(I), "", q)
        dxy = cpml("dvx_dy", dfwd(vx, s, 1), cy(J), "_half", q)
        sum13 = 0.0
        for l in range(3):
            old = e13[l][q].copy()
            e13[l][q] = (old + (dyx + dxy) * phi2[l] - old * half2[l]) * mul2[l]
            sum13 = sum13 + e13[l][q] + old
        sxy[q] = sxy[q] + mhy * (dyx + dxy + 0.5 * sum13) * DT

        # vx : j=2..NY, i=2..NX
        I, J = (2, NX), (2, NY)
        s = rng(I, J)
        q = s()
        d1 = cpml("dsxx_dx", dbwd(sxx, s, 0), cx(I), "", q)
        d2 = cpml("dsxy_dy", dbwd(sxy, s, 1), cy(J), "", q)
        vx[q] = vx[q] + (d1 + d2) * DT / R[q]

        # vy : j=1..NY-1, i=1..NX-1
        I, J = (1, NX - 1), (1, NY - 1)
        s = rng(I, J)
        q = s()
        rh = 0.25 * (R[q] + R[s(1, 0)] + R[s(1, 1)] + R[s(0, 1)])
        d1 = cpml("dsxy_dx", dfwd(sxy, s, 0), cx(I), "_half", q)
        d2 = cpml("dsyy_dy", dfwd(syy, s, 1), cy(J), "_half", q)
        vy[q] = vy[q] + (d1 + d2) * DT / rh

        i, j = isource, jsource
        rh = 0.25 * (R[i, j] + R[i + 1, j] + R[i + 1, j + 1] + R[i, j + 1])
        vx[i, j] = vx[i, j] + force_x[it - 1] * DT / R[i, j]
        vy[i, j] = vy[i, j] + force_y[it - 1] * DT / rh

        for f in (vx, vy):
            f[1, :] = 0.0
            f[NX, :] = 0.0
            f[:, 1] = 0.0
            f[:, NY] = 0.0

        for r in range(nrec):
            i, j = ix_rec[r], iy_rec[r]
            sisvx[r, it - 1] = vx[i, j]
            sisvy[r, it - 1] = vy[i, j]
            lhx = 0.5 * (L[i + 1, j] + L[i, j])
            mhx = 0.5 * (M[i + 1, j] + M[i, j])
            exx = ((lhx + 2.0 * mhx) * sxx[i, j] - lhx * syy[i, j]) / (4.0 * mhx * (lhx + mhx))
            eyy = ((lhx + 2.0 * mhx) * syy[i, j] - lhx * sxx[i, j]) / (4.0 * mhx * (lhx + mhx))
            sisp[r, it - 1] = -(lhx + 2.0 / 3.0 * mhx) * (exx + eyy)

    inner = (slice(1, NX + 1), slice(1, NY + 1))
    out = {n: np.ascontiguousarray(f[inner].T) for n, f in (("vx", vx), ("vy", vy), ("sigmaxx", sxx), ("sigmayy", syy), ("sigmaxy", sxy))}
    for name, arr in (("e1", e1), ("e11", e11), ("e13", e13)):
        out[name] = np.stack([np.ascontiguousarray(a[inner].T) for a in arr])
    return dict(sisvx=sisvx, sisvy=sisvy, sispressure=sisp, **out)
