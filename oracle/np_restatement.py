"""Independent numpy restatement of the reference time loops (TEST INFRASTRUCTURE).

Second, independently written restatement used to cross-check oracle/cpml_oracle.c
(SURVEY.md 7.2 "oracle self-check"): whole-array slice arithmetic instead of scalar
loops, single address space with global k for the 3-D program.  Elementwise IEEE
double operations in the same order as the Fortran expressions, so velocities and
stresses must agree with the C oracle bit for bit; only the energy sums may differ
in the last digits (different summation order).

Follows
  seismic_CPML_2D_isotropic_second_order.f90:550-713
  seismic_CPML_2D_isotropic_fourth_order.f90:551-714
  seismic_CPML_3D_isotropic_MPI_OpenMP.f90:802-1180
Arrays are indexed [i, j(, k)] with the Fortran 1-based indices used directly
(index 0 and N+1 are a zero ghost ring / the end halo planes).
"""
from __future__ import annotations

import numpy as np


def _p1(a, n):
    out = np.zeros(n + 2)
    out[1:n + 1] = np.asarray(a, dtype=np.float64)
    return out


def _sl(lo, hi, shift=0):
    return slice(lo + shift, hi + 1 + shift)


def run_2d_np(*, order, nx, ny, deltax, deltay, deltat, nstep, npoints_pml, isource, jsource,
              lam, mu, rho, prof_x, prof_y, force_x, force_y, ix_rec, iy_rec):
    NX, NY, DT = nx, ny, deltat
    sh = (NX + 2, NY + 2)
    fourth = order == 4
    vx, vy, sxx, syy, sxy = (np.zeros(sh) for _ in range(5))
    m_dvx_dx, m_dvx_dy, m_dvy_dx, m_dvy_dy = (np.zeros(sh) for _ in range(4))
    m_dsxx_dx, m_dsyy_dy, m_dsxy_dx, m_dsxy_dy = (np.zeros(sh) for _ in range(4))
    L, M, R = np.zeros(sh), np.zeros(sh), np.zeros(sh)
    L[1:NX + 1, 1:NY + 1] = np.asarray(lam, dtype=np.float64).reshape(NY, NX).T
    M[1:NX + 1, 1:NY + 1] = np.asarray(mu, dtype=np.float64).reshape(NY, NX).T
    R[1:NX + 1, 1:NY + 1] = np.asarray(rho, dtype=np.float64).reshape(NY, NX).T
    X = {k: _p1(prof_x[k], NX)[:, None] for k in prof_x}
    Y = {k: _p1(prof_y[k], NY)[None, :] for k in prof_y}
    nrec = len(ix_rec)
    sisvx, sisvy = np.zeros((nrec, nstep)), np.zeros((nrec, nstep))
    ek, ep = np.zeros(nstep), np.zeros(nstep)

    def d_fwd(f, axis, I, J, delta):
        """forward difference u(n+1)-u(n) (2nd) or its 4-point form (4th order)."""
        s = (lambda a, b: (_sl(I[0], I[1], a), _sl(J[0], J[1], b)))
        e = (1, 0) if axis == 0 else (0, 1)
        if not fourth:
            return (f[s(e[0], e[1])] - f[s(0, 0)]) / delta
        return (27.0 * f[s(e[0], e[1])] - 27.0 * f[s(0, 0)] - f[s(2 * e[0], 2 * e[1])]
                + f[s(-e[0], -e[1])]) / (24.0 * delta)

    def d_bwd(f, axis, I, J, delta):
        s = (lambda a, b: (_sl(I[0], I[1], a), _sl(J[0], J[1], b)))
        e = (1, 0) if axis == 0 else (0, 1)
        if not fourth:
            return (f[s(0, 0)] - f[s(-e[0], -e[1])]) / delta
        return (27.0 * f[s(0, 0)] - 27.0 * f[s(-e[0], -e[1])] - f[s(e[0], e[1])]
                + f[s(-2 * e[0], -2 * e[1])]) / (24.0 * delta)

    if fourth:
        eb = (npoints_pml, NX - npoints_pml + 1, npoints_pml, NY - npoints_pml + 1)
    else:
        eb = (npoints_pml + 1, NX - npoints_pml, npoints_pml + 1, NY - npoints_pml)
    EB = (_sl(eb[0], eb[1]), _sl(eb[2], eb[3]))

    for it in range(1, nstep + 1):
        # sigma_xx, sigma_yy : j = 2..NY, i = 1..NX-1
        I, J = (1, NX - 1), (2, NY)
        c = (_sl(*I), _sl(*J))
        cx, cy = _sl(*I), _sl(*J)
        lam_hx = 0.5 * (L[_sl(*I, 1), cy] + L[c])
        mu_hx = 0.5 * (M[_sl(*I, 1), cy] + M[c])
        l2m_hx = lam_hx + 2.0 * mu_hx
        dvx_dx = d_fwd(vx, 0, I, J, deltax)
        dvy_dy = d_bwd(vy, 1, I, J, deltay)
        m_dvx_dx[c] = X["b_half"][cx] * m_dvx_dx[c] + X["a_half"][cx] * dvx_dx
        m_dvy_dy[c] = Y["b"][:, cy] * m_dvy_dy[c] + Y["a"][:, cy] * dvy_dy
        dvx_dx = dvx_dx / X["K_half"][cx] + m_dvx_dx[c]
        dvy_dy = dvy_dy / Y["K"][:, cy] + m_dvy_dy[c]
        sxx[c] = sxx[c] + (l2m_hx * dvx_dx + lam_hx * dvy_dy) * DT
        syy[c] = syy[c] + (lam_hx * dvx_dx + l2m_hx * dvy_dy) * DT

        # sigma_xy : j = 1..NY-1, i = 2..NX
        I, J = (2, NX), (1, NY - 1)
        c = (_sl(*I), _sl(*J))
        cx, cy = _sl(*I), _sl(*J)
        mu_hy = 0.5 * (M[cx, _sl(*J, 1)] + M[c])
        dvy_dx = d_bwd(vy, 0, I, J, deltax)
        dvx_dy = d_fwd(vx, 1, I, J, deltay)
        m_dvy_dx[c] = X["b"][cx] * m_dvy_dx[c] + X["a"][cx] * dvy_dx
        m_dvx_dy[c] = Y["b_half"][:, cy] * m_dvx_dy[c] + Y["a_half"][:, cy] * dvx_dy
        dvy_dx = dvy_dx / X["K"][cx] + m_dvy_dx[c]
        Kq = Y["K"] if fourth else Y["K_half"]          # quirk B3, 2D-4th :596
        dvx_dy = dvx_dy / Kq[:, cy] + m_dvx_dy[c]
        sxy[c] = sxy[c] + mu_hy * (dvy_dx + dvx_dy) * DT

        # vx : j = 2..NY, i = 2..NX
        I, J = (2, NX), (2, NY)
        c = (_sl(*I), _sl(*J))
        cx, cy = _sl(*I), _sl(*J)
        dsxx_dx = d_bwd(sxx, 0, I, J, deltax)
        dsxy_dy = d_bwd(sxy, 1, I, J, deltay)
        m_dsxx_dx[c] = X["b"][cx] * m_dsxx_dx[c] + X["a"][cx] * dsxx_dx
        m_dsxy_dy[c] = Y["b"][:, cy] * m_dsxy_dy[c] + Y["a"][:, cy] * dsxy_dy
        dsxx_dx = dsxx_dx / X["K"][cx] + m_dsxx_dx[c]
        dsxy_dy = dsxy_dy / Y["K"][:, cy] + m_dsxy_dy[c]
        vx[c] = vx[c] + (dsxx_dx + dsxy_dy) * DT / R[c]

        # vy : j = 1..NY-1, i = 1..NX-1
        I, J = (1, NX - 1), (1, NY - 1)
        c = (_sl(*I), _sl(*J))
        cx, cy = _sl(*I), _sl(*J)
        rho_hh = 0.25 * (R[c] + R[_sl(*I, 1), cy] + R[_sl(*I, 1), _sl(*J, 1)] + R[cx, _sl(*J, 1)])
        dsxy_dx = d_fwd(sxy, 0, I, J, deltax)
        dsyy_dy = d_fwd(syy, 1, I, J, deltay)
        m_dsxy_dx[c] = X["b_half"][cx] * m_dsxy_dx[c] + X["a_half"][cx] * dsxy_dx
        m_dsyy_dy[c] = Y["b_half"][:, cy] * m_dsyy_dy[c] + Y["a_half"][:, cy] * dsyy_dy
        dsxy_dx = dsxy_dx / X["K_half"][cx] + m_dsxy_dx[c]
        dsyy_dy = dsyy_dy / Y["K_half"][:, cy] + m_dsyy_dy[c]
        vy[c] = vy[c] + (dsxy_dx + dsyy_dy) * DT / rho_hh

        # source
        i, j = isource, jsource
        rho_hh_s = 0.25 * (R[i, j] + R[i + 1, j] + R[i + 1, j + 1] + R[i, j + 1])
        vx[i, j] = vx[i, j] + force_x[it - 1] * DT / R[i, j]
        vy[i, j] = vy[i, j] + force_y[it - 1] * DT / rho_hh_s

        # Dirichlet
        for f in (vx, vy):
            f[1, :] = 0.0
            f[NX, :] = 0.0
            f[:, 1] = 0.0
            f[:, NY] = 0.0

        for r in range(nrec):
            sisvx[r, it - 1] = vx[ix_rec[r], iy_rec[r]]
            sisvy[r, it - 1] = vy[ix_rec[r], iy_rec[r]]

        ek[it - 1] = 0.5 * np.sum(R[EB] * (vx[EB] ** 2 + vy[EB] ** 2))
        l, m = L[EB], M[EB]
        exx = ((l + 2.0 * m) * sxx[EB] - l * syy[EB]) / (4.0 * m * (l + m))
        eyy = ((l + 2.0 * m) * syy[EB] - l * sxx[EB]) / (4.0 * m * (l + m))
        exy = sxy[EB] / (2.0 * m)
        ep[it - 1] = np.sum(0.5 * (exx * sxx[EB] + eyy * syy[EB] + 2.0 * exy * sxy[EB]))

    inner = (slice(1, NX + 1), slice(1, NY + 1))
    return dict(sisvx=sisvx, sisvy=sisvy, energy_kinetic=ek, energy_potential=ep,
                vx=vx[inner].T.copy(), vy=vy[inner].T.copy(), sigmaxx=sxx[inner].T.copy(),
                sigmayy=syy[inner].T.copy(), sigmaxy=sxy[inner].T.copy())


def run_3d_iso_np(*, nx, ny, nz, deltax, deltay, deltaz, deltat, lam, mu, lambdaplustwomu, rho,
                  nstep, npoints_pml, isource, jsource, prof_x, prof_y, prof_z,
                  force_x, force_y, ix_rec, iy_rec, energy_bug_compat=True, dtype=np.float64):
    """dtype=np.float32: the single-precision build the reference endorses (3D-iso :114-116) -- wavefields, memory
    variables, profiles and update constants in IEEE single (numpy rounds every operation, no FMA); the update constants
    are rounded once from their double values and the energy is summed in double, like the CUDA kernels do."""
    T = dtype
    NX, NY, NZ, P = nx, ny, nz, npoints_pml
    sh = (NX + 2, NY + 2, NZ + 2)
    vx, vy, vz, sxx, syy, szz, sxy, sxz, syz = (np.zeros(sh, dtype=T) for _ in range(9))
    mem = {n: np.zeros(sh, dtype=T) for n in (
        "dvx_dx", "dvx_dy", "dvx_dz", "dvy_dx", "dvy_dy", "dvy_dz", "dvz_dx", "dvz_dy", "dvz_dz",
        "dsxx_dx", "dsyy_dy", "dszz_dz", "dsxy_dx", "dsxy_dy", "dsxz_dx", "dsxz_dz",
        "dsyz_dy", "dsyz_dz")}
    X = {k: _p1(prof_x[k], NX).astype(T)[:, None, None] for k in prof_x}
    Y = {k: _p1(prof_y[k], NY).astype(T)[None, :, None] for k in prof_y}
    Z = {k: _p1(prof_z[k], NZ).astype(T)[None, None, :] for k in prof_z}
    odx, ody, odz = T(1.0 / deltax), T(1.0 / deltay), T(1.0 / deltaz)
    DT_l, DT_m, DT_l2m, DT_r = T(deltat * lam), T(deltat * mu), T(deltat * lambdaplustwomu), T(deltat / rho)
    nrec = len(ix_rec)
    sisvx, sisvy = np.zeros((nrec, nstep)), np.zeros((nrec, nstep))
    energy = np.zeros(nstep)
    ks = NZ // 2

    def rng(I, J, K):
        return (lambda a=0, b=0, c=0: (_sl(I[0], I[1], a), _sl(J[0], J[1], b), _sl(K[0], K[1], c)))

    def cpml(name, val, A, which, sl1):
        """memory = b*memory + a*value ; value = value / K + memory"""
        m = mem[name]
        m[sl1] = A["b" + which] * m[sl1] + A["a" + which] * val
        return val / A["K" + which] + m[sl1]

    def coef(A, idx, axis):
        """restrict the 1-D profile dict A to the index range idx along `axis`"""
        sl = [slice(None)] * 3
        sl[axis] = _sl(*idx)
        return {k: v[tuple(sl)] for k, v in A.items()}

    EB = (_sl(P + 1, NX - P), _sl(P + 1, NY - P), _sl(P + 1, NZ - P))

    for it in range(1, nstep + 1):
        # sigmaxx/yy/zz : i=1..NX-1, j=2..NY, k=2..NZ
        I, J, K = (1, NX - 1), (2, NY), (2, NZ)
        s = rng(I, J, K)
        Xc, Yc, Zc = coef(X, I, 0), coef(Y, J, 1), coef(Z, K, 2)
        dxx = (vx[s(1, 0, 0)] - vx[s()]) * odx
        dyy = (vy[s()] - vy[s(0, -1, 0)]) * ody
        dzz = (vz[s()] - vz[s(0, 0, -1)]) * odz
        dxx = cpml("dvx_dx", dxx, Xc, "_half", s())
        dyy = cpml("dvy_dy", dyy, Yc, "", s())
        dzz = cpml("dvz_dz", dzz, Zc, "", s())
        sxx[s()] = DT_l2m * dxx + DT_l * (dyy + dzz) + sxx[s()]
        syy[s()] = DT_l * (dxx + dzz) + DT_l2m * dyy + syy[s()]
        szz[s()] = DT_l * (dxx + dyy) + DT_l2m * dzz + szz[s()]

        # sigmaxy : i=2..NX, j=1..NY-1, k=1..NZ
        I, J, K = (2, NX), (1, NY - 1), (1, NZ)
        s = rng(I, J, K)
        Xc, Yc = coef(X, I, 0), coef(Y, J, 1)
        dyx = (vy[s()] - vy[s(-1, 0, 0)]) * odx
        dxy = (vx[s(0, 1, 0)] - vx[s()]) * ody
        dyx = cpml("dvy_dx", dyx, Xc, "", s())
        dxy = cpml("dvx_dy", dxy, Yc, "_half", s())
        sxy[s()] = DT_m * (dyx + dxy) + sxy[s()]

        # sigmaxz : i=2..NX, j=1..NY, k=1..NZ-1
        I, J, K = (2, NX), (1, NY), (1, NZ - 1)
        s = rng(I, J, K)
        Xc, Zc = coef(X, I, 0), coef(Z, K, 2)
        dzx = (vz[s()] - vz[s(-1, 0, 0)]) * odx
        dxz = (vx[s(0, 0, 1)] - vx[s()]) * odz
        dzx = cpml("dvz_dx", dzx, Xc, "", s())
        dxz = cpml("dvx_dz", dxz, Zc, "_half", s())
        sxz[s()] = DT_m * (dzx + dxz) + sxz[s()]

        # sigmayz : i=1..NX, j=1..NY-1, k=1..NZ-1
        I, J, K = (1, NX), (1, NY - 1), (1, NZ - 1)
        s = rng(I, J, K)
        Yc, Zc = coef(Y, J, 1), coef(Z, K, 2)
        dzy = (vz[s(0, 1, 0)] - vz[s()]) * ody
        dyz = (vy[s(0, 0, 1)] - vy[s()]) * odz
        dzy = cpml("dvz_dy", dzy, Yc, "_half", s())
        dyz = cpml("dvy_dz", dyz, Zc, "_half", s())
        syz[s()] = DT_m * (dzy + dyz) + syz[s()]

        # vx : i=2..NX, j=2..NY, k=2..NZ
        I, J, K = (2, NX), (2, NY), (2, NZ)
        s = rng(I, J, K)
        Xc, Yc, Zc = coef(X, I, 0), coef(Y, J, 1), coef(Z, K, 2)
        d1 = (sxx[s()] - sxx[s(-1, 0, 0)]) * odx
        d2 = (sxy[s()] - sxy[s(0, -1, 0)]) * ody
        d3 = (sxz[s()] - sxz[s(0, 0, -1)]) * odz
        d1 = cpml("dsxx_dx", d1, Xc, "", s())
        d2 = cpml("dsxy_dy", d2, Yc, "", s())
        d3 = cpml("dsxz_dz", d3, Zc, "", s())
        vx[s()] = DT_r * (d1 + d2 + d3) + vx[s()]

        # vy : i=1..NX-1, j=1..NY-1, k=2..NZ
        I, J, K = (1, NX - 1), (1, NY - 1), (2, NZ)
        s = rng(I, J, K)
        Xc, Yc, Zc = coef(X, I, 0), coef(Y, J, 1), coef(Z, K, 2)
        d1 = (sxy[s(1, 0, 0)] - sxy[s()]) * odx
        d2 = (syy[s(0, 1, 0)] - syy[s()]) * ody
        d3 = (syz[s()] - syz[s(0, 0, -1)]) * odz
        d1 = cpml("dsxy_dx", d1, Xc, "_half", s())
        d2 = cpml("dsyy_dy", d2, Yc, "_half", s())
        d3 = cpml("dsyz_dz", d3, Zc, "", s())
        vy[s()] = DT_r * (d1 + d2 + d3) + vy[s()]

        # vz : i=1..NX-1, j=2..NY, k=1..NZ-1
        I, J, K = (1, NX - 1), (2, NY), (1, NZ - 1)
        s = rng(I, J, K)
        Xc, Yc, Zc = coef(X, I, 0), coef(Y, J, 1), coef(Z, K, 2)
        d1 = (sxz[s(1, 0, 0)] - sxz[s()]) * odx
        d2 = (syz[s()] - syz[s(0, -1, 0)]) * ody
        d3 = (szz[s(0, 0, 1)] - szz[s()]) * odz
        d1 = cpml("dsxz_dx", d1, Xc, "_half", s())
        d2 = cpml("dsyz_dy", d2, Yc, "", s())
        d3 = cpml("dszz_dz", d3, Zc, "_half", s())
        vz[s()] = DT_r * (d1 + d2 + d3) + vz[s()]

        # source at (ISOURCE, JSOURCE, NZ/2)
        vx[isource, jsource, ks] = vx[isource, jsource, ks] + T(force_x[it - 1] * deltat / rho)
        vy[isource, jsource, ks] = vy[isource, jsource, ks] + T(force_y[it - 1] * deltat / rho)

        # Dirichlet on the six faces
        for f in (vx, vy, vz):
            f[1, :, :] = 0.0
            f[NX, :, :] = 0.0
            f[:, 1, :] = 0.0
            f[:, NY, :] = 0.0
            f[:, :, 1] = 0.0
            f[:, :, NZ] = 0.0

        for r in range(nrec):
            sisvx[r, it - 1] = vx[ix_rec[r], iy_rec[r], ks]
            sisvy[r, it - 1] = vy[ix_rec[r], iy_rec[r], ks]

        evx, evy, evz = (f[EB].astype(np.float64) for f in (vx, vy, vz))
        esxx, esyy, eszz, esxy, esxz, esyz = (f[EB].astype(np.float64) for f in (sxx, syy, szz, sxy, sxz, syz))
        kin = np.sum(0.5 * rho * (evx ** 2 + evy ** 2 + evz ** 2))
        den = 2.0 * mu * (3.0 * lam + 2.0 * mu)
        exx = (2.0 * (lam + mu) * esxx - lam * esyy - lam * eszz) / den
        eyy = (2.0 * (lam + mu) * esyy - lam * esxx - lam * eszz) / den
        ezz = (2.0 * (lam + mu) * eszz - lam * esxx - lam * esyy) / den
        exy, exz, eyz = esxy / (2.0 * mu), esxz / (2.0 * mu), esyz / (2.0 * mu)
        third = eyy * esyy if energy_bug_compat else ezz * eszz
        pot = np.sum(0.5 * (exx * esxx + eyy * esyy + third + 2.0 * exy * esxy
                            + 2.0 * exz * esxz + 2.0 * eyz * esyz))
        energy[it - 1] = kin + pot

    inner = (slice(1, NX + 1), slice(1, NY + 1), slice(1, NZ + 1))
    fields = {n: np.ascontiguousarray(f[inner].transpose(2, 1, 0)) for n, f in (
        ("vx", vx), ("vy", vy), ("vz", vz), ("sigmaxx", sxx), ("sigmayy", syy), ("sigmazz", szz),
        ("sigmaxy", sxy), ("sigmaxz", sxz), ("sigmayz", syz))}
    return dict(sisvx=sisvx, sisvy=sisvy, total_energy=energy, **fields)
