"""Independent numpy restatement of the 3-D viscoelastic time loop (TEST INFRASTRUCTURE).

Second, independently written restatement of seismic_CPML_3D_viscoelastic_MPI.f90:954-1430 used
to cross-check oracle/cpml_oracle_visco.c: whole-array slice arithmetic in ONE address space with
global k, instead of scalar loops over emulated MPI slabs.  The reference's incomplete z-halo
exchange is expressed here the way the CUDA kernels express it: at every
interface between two of the `emulate_nproc` reference slabs the stencil taps that the MPI
exchange never delivers read zero,
    stress phase   : vz(k+1) on the last plane of a slab (dvz_dz, :991);
                     vx(k-1), vy(k-1) on the first plane of a slab (dvx_dz :1149, dvy_dz :1189);
    velocity phase : sigmaxz(k+1), sigmayz(k+1) on the last plane (:1251, :1271);
                     sigmazz(k-1) on the first plane (:1294).
Agreement with the slab-emulating C oracle bit for bit (fields, seismograms) therefore checks
both restatements AND this tap analysis.  Elementwise IEEE double operations in the order of the
Fortran expressions; only the energy sums may differ in the last digits.

Arrays are indexed [i, j, k + 1] with the Fortran indices i = 0..NX+1, j = 0..NY+1,
k = -1..NZ+2 (global).
"""
from __future__ import annotations

import numpy as np


def _p1(a, n):
    out = np.zeros(n + 2)
    out[1:n + 1] = np.asarray(a, dtype=np.float64)
    return out


def run_3d_visco_np(*, nx, ny, nz, deltax, deltay, deltaz, deltat, lam, mu, rho, nstep, npoints_pml,
                    isource, jsource, tau_epsilon_nu1, tau_sigma_nu1, tau_epsilon_nu2, tau_sigma_nu2,
                    prof_x, prof_y, prof_z, force_x, force_y, ix_rec, iy_rec, emulate_nproc=1,
                    sigmazz_isotropic=False, **_ignored):
    """sigmazz_isotropic=True replaces the reference's memory-variable term of sigmazz (:1058-1060, quirk B14)
    by the isotropic form its sigmaxx / sigmayy use -- NOT the reference, only for the analytical check in
    tests/test_analytical_visco3d.py."""
    NX, NY, NZ, DT = nx, ny, nz, deltat
    sh = (NX + 2, NY + 2, NZ + 4)
    vx, vy, vz, sxx, syy, szz, sxy, sxz, syz = (np.zeros(sh) for _ in range(9))
    sxx_R, syy_R, szz_R, sxy_R, sxz_R, syz_R = (np.zeros(sh) for _ in range(6))
    mem = {n: np.zeros(sh) for n in (
        "dvx_dx", "dvx_dy", "dvx_dz", "dvy_dx", "dvy_dy", "dvy_dz", "dvz_dx", "dvz_dy", "dvz_dz",
        "dsxx_dx", "dsyy_dy", "dszz_dz", "dsxy_dx", "dsxy_dy", "dsxz_dx", "dsxz_dz", "dsyz_dy", "dsyz_dz")}
    e1, e11, e22, e12, e13, e23 = ([np.zeros(sh), np.zeros(sh)] for _ in range(6))
    X = {k: _p1(prof_x[k], NX)[:, None, None] for k in prof_x}
    Y = {k: _p1(prof_y[k], NY)[None, :, None] for k in prof_y}
    Zp = {}
    for k in prof_z:
        z = np.zeros(NZ + 4)
        z[2:NZ + 2] = np.asarray(prof_z[k], dtype=np.float64)
        Zp[k] = z[None, None, :]
    nrec = len(ix_rec)
    sisvx, sisvy = np.zeros((nrec, nstep)), np.zeros((nrec, nstep))
    et, ek, ep = np.zeros(nstep), np.zeros(nstep), np.zeros(nstep)

    odx, ody, odz = 1.0 / deltax, 1.0 / deltay, 1.0 / deltaz
    DT_r = deltat / rho
    ks = NZ // 2
    # :458-477
    inv1 = [1.0 / tau_sigma_nu1[0], 1.0 / tau_sigma_nu1[1]]
    inv2 = [1.0 / tau_sigma_nu2[0], 1.0 / tau_sigma_nu2[1]]
    phi1 = [(1.0 - tau_epsilon_nu1[l] / tau_sigma_nu1[l]) / tau_sigma_nu1[l] for l in range(2)]
    phi2 = [(1.0 - tau_epsilon_nu2[l] / tau_sigma_nu2[l]) / tau_sigma_nu2[l] for l in range(2)]
    Mu1 = 1.0 - (1.0 - tau_epsilon_nu1[0] / tau_sigma_nu1[0]) - (1.0 - tau_epsilon_nu1[1] / tau_sigma_nu1[1])
    Mu2 = 1.0 - (1.0 - tau_epsilon_nu2[0] / tau_sigma_nu2[0]) - (1.0 - tau_epsilon_nu2[1] / tau_sigma_nu2[1])
    # :982-987
    l2m_r = lam + 2.0 * mu
    lam_u = (lam + 2.0 / 3.0 * mu) * Mu1 - 2.0 / 3.0 * mu * Mu2
    mu_u = mu * Mu2
    l2m_u = lam_u + 2.0 * mu_u

    # quirk B6 masks over global k (1.0 = the tap is delivered, 0.0 = it reads the zero halo slot)
    nzl = NZ // emulate_nproc
    kk = np.arange(-1, NZ + 3)
    last_of_slab = (kk % nzl == 0) & (kk >= 1) & (kk < NZ)
    first_of_slab = (kk % nzl == 1) & (kk > 1) & (kk <= NZ)
    keep_up = np.where(last_of_slab, 0.0, 1.0)[None, None, :]     # multiplies f(k+1) taps
    keep_dn = np.where(first_of_slab, 0.0, 1.0)[None, None, :]    # multiplies f(k-1) taps

    def rng(I, J, K):
        def s(di=0, dj=0, dk=0):
            return (slice(I[0] + di, I[1] + 1 + di), slice(J[0] + dj, J[1] + 1 + dj),
                    slice(K[0] + 1 + dk, K[1] + 2 + dk))
        return s

    def coef(P, R, axis, off=0):
        sl = [slice(None)] * 3
        sl[axis] = slice(R[0] + off, R[1] + 1 + off)
        return {k: v[tuple(sl)] for k, v in P.items()}

    def d4(f, s, axis, forward, up=None, dn=None):
        """(27 a - 27 b - c + d) * (1/delta) / 24 with the reference's tap order."""
        sh3 = [[0, 0, 0] for _ in range(4)]
        if forward:      # 27 f(+1) - 27 f(0) - f(+2) + f(-1)
            offs = (1, 0, 2, -1)
        else:            # 27 f(0) - 27 f(-1) - f(+1) + f(-2)
            offs = (0, -1, 1, -2)
        for q, o in enumerate(offs):
            sh3[q][axis] = o
        a, b, c, d = (f[s(*t)] for t in sh3)
        if forward and dn is not None:
            d = d * dn           # f(k-1)
        if (not forward) and up is not None:
            c = c * up           # f(k+1)
        od = (odx, ody, odz)[axis]
        return (27.0 * a - 27.0 * b - c + d) * od / 24.0

    def cpml(name, value, C, half, sl):
        m = mem[name]
        m[sl] = C["b" + half] * m[sl] + C["a" + half] * value
        return value / C["K" + half] + m[sl]

    def evolve(e, l, S, inv, sl):
        tauinv = -inv[l]
        Un = e[l][sl]
        tauinvUn = tauinv * Un
        e[l][sl] = (Un + DT * (S + 0.5 * tauinvUn)) / (1.0 - DT * 0.5 * tauinv)

    P = npoints_pml
    EB = (slice(P, NX - P + 2), slice(P, NY - P + 2), slice(P + 1, NZ - P + 3))

    for it in range(1, nstep + 1):
        # sigmaxx, sigmayy, sigmazz : i=1..NX-1, j=2..NY, k=2..NZ
        I, J, K = (1, NX - 1), (2, NY), (2, NZ)
        s = rng(I, J, K)
        Xc, Yc, Zc = coef(X, I, 0), coef(Y, J, 1), coef(Zp, K, 2, 1)
        up = keep_up[:, :, K[0] + 1:K[1] + 2]
        dxx = d4(vx, s, 0, True)
        dyy = d4(vy, s, 1, False)
        dzz = d4(vz, s, 2, False, up=up)
        duxdx = cpml("dvx_dx", dxx, Xc, "_half", s())
        duydy = cpml("dvy_dy", dyy, Yc, "", s())
        duzdz = cpml("dvz_dz", dzz, Zc, "", s())
        div = duxdx + duydy + duzdz
        for l in range(2):
            evolve(e1, l, div * phi1[l], inv1, s())
        for l in range(2):
            evolve(e11, l, (duxdx - div / 3.0) * phi2[l], inv2, s())
        for l in range(2):
            evolve(e22, l, (duydy - div / 3.0) * phi2[l], inv2, s())
        q = s()
        sxx[q] = sxx[q] + DT * ((lam + 2.0 / 3.0 * mu) * (e1[0][q] + e1[1][q]) + 2.0 * mu * (e11[0][q] + e11[1][q]))
        syy[q] = syy[q] + DT * ((lam + 2.0 / 3.0 * mu) * (e1[0][q] + e1[1][q]) + 2.0 * mu * (e22[0][q] + e22[1][q]))
        if sigmazz_isotropic:
            szz[q] = szz[q] + DT * ((lam + 2.0 / 3.0 * mu) * (e1[0][q] + e1[1][q])
                                    - 2.0 * mu * (e11[0][q] + e11[1][q] + e22[0][q] + e22[1][q]))
        else:
            szz[q] = szz[q] + DT * ((lam + 2.0 * mu) * (e1[0][q] + e1[1][q])
                                    - 2.0 / 3.0 * mu * (e11[0][q] + e11[1][q] + e22[0][q] + e22[1][q]))
        sxx[q] = sxx[q] + (l2m_u * duxdx + lam_u * duydy + lam_u * duzdz) * DT
        syy[q] = syy[q] + (lam_u * duxdx + l2m_u * duydy + lam_u * duzdz) * DT
        szz[q] = szz[q] + (lam_u * duxdx + lam_u * duydy + l2m_u * duzdz) * DT
        sxx_R[q] = sxx_R[q] + (l2m_r * duxdx + lam * duydy + lam * duzdz) * DT
        syy_R[q] = syy_R[q] + (lam * duxdx + l2m_r * duydy + lam * duzdz) * DT
        szz_R[q] = szz_R[q] + (lam * duxdx + lam * duydy + l2m_r * duzdz) * DT

        # sigmaxy : i=2..NX, j=1..NY-1, k=1..NZ
        I, J, K = (2, NX), (1, NY - 1), (1, NZ)
        s = rng(I, J, K)
        Xc, Yc = coef(X, I, 0), coef(Y, J, 1)
        dyx = d4(vy, s, 0, False)
        dxy = d4(vx, s, 1, True)
        duydx = cpml("dvy_dx", dyx, Xc, "", s())
        duxdy = cpml("dvx_dy", dxy, Yc, "_half", s())
        for l in range(2):
            evolve(e12, l, (duxdy + duydx) * phi2[l], inv2, s())
        q = s()
        sxy[q] = sxy[q] + DT * mu * (e12[0][q] + e12[1][q])
        sxy[q] = sxy[q] + mu_u * (duxdy + duydx) * DT
        sxy_R[q] = sxy_R[q] + mu * (duxdy + duydx) * DT

        # sigmaxz : i=2..NX, j=1..NY, k=1..NZ-1
        I, J, K = (2, NX), (1, NY), (1, NZ - 1)
        s = rng(I, J, K)
        Xc, Zc = coef(X, I, 0), coef(Zp, K, 2, 1)
        dn = keep_dn[:, :, K[0] + 1:K[1] + 2]
        dzx = d4(vz, s, 0, False)
        dxz = d4(vx, s, 2, True, dn=dn)
        duzdx = cpml("dvz_dx", dzx, Xc, "", s())
        duxdz = cpml("dvx_dz", dxz, Zc, "_half", s())
        for l in range(2):
            evolve(e13, l, (duxdz + duzdx) * phi2[l], inv2, s())
        q = s()
        sxz[q] = sxz[q] + DT * mu * (e13[0][q] + e13[1][q])
        sxz[q] = sxz[q] + mu_u * (duxdz + duzdx) * DT
        sxz_R[q] = sxz_R[q] + mu * (duxdz + duzdx) * DT

        # sigmayz : i=1..NX, j=1..NY-1, k=1..NZ-1
        I, J, K = (1, NX), (1, NY - 1), (1, NZ - 1)
        s = rng(I, J, K)
        Yc, Zc = coef(Y, J, 1), coef(Zp, K, 2, 1)
        dzy = d4(vz, s, 1, True)
        dyz = d4(vy, s, 2, True, dn=dn)
        duzdy = cpml("dvz_dy", dzy, Yc, "_half", s())
        duydz = cpml("dvy_dz", dyz, Zc, "_half", s())
        for l in range(2):
            evolve(e23, l, (duydz + duzdy) * phi2[l], inv2, s())
        q = s()
        syz[q] = syz[q] + DT * mu * (e23[0][q] + e23[1][q])
        syz[q] = syz[q] + mu_u * (duydz + duzdy) * DT
        syz_R[q] = syz_R[q] + mu * (duydz + duzdy) * DT

        # vx : i=2..NX, j=2..NY, k=2..NZ
        I, J, K = (2, NX), (2, NY), (2, NZ)
        s = rng(I, J, K)
        Xc, Yc, Zc = coef(X, I, 0), coef(Y, J, 1), coef(Zp, K, 2, 1)
        up = keep_up[:, :, K[0] + 1:K[1] + 2]
        d1 = cpml("dsxx_dx", d4(sxx, s, 0, False), Xc, "", s())
        d2 = cpml("dsxy_dy", d4(sxy, s, 1, False), Yc, "", s())
        d3 = cpml("dsxz_dz", d4(sxz, s, 2, False, up=up), Zc, "", s())
        vx[s()] = DT_r * (d1 + d2 + d3) + vx[s()]

        # vy : i=1..NX-1, j=1..NY-1, k=2..NZ
        I, J, K = (1, NX - 1), (1, NY - 1), (2, NZ)
        s = rng(I, J, K)
        Xc, Yc, Zc = coef(X, I, 0), coef(Y, J, 1), coef(Zp, K, 2, 1)
        d1 = cpml("dsxy_dx", d4(sxy, s, 0, True), Xc, "_half", s())
        d2 = cpml("dsyy_dy", d4(syy, s, 1, True), Yc, "_half", s())
        d3 = cpml("dsyz_dz", d4(syz, s, 2, False, up=up), Zc, "", s())
        vy[s()] = DT_r * (d1 + d2 + d3) + vy[s()]

        # vz : i=1..NX-1, j=2..NY, k=1..NZ-1
        I, J, K = (1, NX - 1), (2, NY), (1, NZ - 1)
        s = rng(I, J, K)
        Xc, Yc, Zc = coef(X, I, 0), coef(Y, J, 1), coef(Zp, K, 2, 1)
        dn = keep_dn[:, :, K[0] + 1:K[1] + 2]
        d1 = cpml("dsxz_dx", d4(sxz, s, 0, True), Xc, "_half", s())
        d2 = cpml("dsyz_dy", d4(syz, s, 1, False), Yc, "", s())
        d3 = cpml("dszz_dz", d4(szz, s, 2, True, dn=dn), Zc, "_half", s())
        vz[s()] = DT_r * (d1 + d2 + d3) + vz[s()]

        # source at (ISOURCE, JSOURCE, NZ/2)
        vx[isource, jsource, ks + 1] = vx[isource, jsource, ks + 1] + force_x[it - 1] * deltat / rho
        vy[isource, jsource, ks + 1] = vy[isource, jsource, ks + 1] + force_y[it - 1] * deltat / rho

        # Dirichlet, two planes per face (global faces only)
        for f in (vx, vy, vz):
            f[0:2, :, :] = 0.0
            f[NX:NX + 2, :, :] = 0.0
            f[:, 0:2, :] = 0.0
            f[:, NY:NY + 2, :] = 0.0
            f[:, :, 1:3] = 0.0             # k = 0..1
            f[:, :, NZ + 1:NZ + 3] = 0.0   # k = NZ..NZ+1

        for r in range(nrec):
            sisvx[r, it - 1] = vx[ix_rec[r], iy_rec[r], ks + 1]
            sisvy[r, it - 1] = vy[ix_rec[r], iy_rec[r], ks + 1]

        kin = np.sum(0.5 * rho * (vx[EB] ** 2 + vy[EB] ** 2 + vz[EB] ** 2))
        den = 2.0 * mu * (3.0 * lam + 2.0 * mu)
        exx = (2.0 * (lam + mu) * sxx[EB] - lam * syy[EB] - lam * szz[EB]) / den
        eyy = (2.0 * (lam + mu) * syy[EB] - lam * sxx[EB] - lam * szz[EB]) / den
        exy, exz, eyz = sxy_R[EB] / (2.0 * mu), sxz_R[EB] / (2.0 * mu), syz_R[EB] / (2.0 * mu)
        pot = np.sum(0.5 * (exx * sxx_R[EB] + eyy * syy_R[EB] + eyy * syy_R[EB] + 2.0 * exy * sxy_R[EB]
                            + 2.0 * exz * sxz_R[EB] + 2.0 * eyz * syz_R[EB]))
        et[it - 1], ek[it - 1], ep[it - 1] = kin + pot, kin, pot

    inner = (slice(1, NX + 1), slice(1, NY + 1), slice(2, NZ + 2))
    names = ("vx", "vy", "vz", "sigmaxx", "sigmayy", "sigmazz", "sigmaxy", "sigmaxz", "sigmayz",
             "sigmaxx_R", "sigmayy_R", "sigmazz_R", "sigmaxy_R", "sigmaxz_R", "sigmayz_R")
    arrs = (vx, vy, vz, sxx, syy, szz, sxy, sxz, syz, sxx_R, syy_R, szz_R, sxy_R, sxz_R, syz_R)
    fields = {n: np.ascontiguousarray(f[inner].transpose(2, 1, 0)) for n, f in zip(names, arrs)}
    return dict(sisvx=sisvx, sisvy=sisvy, total_energy=et, energy_kinetic=ek, energy_potential=ep, **fields)
