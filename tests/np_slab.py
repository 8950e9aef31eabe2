"""numpy z-slab stepper (TEST INFRASTRUCTURE): one emulated MPI rank of
seismic_CPML_3D_isotropic_MPI_OpenMP.f90:802-1180 with the same step_* / plane interface
as the GPU slab, so that seismic_cpml_b200.slab.SlabDriver (the product's exchange and
reduction logic) can be exercised on CPU with gloo.  Arrays are [k, j, i] so that a z plane
is contiguous; index 0 / N+1 along j, i is an unused zero ring."""
import numpy as np
import torch


class NumpySlab:
    def __init__(self, c, nslabs, rank):
        self.c = c
        self.nx, self.ny, self.nz = c["nx"], c["ny"], c["nz"]
        self.nzl = self.nz // nslabs
        self.koff = rank * self.nzl
        sh = (self.nzl + 2, self.ny + 2, self.nx + 2)
        names = ("vx", "vy", "vz", "sxx", "syy", "szz", "sxy", "sxz", "syz")
        self.f = {n: np.zeros(sh) for n in names}
        self.order = names
        self.mem = {}
        nstep, nrec = c["nstep"], len(c["ix_rec"])
        self.sisvx, self.sisvy = np.zeros((nrec, nstep)), np.zeros((nrec, nstep))
        self.energy = np.zeros(nstep)
        ks = self.nz // 2 - self.koff
        self.ksrc = ks if 1 <= ks <= self.nzl else 0

        def p1(a, n):
            o = np.zeros(n + 2)
            o[1:n + 1] = a
            return o
        self.X = {k: p1(v, self.nx)[None, None, :] for k, v in c["prof_x"].items()}
        self.Y = {k: p1(v, self.ny)[None, :, None] for k, v in c["prof_y"].items()}
        zfull = {k: p1(v, self.nz) for k, v in c["prof_z"].items()}
        # local view of the global z profiles: local k -> global k + koff
        self.Z = {}
        for k, v in zfull.items():
            loc = np.zeros(self.nzl + 2)
            loc[1:self.nzl + 1] = v[self.koff + 1:self.koff + self.nzl + 1]
            self.Z[k] = loc[:, None, None]

    # -- interface used by SlabDriver
    def plane(self, field, klocal):
        return torch.from_numpy(self.f[self.order[field]][klocal])

    def synchronize(self):
        pass

    def get_seismograms(self):
        return self.sisvx, self.sisvy

    def get_energy(self):
        return self.energy, None, None

    def get_maxnorm(self):
        f = self.f
        return float(np.sqrt(f["vx"] ** 2 + f["vy"] ** 2 + f["vz"] ** 2)[1:self.nzl + 1].max())

    # -- helpers
    def _rng(self, I, J, K):
        def s(di=0, dj=0, dk=0):
            return (slice(K[0] + dk, K[1] + 1 + dk), slice(J[0] + dj, J[1] + 1 + dj), slice(I[0] + di, I[1] + 1 + di))
        return s

    def _cpml(self, name, val, A, which, sl, idx, axis):
        m = self.mem.setdefault(name, np.zeros_like(self.f["vx"]))
        cs = [slice(None)] * 3
        cs[axis] = slice(idx[0], idx[1] + 1)
        cs = tuple(cs)
        m[sl] = A["b" + which][cs] * m[sl] + A["a" + which][cs] * val
        return val / A["K" + which][cs] + m[sl]

    def step_stress(self, it):
        c, f = self.c, self.f
        NX, NY, nzl = self.nx, self.ny, self.nzl
        odx, ody, odz = 1.0 / c["deltax"], 1.0 / c["deltay"], 1.0 / c["deltaz"]
        dt = c["deltat"]
        DT_l, DT_m, DT_l2m = dt * c["lam"], dt * c["mu"], dt * c["lambdaplustwomu"]
        k2 = 2 if self.koff == 0 else 1                            # k2begin :792
        km1 = nzl - 1 if self.koff + nzl == self.nz else nzl       # kminus1end :795
        X, Y, Z = self.X, self.Y, self.Z
        vx, vy, vz = f["vx"], f["vy"], f["vz"]
        I, J, K = (1, NX - 1), (2, NY), (k2, nzl)
        s = self._rng(I, J, K)
        dxx = (vx[s(1)] - vx[s()]) * odx
        dyy = (vy[s()] - vy[s(0, -1)]) * ody
        dzz = (vz[s()] - vz[s(0, 0, -1)]) * odz
        dxx = self._cpml("dvx_dx", dxx, X, "_half", s(), I, 2)
        dyy = self._cpml("dvy_dy", dyy, Y, "", s(), J, 1)
        dzz = self._cpml("dvz_dz", dzz, Z, "", s(), K, 0)
        f["sxx"][s()] = DT_l2m * dxx + DT_l * (dyy + dzz) + f["sxx"][s()]
        f["syy"][s()] = DT_l * (dxx + dzz) + DT_l2m * dyy + f["syy"][s()]
        f["szz"][s()] = DT_l * (dxx + dyy) + DT_l2m * dzz + f["szz"][s()]
        I, J, K = (2, NX), (1, NY - 1), (1, nzl)
        s = self._rng(I, J, K)
        a = (vy[s()] - vy[s(-1)]) * odx
        b = (vx[s(0, 1)] - vx[s()]) * ody
        a = self._cpml("dvy_dx", a, X, "", s(), I, 2)
        b = self._cpml("dvx_dy", b, Y, "_half", s(), J, 1)
        f["sxy"][s()] = DT_m * (a + b) + f["sxy"][s()]
        I, J, K = (2, NX), (1, NY), (1, km1)
        s = self._rng(I, J, K)
        a = (vz[s()] - vz[s(-1)]) * odx
        b = (vx[s(0, 0, 1)] - vx[s()]) * odz
        a = self._cpml("dvz_dx", a, X, "", s(), I, 2)
        b = self._cpml("dvx_dz", b, Z, "_half", s(), K, 0)
        f["sxz"][s()] = DT_m * (a + b) + f["sxz"][s()]
        I, J, K = (1, NX), (1, NY - 1), (1, km1)
        s = self._rng(I, J, K)
        a = (vz[s(0, 1)] - vz[s()]) * ody
        b = (vy[s(0, 0, 1)] - vy[s()]) * odz
        a = self._cpml("dvz_dy", a, Y, "_half", s(), J, 1)
        b = self._cpml("dvy_dz", b, Z, "_half", s(), K, 0)
        f["syz"][s()] = DT_m * (a + b) + f["syz"][s()]

    def step_velocity(self, it):
        c, f = self.c, self.f
        NX, NY, nzl = self.nx, self.ny, self.nzl
        odx, ody, odz = 1.0 / c["deltax"], 1.0 / c["deltay"], 1.0 / c["deltaz"]
        DT_r = c["deltat"] / c["rho"]
        k2 = 2 if self.koff == 0 else 1
        km1 = nzl - 1 if self.koff + nzl == self.nz else nzl
        X, Y, Z = self.X, self.Y, self.Z
        I, J, K = (2, NX), (2, NY), (k2, nzl)
        s = self._rng(I, J, K)
        d1 = (f["sxx"][s()] - f["sxx"][s(-1)]) * odx
        d2 = (f["sxy"][s()] - f["sxy"][s(0, -1)]) * ody
        d3 = (f["sxz"][s()] - f["sxz"][s(0, 0, -1)]) * odz
        d1 = self._cpml("dsxx_dx", d1, X, "", s(), I, 2)
        d2 = self._cpml("dsxy_dy", d2, Y, "", s(), J, 1)
        d3 = self._cpml("dsxz_dz", d3, Z, "", s(), K, 0)
        f["vx"][s()] = DT_r * (d1 + d2 + d3) + f["vx"][s()]
        I, J, K = (1, NX - 1), (1, NY - 1), (k2, nzl)
        s = self._rng(I, J, K)
        d1 = (f["sxy"][s(1)] - f["sxy"][s()]) * odx
        d2 = (f["syy"][s(0, 1)] - f["syy"][s()]) * ody
        d3 = (f["syz"][s()] - f["syz"][s(0, 0, -1)]) * odz
        d1 = self._cpml("dsxy_dx", d1, X, "_half", s(), I, 2)
        d2 = self._cpml("dsyy_dy", d2, Y, "_half", s(), J, 1)
        d3 = self._cpml("dsyz_dz", d3, Z, "", s(), K, 0)
        f["vy"][s()] = DT_r * (d1 + d2 + d3) + f["vy"][s()]
        I, J, K = (1, NX - 1), (2, NY), (1, km1)
        s = self._rng(I, J, K)
        d1 = (f["sxz"][s(1)] - f["sxz"][s()]) * odx
        d2 = (f["syz"][s()] - f["syz"][s(0, -1)]) * ody
        d3 = (f["szz"][s(0, 0, 1)] - f["szz"][s()]) * odz
        d1 = self._cpml("dsxz_dx", d1, X, "_half", s(), I, 2)
        d2 = self._cpml("dsyz_dy", d2, Y, "", s(), J, 1)
        d3 = self._cpml("dszz_dz", d3, Z, "_half", s(), K, 0)
        f["vz"][s()] = DT_r * (d1 + d2 + d3) + f["vz"][s()]
        if self.ksrc:
            i, j, k = c["isource"], c["jsource"], self.ksrc
            f["vx"][k, j, i] = f["vx"][k, j, i] + c["force_x"][it - 1] * c["deltat"] / c["rho"]
            f["vy"][k, j, i] = f["vy"][k, j, i] + c["force_y"][it - 1] * c["deltat"] / c["rho"]
        for n in ("vx", "vy", "vz"):
            v = f[n]
            v[:, :, 1] = 0.0
            v[:, :, NX] = 0.0
            v[:, 1, :] = 0.0
            v[:, NY, :] = 0.0
            if self.koff == 0:
                v[1] = 0.0
            if self.koff + nzl == self.nz:
                v[nzl] = 0.0

    def step_finish(self, it):
        c, f = self.c, self.f
        if self.ksrc:
            for r in range(len(c["ix_rec"])):
                self.sisvx[r, it - 1] = f["vx"][self.ksrc, c["iy_rec"][r], c["ix_rec"][r]]
                self.sisvy[r, it - 1] = f["vy"][self.ksrc, c["iy_rec"][r], c["ix_rec"][r]]
        P, lam, mu, rho = c["npoints_pml"], c["lam"], c["mu"], c["rho"]
        kmin = P + 1 if self.koff == 0 else 1
        kmax = self.nzl - P if self.koff + self.nzl == self.nz else self.nzl
        EB = (slice(kmin, kmax + 1), slice(P + 1, self.ny - P + 1), slice(P + 1, self.nx - P + 1))
        sxx, syy, szz, sxy, sxz, syz = (f[n][EB] for n in ("sxx", "syy", "szz", "sxy", "sxz", "syz"))
        kin = np.sum(0.5 * rho * (f["vx"][EB] ** 2 + f["vy"][EB] ** 2 + f["vz"][EB] ** 2))
        den = 2.0 * mu * (3.0 * lam + 2.0 * mu)
        exx = (2.0 * (lam + mu) * sxx - lam * syy - lam * szz) / den
        eyy = (2.0 * (lam + mu) * syy - lam * sxx - lam * szz) / den
        pot = np.sum(0.5 * (exx * sxx + eyy * syy + eyy * syy + 2.0 * (sxy / (2.0 * mu)) * sxy
                            + 2.0 * (sxz / (2.0 * mu)) * sxz + 2.0 * (syz / (2.0 * mu)) * syz))
        self.energy[it - 1] = kin + pot
