// 3-D viscoelastic C-PML kernels for sm_100a (fourth order in space, N_SLS = 2).
//
// Two fused kernels per time step replace the five loop nests, the Dirichlet pass and the
// energy pass of seismic_CPML_3D_viscoelastic_MPI.f90:977-1430:
//
//   k_vstress3d    sigmaxx/yy/zz + e1,e11,e22 + sigma*_R (:977-1096), sigmaxy + e12 (:1098-1139),
//                  sigmaxz + e13 and sigmayz + e23 (:1141-1223), and the POTENTIAL part of the
//                  energy (:1401-1419), which only needs the stresses this kernel has just
//                  produced -- so the six sigma*_R arrays are never read again in the step;
//   k_vvelocity3d  vx/vy (:1244-1285), vz (:1287-1308), source (:1310-1335), Dirichlet on two
//                  planes per face (:1337-1371), the KINETIC part of the energy (:1396-1397).
// The step is finished by k_post3d (energy sums + seismogram sample), shared with the
// isotropic solver.
//
// Mapping: one thread per (i,j) column, x across the warp, each block marches a chunk of z
// planes.  The four-plane z windows of the fourth-order operator (vx, vy at k-1..k+2 and vz at
// k-2..k+1 in the stress kernel; sigmaxz, sigmayz at k-2..k+1 and sigmazz at k-1..k+2 in the
// velocity kernel) live in registers, so every plane is fetched from HBM once per kernel; the
// radius-2 in-plane taps are re-read through L1.  The 24 (stress) / 3 (velocity) streamed
// read-modify-write words per point use evict-first loads and stores; both relaxation mechanisms
// of a memory variable sit next to each other (the reference's e(N_SLS,...) layout) and move as
// one 16-byte access.  C-PML memory variables exist only inside the shells (K_MAX_PML = 7 there).
//
// Reference quirk B6: the MPI exchange of the reference delivers only half of the z
// halo its stencils read; the slots never received stay zero.  With `nzl_e` = plane count of one
// reference slab the same taps read zero here: f(k+1) of the backward z differences on the last
// plane of a slab, f(k-1) of the forward z differences on the first plane of a slab.
//
// Compiled with -fmad=false; every division of the reference (/24, /K, /3, /(1 - dt/2 tauinv)) is
// a division by a constant or by a profile value, done by div_exact (cpml_internal.h): the
// correctly rounded quotient from the divisor's reciprocal and two FMA residual corrections, five
// FP64 operations instead of the generic ~60-instruction sequence.  Velocities, stresses and
// memory variables stay bit-identical to an IEEE (non-FMA) build of the reference.
#include "visco_common.cuh"

namespace cpml {

// Unp1 = (Un + deltat*(Sn + 0.5*tauinv*Un)) / (1 - deltat*0.5*tauinv)   (:1003-1009), both mechanisms of
// one memory variable behind one range test
__device__ __forceinline__ double2 evolve2(double2 U, double S0, double S1, const double (&tauinv)[2], const double (&den)[2],
                                           const double (&rden)[2], double dt)
{
    const double t0 = tauinv[0] * U.x, t1 = tauinv[1] * U.y;
    double n0 = U.x + dt * (S0 + 0.5 * t0), n1 = U.y + dt * (S1 + 0.5 * t1);
    div_exact2(n0, n1, den[0], rden[0], den[1], rden[1]);
    return make_double2(n0, n1);
}

template <int NT>
__device__ __forceinline__ double block_sum1(double a, double *smem /* NT/32 */)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
    const int t = threadIdx.y * blockDim.x + threadIdx.x;
    const int w = t >> 5, l = t & 31;
    if (l == 0) smem[w] = a;
    __syncthreads();
    if (w == 0) {
        a = (l < NT / 32) ? smem[l] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
    }
    return a;
}

// Coefficient tables of a block: rows a, b, K, 1/K, a_half, b_half, K_half, 1/K_half; columns /
// rows beyond the grid get the identity (never used).
template <int TX, int TY>
__device__ __forceinline__ void load_coef_tables(const ParamsV3D &p, double (*cxs)[TX], double (*cys)[TY])
{
    const int t = threadIdx.y * TX + threadIdx.x;
    for (int e = t; e < 8 * (TX + TY); e += TX * TY) {
        const bool isx = e < 8 * TX;
        const int m = isx ? e / TX : (e - 8 * TX) / TY;
        const int c = isx ? e % TX : (e - 8 * TX) % TY;
        const int idx = isx ? blockIdx.x * TX + c + 1 : blockIdx.y * TY + c + 1;
        const AxisCoef &a = isx ? p.cx : p.cy;
        const double *src = m == 0 ? a.a : m == 1 ? a.b : m == 2 ? a.K : m == 3 ? a.rK : m == 4 ? a.a_half : m == 5 ? a.b_half
                          : m == 6 ? a.K_half : a.rK_half;
        const double v = idx <= (isx ? p.nx : p.ny) ? src[idx] : ((m == 0 || m == 4) ? 0.0 : 1.0);
        if (isx) cxs[m][c] = v; else cys[m][c] = v;
    }
    __syncthreads();
}

// PART selects what one launch updates: 0 = all six stresses (one launch per step, the default),
// 1 = the normal stresses (sigmaxx/yy/zz, e1, e11, e22, their sigma_R), 2 = the shear stresses
// (sigmaxy/xz/yz, e12, e13, e23, their sigma_R).  Every load is scoped to the nest that uses it,
// which is what lets the fused kernel run in 128 registers (16 warps per SM) without spilling; the
// split launches (CPML_VSPLIT=1) need fewer still but read the velocity planes twice and measured
// 7 % slower on B200 (profiles/r01_v6_visco_sweep.txt).
template <int TX, int TY, int MINB, int PART>
__global__ void __launch_bounds__(TX *TY, MINB)
k_vstress3d(const __grid_constant__ ParamsV3D p)
{
    constexpr bool NORMAL = PART != 2, SHEAR = PART != 1;
    __shared__ double red[TX * TY / 32];
    __shared__ double cxs[8][TX], cys[8][TY];        // a, b, K, 1/K, a_half, b_half, K_half, 1/K_half
    const int i = blockIdx.x * TX + threadIdx.x + 1;
    const int j = blockIdx.y * TY + threadIdx.y + 1;
    double epot = 0.0;
    load_coef_tables<TX, TY>(p, cxs, cys);

    if (i <= p.nx && j <= p.ny) {
        const int kb = 1 + blockIdx.z * p.kchunk;
        const int ke = min(p.nzl, kb + p.kchunk - 1);
        const int pitch = p.pitch;
        const int pl = (int)p.plane;           // < 2^31 elements per field (checked by cpml_create)
        int q = kb * pl + (j - 1) * pitch + (i - 1);

        const bool in_x = (i <= p.xlo) || (i >= p.xhi);
        const bool in_y = (j <= p.ylo) || (j >= p.yhi);
        const int sx = in_x ? vshell(i, p.xlo, p.xhi) : 0;
        const int sy = in_y ? vshell(j, p.ylo, p.yhi) : 0;

        const bool do_n = (i <= p.nx - 1) && (j >= 2);     // :979-980
        const bool do_xy = (i >= 2) && (j <= p.ny - 1);    // :1099-1100
        const bool do_xz = (i >= 2);                       // :1143-1144
        const bool do_yz = (j <= p.ny - 1);                // :1183-1184
        const bool ebox_ij = (i >= p.npml) && (i <= p.nx - p.npml + 1) &&
                             (j >= p.npml) && (j <= p.ny - p.npml + 1);            // :1391-1392

        const double odx = p.odx, ody = p.ody, odz = p.odz, dt = p.dt;
        // x / y C-PML coefficients of this block's columns and rows sit in shared memory (cxs, cys):
        // only shell points read them, at the point of use, so they cost no registers in the interior
        const double *const cxc = &cxs[0][threadIdx.x], *const cyc = &cys[0][threadIdx.y];

        // z windows of the fourth-order operator, carried in registers
        double vx_m = 0, vx_c = 0, vx_p = 0, vy_m = 0, vy_c = 0, vy_p = 0, vz_mm = 0, vz_m = 0, vz_c = 0;
        if (SHEAR) {
            vx_m = p.vx[q - pl]; vx_c = p.vx[q]; vx_p = p.vx[q + pl];
            vy_m = p.vy[q - pl]; vy_c = p.vy[q]; vy_p = p.vy[q + pl];
        }
        if (NORMAL) { vz_mm = p.vz[q - 2 * pl]; vz_m = p.vz[q - pl]; }
        vz_c = p.vz[q];

        int kmod = (kb + p.koff) % p.nzl_e;                  // position inside the emulated reference slab (quirk B6),
                                                              // advanced with the plane instead of a modulo per plane
        for (int k = kb; k <= ke; ++k, q += pl, kmod = (kmod + 1 == p.nzl_e) ? 0 : kmod + 1) {
            const int kg = k + p.koff;                      // :978
            const bool in_z = (kg <= p.zlo) || (kg >= p.zhi);
            const bool in_pml = in_x | in_y | in_z;          // one branch per nest for the interior points
            int qx = 0, qy = 0, qz = 0;
            double az = 0, bz = 0, Kz = 1, azh = 0, bzh = 0, Kzh = 1, rKz = 1, rKzh = 1;
            if (in_x) qx = ((k - 1) * p.ny + (j - 1)) * p.sxp + sx;
            if (in_y) qy = ((k - 1) * p.sy + sy) * pitch + (i - 1);
            if (in_z) {
                qz = ((vshell(kg, p.zlo, p.zhi) - p.zbase) * p.ny + (j - 1)) * pitch + (i - 1);
                if (NORMAL) { az = p.cz.a[kg]; bz = p.cz.b[kg]; Kz = p.cz.K[kg]; rKz = p.cz.rK[kg]; }
                if (SHEAR) { azh = p.cz.a_half[kg]; bzh = p.cz.b_half[kg]; Kzh = p.cz.K_half[kg]; rKzh = p.cz.rK_half[kg]; }
            }
            // quirk B6: taps the reference's MPI exchange never delivers
            const bool cut_up = (kmod == 0);                // last plane of a reference slab
            const bool cut_dn = (kmod == 1);                // first plane of a reference slab
            const bool ebox = ebox_ij && kg >= p.npml && kg <= p.nz - p.npml + 1;      // :1387-1392
            const bool bnd = (k <= 2) | (k >= p.nzl - 1);   // planes a neighbour slab needs (uniform)
            // C-PML memory variables of the NEXT plane (and of the first plane of the chunk) into L2: each
            // recursion is a dependent load -> update -> store, up to three in a row per nest, and the shell
            // points (23 % of the default grid) took 21 % / 31 % of the stall samples of the two kernels with
            // 5 % / 10 % of their instructions (profiles/r01_v8_ncu_cfg5d.txt, source page)
            if ((p.pf & 4) && (in_pml | (kg + 1 <= p.zlo) | (kg + 1 >= p.zhi))) {
                for (int d = (k == kb ? 0 : 1); d <= 1; ++d) {
                    if (k + d > p.nzl) break;
                    if (in_x) {
                        const int qn = ((k + d - 1) * p.ny + (j - 1)) * p.sxp + sx;
                        if (NORMAL) pf_l2(p.mx[0] + qn);
                        if (SHEAR) { pf_l2(p.mx[1] + qn); pf_l2(p.mx[2] + qn); }
                    }
                    if (in_y) {
                        const int qn = ((k + d - 1) * p.sy + sy) * pitch + (i - 1);
                        if (NORMAL) pf_l2(p.my[0] + qn);
                        if (SHEAR) { pf_l2(p.my[1] + qn); pf_l2(p.my[2] + qn); }
                    }
                    const int kgn = kg + d;
                    if ((kgn <= p.zlo) || (kgn >= p.zhi)) {
                        const int qn = ((vshell(kgn, p.zlo, p.zhi) - p.zbase) * p.ny + (j - 1)) * pitch + (i - 1);
                        if (NORMAL) pf_l2(p.mz[0] + qn);
                        if (SHEAR) { pf_l2(p.mz[1] + qn); pf_l2(p.mz[2] + qn); }
                    }
                }
            }

            // L2 prefetch (see pf_l2).  pf = 1: everything plane k+1 streams, at the head of plane k.
            // pf = 2: staggered by half a plane -- the shear words of THIS plane here (used after the
            // normal-stress nest), the normal words of plane k+1 at the head of the shear nests.
            if ((p.pf & 3) == 1 && k + 1 <= p.nzl) {
                const int qn = q + pl;
                if (NORMAL) {
                    pf_l2(p.sxx + qn); pf_l2(p.syy + qn); pf_l2(p.szz + qn);
                    pf_l2(p.rxx + qn); pf_l2(p.ryy + qn); pf_l2(p.rzz + qn);
                    pf_l2(p.e1 + qn); pf_l2(p.e11 + qn); pf_l2(p.e22 + qn);
                }
                if (SHEAR) {
                    pf_l2(p.sxy + qn); pf_l2(p.sxz + qn); pf_l2(p.syz + qn);
                    pf_l2(p.rxy + qn); pf_l2(p.rxz + qn); pf_l2(p.ryz + qn);
                    pf_l2(p.e12 + qn); pf_l2(p.e13 + qn); pf_l2(p.e23 + qn);
                    pf_l2(p.vx + qn + 2 * pl); pf_l2(p.vy + qn + 2 * pl);
                }
                pf_l2(p.vz + qn + pl);
            }
            if ((p.pf & 3) == 2 && SHEAR) {
                pf_l2(p.sxy + q); pf_l2(p.sxz + q); pf_l2(p.syz + q);
                pf_l2(p.rxy + q); pf_l2(p.rxz + q); pf_l2(p.ryz + q);
                pf_l2(p.e12 + q); pf_l2(p.e13 + q); pf_l2(p.e23 + q);
                pf_l2(p.vx + q + 2 * pl); pf_l2(p.vy + q + 2 * pl);
            }
            double vz_p = 0.0;
            if (NORMAL) vz_p = p.vz[q + pl];

            // ---- sigmaxx, sigmayy, sigmazz, e1, e11, e22, sigma*_R  (:977-1096)
            if (NORMAL) {
                const double vx_im1 = p.vx[q - 1], vx_0 = SHEAR ? vx_c : p.vx[q], vx_ip1 = p.vx[q + 1], vx_ip2 = p.vx[q + 2];
                const double vy_jm2 = p.vy[q - 2 * pitch], vy_jm1 = p.vy[q - pitch], vy_0 = SHEAR ? vy_c : p.vy[q], vy_jp1 = p.vy[q + pitch];
                double sxx = vld(p.sxx + q), syy = vld(p.syy + q), szz = vld(p.szz + q);
                double rxx = vld(p.rxx + q), ryy = vld(p.ryy + q), rzz = vld(p.rzz + q);
                double2 e1 = vld2(p.e1 + q), e11 = vld2(p.e11 + q), e22 = vld2(p.e22 + q);
                if (do_n && kg >= 2) {                      // k2begin, :942-943
                    double duxdx = d4n(vx_ip1, vx_0, vx_ip2, vx_im1, odx);
                    double duydy = d4n(vy_0, vy_jm1, vy_jp1, vy_jm2, ody);
                    double duzdz = d4n(vz_c, vz_m, cut_up ? 0.0 : vz_p, vz_mm, odz);
                    DIV24_3(duxdx, duydy, duzdz);
                    if (in_pml) {
                        if (in_x) duxdx = vcpml(p.mx[0], qx, cxc[5 * TX], cxc[4 * TX], cxc[6 * TX], cxc[7 * TX], duxdx);
                        if (in_y) duydy = vcpml(p.my[0], qy, cyc[1 * TY], cyc[0 * TY], cyc[2 * TY], cyc[3 * TY], duydy);
                        if (in_z) duzdz = vcpml(p.mz[0], qz, bz, az, Kz, rKz, duzdz);
                    }
                    const double div = duxdx + duydy + duzdz;
                    const double div3 = div_small(div, 3.0, 1.0 / 3.0);      // div/DIM

                    e1 = evolve2(e1, div * p.phi1[0], div * p.phi1[1], p.tauinv1, p.den1, p.rden1, dt);
                    e11 = evolve2(e11, (duxdx - div3) * p.phi2[0], (duxdx - div3) * p.phi2[1], p.tauinv2, p.den2, p.rden2, dt);
                    e22 = evolve2(e22, (duydy - div3) * p.phi2[0], (duydy - div3) * p.phi2[1], p.tauinv2, p.den2, p.rden2, dt);
                    vst2(p.e1 + q, e1); vst2(p.e11 + q, e11); vst2(p.e22 + q, e22);

                    // relaxed moduli times the memory variables (:1054-1060)
                    sxx = sxx + dt * (p.lam23mu * (e1.x + e1.y) + p.two_mu * (e11.x + e11.y));
                    syy = syy + dt * (p.lam23mu * (e1.x + e1.y) + p.two_mu * (e22.x + e22.y));
                    szz = szz + dt * (p.szz_e1 * (e1.x + e1.y) - p.szz_dev * (e11.x + e11.y + e22.x + e22.y));   // quirk B14
                    // unrelaxed elastic term (:1064-1077)
                    sxx = sxx + (p.l2m_u * duxdx + p.lam_u * duydy + p.lam_u * duzdz) * dt;
                    syy = syy + (p.lam_u * duxdx + p.l2m_u * duydy + p.lam_u * duzdz) * dt;
                    szz = szz + (p.lam_u * duxdx + p.lam_u * duydy + p.l2m_u * duzdz) * dt;
                    // relaxed stresses (:1079-1092)
                    rxx = rxx + (p.l2m_r * duxdx + p.lam * duydy + p.lam * duzdz) * dt;
                    ryy = ryy + (p.lam * duxdx + p.l2m_r * duydy + p.lam * duzdz) * dt;
                    rzz = rzz + (p.lam * duxdx + p.lam * duydy + p.l2m_r * duzdz) * dt;
                    vst(p.sxx + q, sxx); vst(p.syy + q, syy); vst(p.szz + q, szz);
                    vst(p.rxx + q, rxx); vst(p.ryy + q, ryy); vst(p.rzz + q, rzz);
                }
                if (bnd) peer_put<true>(p.peer_lo[2], p.peer_hi[2], k, p.nzl, q - k * pl, pl, szz);
                // potential energy over the PML-free box (:1387-1419), normal-stress terms; quirk B2:
                // epsilon_yy * sigmayy_R is counted twice and the zz term is missing
                if (ebox) {
                    const double epsilon_xx = (p.c2lm * sxx - p.lam * syy - p.lam * szz) * p.inv_den;
                    const double epsilon_yy = (p.c2lm * syy - p.lam * sxx - p.lam * szz) * p.inv_den;
                    epot += 0.5 * (epsilon_xx * rxx + epsilon_yy * ryy + epsilon_yy * ryy);
                }
            }

            if (SHEAR) {
                if ((p.pf & 3) == 2 && NORMAL && k + 1 <= p.nzl) {
                    const int qn = q + pl;
                    pf_l2(p.sxx + qn); pf_l2(p.syy + qn); pf_l2(p.szz + qn);
                    pf_l2(p.rxx + qn); pf_l2(p.ryy + qn); pf_l2(p.rzz + qn);
                    pf_l2(p.e1 + qn); pf_l2(p.e11 + qn); pf_l2(p.e22 + qn);
                    pf_l2(p.vz + qn + pl);
                }
                const double vx_pp = p.vx[q + 2 * pl], vy_pp = p.vy[q + 2 * pl];
                double esh = 0.0;
                // ---- sigmaxy, e12  (:1098-1139)
                {
                    const double vy_im2 = p.vy[q - 2], vy_im1 = p.vy[q - 1], vy_ip1 = p.vy[q + 1];
                    const double vx_jm1 = p.vx[q - pitch], vx_jp1 = p.vx[q + pitch], vx_jp2 = p.vx[q + 2 * pitch];
                    double sxy = vld(p.sxy + q), rxy = vld(p.rxy + q);
                    double2 e12 = vld2(p.e12 + q);
                    if (do_xy) {
                        double duydx = d4n(vy_c, vy_im1, vy_ip1, vy_im2, odx);
                        double duxdy = d4n(vx_jp1, vx_c, vx_jp2, vx_jm1, ody);
                        DIV24_2(duydx, duxdy);
                        if (in_pml) {
                            if (in_x) duydx = vcpml(p.mx[1], qx, cxc[1 * TX], cxc[0 * TX], cxc[2 * TX], cxc[3 * TX], duydx);
                            if (in_y) duxdy = vcpml(p.my[1], qy, cyc[5 * TY], cyc[4 * TY], cyc[6 * TY], cyc[7 * TY], duxdy);
                        }
                        const double g = duxdy + duydx;
                        e12 = evolve2(e12, g * p.phi2[0], g * p.phi2[1], p.tauinv2, p.den2, p.rden2, dt);
                        vst2(p.e12 + q, e12);
                        sxy = sxy + dt * p.mu * (e12.x + e12.y);
                        sxy = sxy + p.mu_u * g * dt;
                        rxy = rxy + p.mu * g * dt;
                        vst(p.sxy + q, sxy); vst(p.rxy + q, rxy);
                    }
                    esh += 2.0 * (rxy * p.inv_2mu) * rxy;
                }
                // ---- sigmaxz, e13 and sigmayz, e23  (:1141-1223)
                {
                    const double vz_im2 = p.vz[q - 2], vz_im1 = p.vz[q - 1], vz_ip1 = p.vz[q + 1];
                    double sxz = vld(p.sxz + q), rxz = vld(p.rxz + q);
                    double2 e13 = vld2(p.e13 + q);
                    if (do_xz && kg <= p.nz - 1) {                  // kminus1end, :945-946
                        double duzdx = d4n(vz_c, vz_im1, vz_ip1, vz_im2, odx);
                        double duxdz = d4n(vx_p, vx_c, vx_pp, cut_dn ? 0.0 : vx_m, odz);
                        DIV24_2(duzdx, duxdz);
                        if (in_pml) {
                            if (in_x) duzdx = vcpml(p.mx[2], qx, cxc[1 * TX], cxc[0 * TX], cxc[2 * TX], cxc[3 * TX], duzdx);
                            if (in_z) duxdz = vcpml(p.mz[1], qz, bzh, azh, Kzh, rKzh, duxdz);
                        }
                        const double g = duxdz + duzdx;
                        e13 = evolve2(e13, g * p.phi2[0], g * p.phi2[1], p.tauinv2, p.den2, p.rden2, dt);
                        vst2(p.e13 + q, e13);
                        sxz = sxz + dt * p.mu * (e13.x + e13.y);
                        sxz = sxz + p.mu_u * g * dt;
                        rxz = rxz + p.mu * g * dt;
                        vst(p.sxz + q, sxz); vst(p.rxz + q, rxz);
                    }
                    if (bnd) peer_put<false>(p.peer_lo[4], p.peer_hi[4], k, p.nzl, q - k * pl, pl, sxz);
                    esh += 2.0 * (rxz * p.inv_2mu) * rxz;
                }
                {
                    const double vz_jm1 = p.vz[q - pitch], vz_jp1 = p.vz[q + pitch], vz_jp2 = p.vz[q + 2 * pitch];
                    double syz = vld(p.syz + q), ryz = vld(p.ryz + q);
                    double2 e23 = vld2(p.e23 + q);
                    if (do_yz && kg <= p.nz - 1) {
                        double duzdy = d4n(vz_jp1, vz_c, vz_jp2, vz_jm1, ody);
                        double duydz = d4n(vy_p, vy_c, vy_pp, cut_dn ? 0.0 : vy_m, odz);
                        DIV24_2(duzdy, duydz);
                        if (in_pml) {
                            if (in_y) duzdy = vcpml(p.my[2], qy, cyc[5 * TY], cyc[4 * TY], cyc[6 * TY], cyc[7 * TY], duzdy);
                            if (in_z) duydz = vcpml(p.mz[2], qz, bzh, azh, Kzh, rKzh, duydz);
                        }
                        const double g = duydz + duzdy;
                        e23 = evolve2(e23, g * p.phi2[0], g * p.phi2[1], p.tauinv2, p.den2, p.rden2, dt);
                        vst2(p.e23 + q, e23);
                        syz = syz + dt * p.mu * (e23.x + e23.y);
                        syz = syz + p.mu_u * g * dt;
                        ryz = ryz + p.mu * g * dt;
                        vst(p.syz + q, syz); vst(p.ryz + q, ryz);
                    }
                    if (bnd) peer_put<false>(p.peer_lo[5], p.peer_hi[5], k, p.nzl, q - k * pl, pl, syz);
                    esh += 2.0 * (ryz * p.inv_2mu) * ryz;
                }
                // potential energy, shear terms (:1411-1419)
                if (ebox) epot += 0.5 * esh;
                vx_m = vx_c; vx_c = vx_p; vx_p = vx_pp;
                vy_m = vy_c; vy_c = vy_p; vy_p = vy_pp;
            }
            if (NORMAL) { vz_mm = vz_m; vz_m = vz_c; vz_c = vz_p; }
            else vz_c = p.vz[q + pl];
        }
    }

    epot = block_sum1<TX * TY>(epot, red);
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        const int b = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        p.partials[(PART == 2 ? 2 : 1) * p.nblocks + b] = epot;      // [nb,2nb): normal (or all), [2nb,3nb): shear
    }
}

template <int TX, int TY, int MINB>
__global__ void __launch_bounds__(TX *TY, MINB)
k_vvelocity3d(const __grid_constant__ ParamsV3D p)
{
    __shared__ double red[TX * TY / 32];
    __shared__ double cxs[8][TX], cys[8][TY];
    const int i = blockIdx.x * TX + threadIdx.x + 1;
    const int j = blockIdx.y * TY + threadIdx.y + 1;
    double ekin = 0.0;
    load_coef_tables<TX, TY>(p, cxs, cys);

    if (i <= p.nx && j <= p.ny) {
        const int kb = 1 + blockIdx.z * p.kchunk;
        const int ke = min(p.nzl, kb + p.kchunk - 1);
        const int pitch = p.pitch;
        const int pl = (int)p.plane;           // < 2^31 elements per field (checked by cpml_create)
        int q = kb * pl + (j - 1) * pitch + (i - 1);

        const bool in_x = (i <= p.xlo) || (i >= p.xhi);
        const bool in_y = (j <= p.ylo) || (j >= p.yhi);
        const int sx = in_x ? vshell(i, p.xlo, p.xhi) : 0;
        const int sy = in_y ? vshell(j, p.ylo, p.yhi) : 0;

        const bool do_vx = (i >= 2) && (j >= 2);                    // :1246-1247
        const bool do_vy = (i <= p.nx - 1) && (j <= p.ny - 1);      // :1266-1267
        const bool do_vz = (i <= p.nx - 1) && (j >= 2);             // :1289-1290
        const bool edge_ij = (i <= 1) || (i >= p.nx) || (j <= 1) || (j >= p.ny);   // :1340-1358 (two cells per face)
        const bool ebox_ij = (i >= p.npml) && (i <= p.nx - p.npml + 1) &&
                             (j >= p.npml) && (j <= p.ny - p.npml + 1);
        const bool src_ij = (i == p.isrc) && (j == p.jsrc);

        const double odx = p.odx, ody = p.ody, odz = p.odz, dt_r = p.dt_over_rho;

        // x / y C-PML coefficients of this block's columns and rows sit in shared memory (cxs, cys):
        // only shell points read them, at the point of use, so they cost no registers in the interior
        const double *const cxc = &cxs[0][threadIdx.x], *const cyc = &cys[0][threadIdx.y];

        // z windows
        double sxz_mm = p.sxz[q - 2 * pl], sxz_m = p.sxz[q - pl], sxz_c = p.sxz[q];
        double syz_mm = p.syz[q - 2 * pl], syz_m = p.syz[q - pl], syz_c = p.syz[q];
        double szz_m = p.szz[q - pl], szz_c = p.szz[q], szz_p = p.szz[q + pl];

        int kmod = (kb + p.koff) % p.nzl_e;                  // see k_vstress3d
        for (int k = kb; k <= ke; ++k, q += pl, kmod = (kmod + 1 == p.nzl_e) ? 0 : kmod + 1) {
            const int kg = k + p.koff;
            // ---- loads of this plane
            const double sxz_p = p.sxz[q + pl], syz_p = p.syz[q + pl], szz_pp = p.szz[q + 2 * pl];
            const double sxx_im2 = p.sxx[q - 2], sxx_im1 = p.sxx[q - 1], sxx_c = p.sxx[q], sxx_ip1 = p.sxx[q + 1];
            const double sxy_jm2 = p.sxy[q - 2 * pitch], sxy_jm1 = p.sxy[q - pitch], sxy_c = p.sxy[q], sxy_jp1 = p.sxy[q + pitch];
            const double sxy_im1 = p.sxy[q - 1], sxy_ip1 = p.sxy[q + 1], sxy_ip2 = p.sxy[q + 2];
            const double syy_jm1 = p.syy[q - pitch], syy_c = p.syy[q], syy_jp1 = p.syy[q + pitch], syy_jp2 = p.syy[q + 2 * pitch];
            const double sxz_im1 = p.sxz[q - 1], sxz_ip1 = p.sxz[q + 1], sxz_ip2 = p.sxz[q + 2];
            const double syz_jm2 = p.syz[q - 2 * pitch], syz_jm1 = p.syz[q - pitch], syz_jp1 = p.syz[q + pitch];
            double vx = vld(p.vx + q), vy = vld(p.vy + q), vz = vld(p.vz + q);

            const bool in_z = (kg <= p.zlo) || (kg >= p.zhi);
            const bool in_pml = in_x | in_y | in_z;          // one branch per nest for the interior points
            int qx = 0, qy = 0, qz = 0;
            double az = 0, bz = 0, Kz = 1, azh = 0, bzh = 0, Kzh = 1, rKz = 1, rKzh = 1;
            if (in_x) qx = ((k - 1) * p.ny + (j - 1)) * p.sxp + sx;
            if (in_y) qy = ((k - 1) * p.sy + sy) * pitch + (i - 1);
            if (in_z) {
                qz = ((vshell(kg, p.zlo, p.zhi) - p.zbase) * p.ny + (j - 1)) * pitch + (i - 1);
                az = p.cz.a[kg]; bz = p.cz.b[kg]; Kz = p.cz.K[kg];
                azh = p.cz.a_half[kg]; bzh = p.cz.b_half[kg]; Kzh = p.cz.K_half[kg];
                rKz = p.cz.rK[kg]; rKzh = p.cz.rK_half[kg];
            }
            const bool cut_up = (kmod == 0);
            const bool cut_dn = (kmod == 1);
            if ((p.pf & 4) && (in_pml | (kg + 1 <= p.zlo) | (kg + 1 >= p.zhi))) {      // memory variables of the next plane into L2, see k_vstress3d
                for (int d = (k == kb ? 0 : 1); d <= 1; ++d) {
                    if (k + d > p.nzl) break;
                    if (in_x) {
                        const int qn = ((k + d - 1) * p.ny + (j - 1)) * p.sxp + sx;
                        pf_l2(p.mx[3] + qn); pf_l2(p.mx[4] + qn); pf_l2(p.mx[5] + qn);
                    }
                    if (in_y) {
                        const int qn = ((k + d - 1) * p.sy + sy) * pitch + (i - 1);
                        pf_l2(p.my[3] + qn); pf_l2(p.my[4] + qn); pf_l2(p.my[5] + qn);
                    }
                    const int kgn = kg + d;
                    if ((kgn <= p.zlo) || (kgn >= p.zhi)) {
                        const int qn = ((vshell(kgn, p.zlo, p.zhi) - p.zbase) * p.ny + (j - 1)) * pitch + (i - 1);
                        pf_l2(p.mz[3] + qn); pf_l2(p.mz[4] + qn); pf_l2(p.mz[5] + qn);
                    }
                }
            }

            if (kg >= 2) {                                           // k2begin
                if (do_vx) {                                         // :1244-1262
                    double d1 = d4n(sxx_c, sxx_im1, sxx_ip1, sxx_im2, odx);
                    double d2 = d4n(sxy_c, sxy_jm1, sxy_jp1, sxy_jm2, ody);
                    double d3 = d4n(sxz_c, sxz_m, cut_up ? 0.0 : sxz_p, sxz_mm, odz);
                    DIV24_3(d1, d2, d3);
                    if (in_pml) {
                        if (in_x) d1 = vcpml(p.mx[3], qx, cxc[1 * TX], cxc[0 * TX], cxc[2 * TX], cxc[3 * TX], d1);
                        if (in_y) d2 = vcpml(p.my[3], qy, cyc[1 * TY], cyc[0 * TY], cyc[2 * TY], cyc[3 * TY], d2);
                        if (in_z) d3 = vcpml(p.mz[3], qz, bz, az, Kz, rKz, d3);
                    }
                    vx = dt_r * (d1 + d2 + d3) + vx;
                }
                if (do_vy) {                                         // :1266-1284
                    double d1 = d4n(sxy_ip1, sxy_c, sxy_ip2, sxy_im1, odx);
                    double d2 = d4n(syy_jp1, syy_c, syy_jp2, syy_jm1, ody);
                    double d3 = d4n(syz_c, syz_m, cut_up ? 0.0 : syz_p, syz_mm, odz);
                    DIV24_3(d1, d2, d3);
                    if (in_pml) {
                        if (in_x) d1 = vcpml(p.mx[4], qx, cxc[5 * TX], cxc[4 * TX], cxc[6 * TX], cxc[7 * TX], d1);
                        if (in_y) d2 = vcpml(p.my[4], qy, cyc[5 * TY], cyc[4 * TY], cyc[6 * TY], cyc[7 * TY], d2);
                        if (in_z) d3 = vcpml(p.mz[4], qz, bz, az, Kz, rKz, d3);
                    }
                    vy = dt_r * (d1 + d2 + d3) + vy;
                }
            }
            if (do_vz && kg <= p.nz - 1) {                           // kminus1end, :1287-1308
                double d1 = d4n(sxz_ip1, sxz_c, sxz_ip2, sxz_im1, odx);
                double d2 = d4n(syz_c, syz_jm1, syz_jp1, syz_jm2, ody);
                double d3 = d4n(szz_p, szz_c, szz_pp, cut_dn ? 0.0 : szz_m, odz);
                DIV24_3(d1, d2, d3);
                if (in_pml) {
                    if (in_x) d1 = vcpml(p.mx[5], qx, cxc[5 * TX], cxc[4 * TX], cxc[6 * TX], cxc[7 * TX], d1);
                    if (in_y) d2 = vcpml(p.my[5], qy, cyc[1 * TY], cyc[0 * TY], cyc[2 * TY], cyc[3 * TY], d2);
                    if (in_z) d3 = vcpml(p.mz[5], qz, bzh, azh, Kzh, rKzh, d3);
                }
                vz = dt_r * (d1 + d2 + d3) + vz;
            }

            // source (:1332-1333), after the update of step it and before Dirichlet
            if (src_ij && k == p.ksrc) {
                vx = vx + p.src_x[p.it - 1];
                vy = vy + p.src_y[p.it - 1];
            }
            // Dirichlet, two planes per face (:1337-1371); the ghost cells i,j = 0, N+1 and the
            // outer halo planes are never written and stay zero
            if (edge_ij || kg <= 1 || kg >= p.nz) { vx = 0.0; vy = 0.0; vz = 0.0; }

            vst(p.vx + q, vx);
            vst(p.vy + q, vy);
            vst(p.vz + q, vz);
            if ((k <= 2) | (k >= p.nzl - 1)) {              // planes a neighbour slab needs (:962-975)
                const int rel = q - k * pl;
                peer_put<true>(p.peer_lo[0], p.peer_hi[0], k, p.nzl, rel, pl, vx);
                peer_put<true>(p.peer_lo[1], p.peer_hi[1], k, p.nzl, rel, pl, vy);
                peer_put<false>(p.peer_lo[3], p.peer_hi[3], k, p.nzl, rel, pl, vz);
            }

            // kinetic energy over the PML-free box (:1387-1397)
            if (ebox_ij && kg >= p.npml && kg <= p.nz - p.npml + 1)
                ekin += p.half_rho * (vx * vx + vy * vy + vz * vz);

            sxz_mm = sxz_m; sxz_m = sxz_c; sxz_c = sxz_p;
            syz_mm = syz_m; syz_m = syz_c; syz_c = syz_p;
            szz_m = szz_c; szz_c = szz_p; szz_p = szz_pp;
        }
    }

    ekin = block_sum1<TX * TY>(ekin, red);
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        const int b = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        p.partials[b] = ekin;
    }
}

// ---- launch dispatch ---------------------------------------------------------------

static int g_vtx = 0, g_vty = 0, g_vsplit = 0, g_vminb = 0, g_vpf = 0;

void visco_tile(int *tx, int *ty)
{
    if (!g_vtx) {
        const char *sx = getenv("CPML_VTX"), *sy = getenv("CPML_VTY"), *sp = getenv("CPML_VSPLIT"), *sm = getenv("CPML_VMINB");
        const char *sf = getenv("CPML_VPF");
        g_vpf = sf ? atoi(sf) : 6;            // L2 prefetch mode of ParamsV3D::pf
        if (g_vpf < 0 || g_vpf > 6 || (g_vpf & 3) == 3) g_vpf = 6;
        g_vtx = sx ? atoi(sx) : 32;
        g_vty = sy ? atoi(sy) : 8;
        g_vsplit = sp ? atoi(sp) : 0;         // 0: one stress launch per step (default), 1: normal and shear stresses in two launches
        g_vminb = sm ? atoi(sm) : 0;          // 0: the default register cap of the tile
        const int key = g_vtx * 100 + g_vty;
        if (key != 3208 && key != 3204 && key != 6404 && key != 6402 && key != 1616) { g_vtx = 32; g_vty = 8; }
    }
    *tx = g_vtx; *ty = g_vty;
}

int visco_stress_launches() { int a, b; visco_tile(&a, &b); return g_vsplit ? 2 : 1; }

// MINB = resident blocks per SM the kernels are compiled for: 65536 / (threads * MINB) registers per thread
template <int TX, int TY, int MINB>
static cudaError_t vlaunch(const ParamsV3D &p, dim3 grid, cudaStream_t s, bool stress)
{
    if (!stress) k_vvelocity3d<TX, TY, MINB><<<grid, dim3(TX, TY), 0, s>>>(p);
    else if (g_vsplit) {
        k_vstress3d<TX, TY, MINB, 1><<<grid, dim3(TX, TY), 0, s>>>(p);
        k_vstress3d<TX, TY, MINB, 2><<<grid, dim3(TX, TY), 0, s>>>(p);
    } else k_vstress3d<TX, TY, MINB, 0><<<grid, dim3(TX, TY), 0, s>>>(p);
    return cudaGetLastError();
}

template <int TX, int TY, int LO, int HI>
static cudaError_t vlaunch_minb(const ParamsV3D &p, dim3 grid, cudaStream_t s, bool stress)
{
    const bool hi = g_vminb ? (g_vminb >= HI) : true;
    return hi ? vlaunch<TX, TY, HI>(p, grid, s, stress) : vlaunch<TX, TY, LO>(p, grid, s, stress);
}

static cudaError_t vdispatch(const ParamsV3D &p_in, dim3 grid, cudaStream_t s, bool stress)
{
    int tx, ty;
    visco_tile(&tx, &ty);
    ParamsV3D p = p_in;
    p.pf = g_vpf;
    switch (tx * 100 + ty) {
    case 3204:  return g_vminb == 3 ? vlaunch<32, 4, 3>(p, grid, s, stress) : vlaunch_minb<32, 4, 2, 4>(p, grid, s, stress);
    case 6404:  return vlaunch_minb<64, 4, 1, 2>(p, grid, s, stress);
    case 6402:  return vlaunch_minb<64, 2, 2, 4>(p, grid, s, stress);
    case 1616:  return vlaunch_minb<16, 16, 1, 2>(p, grid, s, stress);
    default:    return vlaunch_minb<32, 8, 1, 2>(p, grid, s, stress);
    }
}

void launch_vstress3d(const ParamsV3D &p, dim3 grid, cudaStream_t s) { vdispatch(p, grid, s, true); }
void launch_vvelocity3d(const ParamsV3D &p, dim3 grid, cudaStream_t s) { vdispatch(p, grid, s, false); }

}  // namespace cpml
