// 2-D isotropic C-PML kernels for sm_100a, second and fourth order in space.
//
//   k_stress2d<ORDER>    sigma_xx/yy (2D-2nd :556-580, 2D-4th :557-581),
//                        sigma_xy    (2D-2nd :582-600, 2D-4th :583-601)
//   k_velocity2d<ORDER>  vx, vy (2D-2nd :606-641), source (:643-667), Dirichlet
//                        (:669-680), per-block kinetic / potential energy (:688-713)
// The step is finished by k_post3d (energy sum + seismogram sample), shared with 3-D.
//
// One thread per grid point, x across the warp; the radius-1 (2nd order) or radius-2
// (4th order) neighbours come through L1/L2.  Heterogeneous lambda, mu, rho arrays are
// streamed and averaged on the fly exactly as the reference does.  Fields carry a
// two-cell zero ghost ring (the reference's fourth-order arrays are (0:NX+1,0:NY+1)),
// so the fourth-order taps at the edges read the same zeros the reference reads.
// Compiled with -fmad=false, divisions kept as divisions: bit-identical fields.
#include "kernels_2d_point.cuh"

namespace cpml {

// The reference divides by DELTAX (2D-2nd :564) or by 24*DELTAX (2D-4th :565): `den` is that
// divisor and `rden` its correctly rounded reciprocal; div_exact returns the correctly rounded
// quotient, bit-identical to the division (cpml_internal.h).
// forward difference u(n+1)-u(n): `s` is the element stride along the axis
template <int ORDER>
__device__ __forceinline__ double d_fwd(const double *f, long long q, long long s, double den, double rden)
{
    if (ORDER == 2) return div_exact(f[q + s] - f[q], den, rden);
    return div_exact(27.0 * f[q + s] - 27.0 * f[q] - f[q + 2 * s] + f[q - s], den, rden);
}
// backward difference u(n)-u(n-1)
template <int ORDER>
__device__ __forceinline__ double d_bwd(const double *f, long long q, long long s, double den, double rden)
{
    if (ORDER == 2) return div_exact(f[q] - f[q - s], den, rden);
    return div_exact(27.0 * f[q] - 27.0 * f[q - s] - f[q + s] + f[q - 2 * s], den, rden);
}

template <int ORDER, int TX, int TY>
__global__ void __launch_bounds__(TX *TY)
k_stress2d(const __grid_constant__ Params2D p)
{
    const int i = blockIdx.x * TX + threadIdx.x + 1;
    const int j = blockIdx.y * TY + threadIdx.y + 1;
    if (i > p.nx || j > p.ny) return;
    const int pitch = p.pitch;
    const long long q = (long long)(j - 1) * pitch + (i - 1);
    const bool in_x = (i <= p.xlo) || (i >= p.xhi);
    const bool in_y = (j <= p.ylo) || (j >= p.yhi);
    const long long qx = in_x ? (long long)(j - 1) * p.sxp + shell_index2(i, p.xlo, p.xhi) : 0;
    const long long qy = in_y ? (long long)shell_index2(j, p.ylo, p.yhi) * pitch + (i - 1) : 0;
    const double DELTAT = p.deltat;

    if (i <= p.nx - 1 && j >= 2) {
        const double lambda_half_x = 0.5 * (p.lambda[q + 1] + p.lambda[q]);
        const double mu_half_x = 0.5 * (p.mu[q + 1] + p.mu[q]);
        const double lambda_plus_two_mu_half_x = lambda_half_x + 2.0 * mu_half_x;
        double value_dvx_dx = d_fwd<ORDER>(p.vx, q, 1, p.denx, p.rdenx);
        double value_dvy_dy = d_bwd<ORDER>(p.vy, q, pitch, p.deny, p.rdeny);
        if (in_x) value_dvx_dx = cpml_apply2(p.mx[0], qx, p.cx.b_half[i], p.cx.a_half[i], p.cx.K_half[i], p.cx.rK_half[i], value_dvx_dx);
        if (in_y) value_dvy_dy = cpml_apply2(p.my[0], qy, p.cy.b[j], p.cy.a[j], p.cy.K[j], p.cy.rK[j], value_dvy_dy);
        p.sxx[q] = p.sxx[q] + (lambda_plus_two_mu_half_x * value_dvx_dx + lambda_half_x * value_dvy_dy) * DELTAT;
        p.syy[q] = p.syy[q] + (lambda_half_x * value_dvx_dx + lambda_plus_two_mu_half_x * value_dvy_dy) * DELTAT;
    }
    if (i >= 2 && j <= p.ny - 1) {
        const double mu_half_y = 0.5 * (p.mu[q + pitch] + p.mu[q]);
        double value_dvy_dx = d_bwd<ORDER>(p.vy, q, 1, p.denx, p.rdenx);
        double value_dvx_dy = d_fwd<ORDER>(p.vx, q, pitch, p.deny, p.rdeny);
        if (in_x) value_dvy_dx = cpml_apply2(p.mx[1], qx, p.cx.b[i], p.cx.a[i], p.cx.K[i], p.cx.rK[i], value_dvy_dx);
        // quirk B3: the fourth-order program divides by K_y(j) here (2D-4th :596), the
        // second-order one by K_y_half(j) (2D-2nd :595)
        if (in_y) value_dvx_dy = cpml_apply2(p.my[1], qy, p.cy.b_half[j], p.cy.a_half[j],
                                             ORDER == 4 ? p.cy.K[j] : p.cy.K_half[j],
                                             ORDER == 4 ? p.cy.rK[j] : p.cy.rK_half[j], value_dvx_dy);
        p.sxy[q] = p.sxy[q] + mu_half_y * (value_dvy_dx + value_dvx_dy) * DELTAT;
    }
}

template <int NT>
__device__ __forceinline__ void block_sum2_2d(double &a, double &b, double *smem)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_down_sync(0xffffffffu, a, o);
        b += __shfl_down_sync(0xffffffffu, b, o);
    }
    const int t = threadIdx.y * blockDim.x + threadIdx.x;
    const int w = t >> 5, l = t & 31;
    if (l == 0) { smem[w] = a; smem[NT / 32 + w] = b; }
    __syncthreads();
    if (w == 0) {
        a = (l < NT / 32) ? smem[l] : 0.0;
        b = (l < NT / 32) ? smem[NT / 32 + l] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_down_sync(0xffffffffu, a, o);
            b += __shfl_down_sync(0xffffffffu, b, o);
        }
    }
}

template <int ORDER, int TX, int TY>
__global__ void __launch_bounds__(TX *TY)
k_velocity2d(const __grid_constant__ Params2D p)
{
    __shared__ double red[2 * TX * TY / 32];
    const int i = blockIdx.x * TX + threadIdx.x + 1;
    const int j = blockIdx.y * TY + threadIdx.y + 1;
    double ekin = 0.0, epot = 0.0;

    if (i <= p.nx && j <= p.ny) {
        const int pitch = p.pitch;
        const long long q = (long long)(j - 1) * pitch + (i - 1);
        const bool in_x = (i <= p.xlo) || (i >= p.xhi);
        const bool in_y = (j <= p.ylo) || (j >= p.yhi);
        const long long qx = in_x ? (long long)(j - 1) * p.sxp + shell_index2(i, p.xlo, p.xhi) : 0;
        const long long qy = in_y ? (long long)shell_index2(j, p.ylo, p.yhi) * pitch + (i - 1) : 0;
        const double DELTAT = p.deltat;
        const double rho = p.rho[q];
        const double rho_half_x_half_y = 0.25 * (rho + p.rho[q + 1] + p.rho[q + 1 + pitch] + p.rho[q + pitch]);
        double vx = p.vx[q], vy = p.vy[q];

        if (i >= 2 && j >= 2) {
            double value_dsigmaxx_dx = d_bwd<ORDER>(p.sxx, q, 1, p.denx, p.rdenx);
            double value_dsigmaxy_dy = d_bwd<ORDER>(p.sxy, q, pitch, p.deny, p.rdeny);
            if (in_x) value_dsigmaxx_dx = cpml_apply2(p.mx[2], qx, p.cx.b[i], p.cx.a[i], p.cx.K[i], p.cx.rK[i], value_dsigmaxx_dx);
            if (in_y) value_dsigmaxy_dy = cpml_apply2(p.my[2], qy, p.cy.b[j], p.cy.a[j], p.cy.K[j], p.cy.rK[j], value_dsigmaxy_dy);
            vx = vx + div_rho((value_dsigmaxx_dx + value_dsigmaxy_dy) * DELTAT, rho, p.rho_exact);
        }
        if (i <= p.nx - 1 && j <= p.ny - 1) {
            double value_dsigmaxy_dx = d_fwd<ORDER>(p.sxy, q, 1, p.denx, p.rdenx);
            double value_dsigmayy_dy = d_fwd<ORDER>(p.syy, q, pitch, p.deny, p.rdeny);
            if (in_x) value_dsigmaxy_dx = cpml_apply2(p.mx[3], qx, p.cx.b_half[i], p.cx.a_half[i], p.cx.K_half[i], p.cx.rK_half[i], value_dsigmaxy_dx);
            if (in_y) value_dsigmayy_dy = cpml_apply2(p.my[3], qy, p.cy.b_half[j], p.cy.a_half[j], p.cy.K_half[j], p.cy.rK_half[j], value_dsigmayy_dy);
            vy = vy + div_rho((value_dsigmaxy_dx + value_dsigmayy_dy) * DELTAT, rho_half_x_half_y, p.rho_exact);
        }
        if (i == p.isrc && j == p.jsrc) {               // 2D-2nd :663-667
            vx = vx + p.force_x[p.it - 1] * DELTAT / rho;
            vy = vy + p.force_y[p.it - 1] * DELTAT / rho_half_x_half_y;
        }
        if (i == 1 || i == p.nx || j == 1 || j == p.ny) { vx = 0.0; vy = 0.0; }   // :669-680
        p.vx[q] = vx;
        p.vy[q] = vy;

        // energy box: 2D-2nd :695-704 (NPML+1..N-NPML), 2D-4th :696-705 (NPML..N-NPML+1)
        const int e0 = ORDER == 4 ? p.npml : p.npml + 1;
        const int ex1 = ORDER == 4 ? p.nx - p.npml + 1 : p.nx - p.npml;
        const int ey1 = ORDER == 4 ? p.ny - p.npml + 1 : p.ny - p.npml;
        if (i >= e0 && i <= ex1 && j >= e0 && j <= ey1) {
            const double l = p.lambda[q], m = p.mu[q];
            const double sxx = p.sxx[q], syy = p.syy[q], sxy = p.sxy[q];
            ekin = 0.5 * (rho * (vx * vx + vy * vy));
            // one division for both 1/(4 mu (lambda + mu)) and 1/(2 mu): the energy is a sum whose
            // order differs from the reference's anyway (tolerance 1e-11, not bitwise)
            const double inv4 = __drcp_rn(4.0 * m * (l + m));
            const double epsilon_xx = ((l + 2.0 * m) * sxx - l * syy) * inv4;
            const double epsilon_yy = ((l + 2.0 * m) * syy - l * sxx) * inv4;
            const double epsilon_xy = sxy * (inv4 * (2.0 * (l + m)));
            epot = 0.5 * (epsilon_xx * sxx + epsilon_yy * syy + 2.0 * epsilon_xy * sxy);
        }
    }
    block_sum2_2d<TX * TY>(ekin, epot, red);
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        const int b = blockIdx.y * gridDim.x + blockIdx.x;
        p.partials[b] = ekin;
        p.partials[p.nblocks + b] = epot;
    }
}

// ---------------------------------------------------------------------------------------
// Paired kernels (default): one thread updates the two x-adjacent points A = (i, j), B = (i+1, j),
// i odd, so that every row access is one aligned 16-byte load or store (the x taps of both points
// come from three of them) and the address, predicate and loop-bound instructions are shared --
// the one-point kernels above spend more instructions on those than on arithmetic (ncu: 268 / 566
// instructions per point, profiles/r01_v6_ncu_cfg2.txt).  The arithmetic of a point is written
// once (stress_point2 / velocity_point2) and is the same sequence of operations as above.
// ---------------------------------------------------------------------------------------

template <int ORDER, int TX, int TY, int MINB>
__global__ void __launch_bounds__(TX *TY, MINB)
k_stress2d_pair(const __grid_constant__ Params2D p)
{
    __shared__ double red[2 * TX * TY / 32];
    const int i = 2 * (blockIdx.x * TX + threadIdx.x) + 1;            // A; B = i + 1
    const int j = blockIdx.y * TY + threadIdx.y + 1;
    double epot = 0.0, unused = 0.0;
    if (i <= p.nx && j <= p.ny) {
        const bool validB = i + 1 <= p.nx;
        const int pitch = p.pitch;
        const long long q = (long long)(j - 1) * pitch + (i - 1);           // even: 16-byte aligned
        const bool in_y = (j <= p.ylo) || (j >= p.yhi);
        const bool in_xA = (i <= p.xlo) || (i >= p.xhi), in_xB = validB && ((i + 1 <= p.xlo) || (i + 1 >= p.xhi));
        const long long qxA = in_xA ? (long long)(j - 1) * p.sxp + shell_index2(i, p.xlo, p.xhi) : 0;
        const long long qxB = in_xB ? (long long)(j - 1) * p.sxp + shell_index2(i + 1, p.xlo, p.xhi) : 0;
        const long long qy = in_y ? (long long)shell_index2(j, p.ylo, p.yhi) * pitch + (i - 1) : 0;

        const double2 lam_c = ld2(p.lambda, q), mu_c = ld2(p.mu, q), mu_jp = ld2(p.mu, q + pitch);
        const double lam_r = p.lambda[q + 2], mu_r = p.mu[q + 2];
        const double2 vx_c = ld2(p.vx, q), vx_r = ld2(p.vx, q + 2), vx_jp = ld2(p.vx, q + pitch);
        const double2 vy_c = ld2(p.vy, q), vy_l = ld2(p.vy, q - 2), vy_jm = ld2(p.vy, q - pitch);
        double2 vx_l = make_double2(0.0, 0.0), vx_jpp = vx_l, vx_jm = vx_l, vy_jp = vx_l, vy_jmm = vx_l;
        double vy_r = 0.0;
        if (ORDER == 4) {
            vx_l = ld2(p.vx, q - 2); vx_jpp = ld2(p.vx, q + 2 * pitch); vx_jm = ld2(p.vx, q - pitch);
            vy_jp = ld2(p.vy, q + pitch); vy_jmm = ld2(p.vy, q - 2 * pitch); vy_r = p.vy[q + 2];
        }
        double2 sxx = ld2(p.sxx, q), syy = ld2(p.syy, q), sxy = ld2(p.sxy, q);

        stress_point2<ORDER>(p, i, j, in_xA, in_y, qxA, qy, lam_c.x, lam_c.y, mu_c.x, mu_c.y, mu_jp.x,
                             vx_c.y, vx_c.x, vx_r.x, vx_l.y, vx_jp.x, vx_jpp.x, vx_jm.x,
                             vy_c.x, vy_l.y, vy_c.y, vy_l.x, vy_jm.x, vy_jp.x, vy_jmm.x, sxx.x, syy.x, sxy.x);
        if (validB)
            stress_point2<ORDER>(p, i + 1, j, in_xB, in_y, qxB, qy + 1, lam_c.y, lam_r, mu_c.y, mu_r, mu_jp.y,
                                 vx_r.x, vx_c.y, vx_r.y, vx_c.x, vx_jp.y, vx_jpp.y, vx_jm.y,
                                 vy_c.y, vy_c.x, vy_r, vy_l.y, vy_jm.y, vy_jp.y, vy_jmm.y, sxx.y, syy.y, sxy.y);
        st2(p.sxx, q, sxx.x, sxx.y);
        st2(p.syy, q, syy.x, syy.y);
        st2(p.sxy, q, sxy.x, sxy.y);
        epot = epot_point2<ORDER>(p, i, j, lam_c.x, mu_c.x, sxx.x, syy.x, sxy.x);
        if (validB) epot += epot_point2<ORDER>(p, i + 1, j, lam_c.y, mu_c.y, sxx.y, syy.y, sxy.y);
    }
    block_sum2_2d<TX * TY>(epot, unused, red);
    if (threadIdx.x == 0 && threadIdx.y == 0) p.partials[p.nblocks + blockIdx.y * gridDim.x + blockIdx.x] = epot;
}

template <int ORDER, int TX, int TY, int MINB>
__global__ void __launch_bounds__(TX *TY, MINB)
k_velocity2d_pair(const __grid_constant__ Params2D p)
{
    __shared__ double red[2 * TX * TY / 32];
    const int i = 2 * (blockIdx.x * TX + threadIdx.x) + 1;
    const int j = blockIdx.y * TY + threadIdx.y + 1;
    double ekin = 0.0, epot = 0.0;

    if (i <= p.nx && j <= p.ny) {
        const bool validB = i + 1 <= p.nx;
        const int pitch = p.pitch;
        const long long q = (long long)(j - 1) * pitch + (i - 1);
        const bool in_y = (j <= p.ylo) || (j >= p.yhi);
        const bool in_xA = (i <= p.xlo) || (i >= p.xhi), in_xB = validB && ((i + 1 <= p.xlo) || (i + 1 >= p.xhi));
        const long long qxA = in_xA ? (long long)(j - 1) * p.sxp + shell_index2(i, p.xlo, p.xhi) : 0;
        const long long qxB = in_xB ? (long long)(j - 1) * p.sxp + shell_index2(i + 1, p.xlo, p.xhi) : 0;
        const long long qy = in_y ? (long long)shell_index2(j, p.ylo, p.yhi) * pitch + (i - 1) : 0;

        const double2 rho_c = ld2(p.rho, q), rho_jp = ld2(p.rho, q + pitch);
        const double rho_r = p.rho[q + 2], rho_jpr = p.rho[q + pitch + 2];
        const double rhoA = rho_c.x, rhoB = rho_c.y;
        const double rho_hA = 0.25 * (rho_c.x + rho_c.y + rho_jp.y + rho_jp.x);          // 2D-2nd :627
        const double rho_hB = 0.25 * (rho_c.y + rho_r + rho_jpr + rho_jp.y);
        const double2 sxx_c = ld2(p.sxx, q), sxx_l = ld2(p.sxx, q - 2);
        const double2 sxy_c = ld2(p.sxy, q), sxy_r = ld2(p.sxy, q + 2), sxy_jm = ld2(p.sxy, q - pitch);
        const double2 syy_c = ld2(p.syy, q), syy_jp = ld2(p.syy, q + pitch);
        double2 z2 = make_double2(0.0, 0.0), sxy_l = z2, sxy_jp = z2, sxy_jmm = z2, syy_jpp = z2, syy_jm = z2;
        double sxx_r = 0.0;
        if (ORDER == 4) {
            sxx_r = p.sxx[q + 2]; sxy_l = ld2(p.sxy, q - 2); sxy_jp = ld2(p.sxy, q + pitch); sxy_jmm = ld2(p.sxy, q - 2 * pitch);
            syy_jpp = ld2(p.syy, q + 2 * pitch); syy_jm = ld2(p.syy, q - pitch);
        }
        double2 v_x = ld2(p.vx, q), v_y = ld2(p.vy, q);

        velocity_point2<ORDER>(p, i, j, in_xA, in_y, qxA, qy, rhoA, rho_hA,
                               sxx_c.x, sxx_l.y, sxx_c.y, sxx_l.x,
                               sxy_c.x, sxy_jm.x, sxy_jp.x, sxy_jmm.x, sxy_c.y, sxy_r.x, sxy_l.y,
                               syy_c.x, syy_jp.x, syy_jpp.x, syy_jm.x, v_x.x, v_y.x, ekin);
        if (validB)
            velocity_point2<ORDER>(p, i + 1, j, in_xB, in_y, qxB, qy + 1, rhoB, rho_hB,
                                   sxx_c.y, sxx_c.x, sxx_r, sxx_l.y,
                                   sxy_c.y, sxy_jm.y, sxy_jp.y, sxy_jmm.y, sxy_r.x, sxy_r.y, sxy_c.x,
                                   syy_c.y, syy_jp.y, syy_jpp.y, syy_jm.y, v_x.y, v_y.y, ekin);
        st2(p.vx, q, v_x.x, v_x.y);
        st2(p.vy, q, v_y.x, v_y.y);
    }
    block_sum2_2d<TX * TY>(ekin, epot, red);
    if (threadIdx.x == 0 && threadIdx.y == 0) p.partials[blockIdx.y * gridDim.x + blockIdx.x] = ekin;
}

static int pair_mode()
{
    static int mode = -1;
    if (mode < 0) { const char *e = getenv("CPML_2D_PAIR"); mode = (e && *e == '0') ? 0 : 1; }
    return mode;
}

static int pair_minb()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("CPML_2D_MINB"); v = e ? atoi(e) : 2; }
    return v;
}

// `grid` is the one-point geometry (32 x 8 points per block); the paired kernels cover 64 x 8.
void launch_stress2d(const Params2D &p, dim3 grid, dim3 block, cudaStream_t s)
{
    (void)block;
    if (pair_mode()) {
        const dim3 g((p.nx + 63) / 64, grid.y);
        if (pair_minb() >= 3) {
            if (p.order == 4) k_stress2d_pair<4, 32, 8, 3><<<g, dim3(32, 8), 0, s>>>(p);
            else              k_stress2d_pair<2, 32, 8, 3><<<g, dim3(32, 8), 0, s>>>(p);
        } else {
            if (p.order == 4) k_stress2d_pair<4, 32, 8, 2><<<g, dim3(32, 8), 0, s>>>(p);
            else              k_stress2d_pair<2, 32, 8, 2><<<g, dim3(32, 8), 0, s>>>(p);
        }
        return;
    }
    if (p.order == 4) k_stress2d<4, 32, 8><<<grid, dim3(32, 8), 0, s>>>(p);
    else              k_stress2d<2, 32, 8><<<grid, dim3(32, 8), 0, s>>>(p);
}

void launch_velocity2d(const Params2D &p, dim3 grid, dim3 block, cudaStream_t s)
{
    (void)block;
    if (pair_mode()) {
        const dim3 g((p.nx + 63) / 64, grid.y);
        if (pair_minb() >= 3) {
            if (p.order == 4) k_velocity2d_pair<4, 32, 8, 3><<<g, dim3(32, 8), 0, s>>>(p);
            else              k_velocity2d_pair<2, 32, 8, 3><<<g, dim3(32, 8), 0, s>>>(p);
        } else {
            if (p.order == 4) k_velocity2d_pair<4, 32, 8, 2><<<g, dim3(32, 8), 0, s>>>(p);
            else              k_velocity2d_pair<2, 32, 8, 2><<<g, dim3(32, 8), 0, s>>>(p);
        }
        return;
    }
    if (p.order == 4) k_velocity2d<4, 32, 8><<<grid, dim3(32, 8), 0, s>>>(p);
    else              k_velocity2d<2, 32, 8><<<grid, dim3(32, 8), 0, s>>>(p);
}

}  // namespace cpml
