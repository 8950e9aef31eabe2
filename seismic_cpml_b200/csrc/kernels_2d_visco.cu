// 2-D viscoelastic C-PML kernels for sm_100a, second and fourth order in space, N_SLS = 3.
//
//   k_vstress2d<ORDER>    sigma_xx/yy + e1, e11 (2D-visco-4th :771-830), sigma_xy + e13 (:832-871)
//   k_vvelocity2d<ORDER>  vx, vy (:879-925), source (:927-972), Dirichlet (:974-985)
//   k_vpressure2d         sispressure(it, irec) (:1004-1035)
//   k_venergy2d           COMPUTE_ENERGY (:1037-1066), only when the driver asks for it
// (line numbers: seismic_CPML_2D_velocity_and_stress_fourth_order_viscoelastic.f90; the
// second-order file differs only in the difference operator.)  The step is finished by k_post3d
// (energy sums + velocity seismograms), shared with the other solvers.
//
// The reference copies the three memory-variable arrays to "_old" copies every step (:767-769)
// and keeps both; a point's update only needs its own old value, so here the old value is the
// register the new one is computed from and each memory variable is one array, read and written
// once per step (18 words per point instead of 36 + the copy's 36).
// These programs multiply by precomputed 9/(8 DELTAX) and 1/(24 DELTAX) (:210-213) instead of
// dividing, so the operator has no division; /K (C-PML) and /rho go through div_exact.  Compiled
// with -fmad=false: fields and memory variables are bit-identical to an IEEE (non-FMA) build of
// the reference.
#include "cpml_internal.h"

namespace cpml {

__device__ __forceinline__ double vapply2(double *__restrict__ mem, long long q, double b, double a, double K, double rK, double value)
{
    double m = mem[q];
    m = b * m + a * value;
    mem[q] = m;
    return div_exact(value, K, rK) + m;
}

// a / rho through the correctly rounded reciprocal (see div_rho in kernels_2d.cu)
__device__ __forceinline__ double vdiv_rho(double a, double rho, int exact_ok)
{
    return exact_ok ? div_exact(a, rho, __drcp_rn(rho)) : a / rho;
}

__device__ __forceinline__ int vshell2(int i, int lo, int hi) { return i <= lo ? i - 1 : lo + (i - hi); }

// forward difference, :724 (fourth order) / second-order file :718
template <int ORDER>
__device__ __forceinline__ double vd_fwd(const double *f, long long q, long long s, double c98, double c24)
{
    if (ORDER == 2) return (f[q + s] - f[q]) * c98;                // c98 carries ONE_OVER_DELTA
    return (f[q + s] - f[q]) * c98 + (f[q - s] - f[q + 2 * s]) * c24;
}
// backward difference, :725
template <int ORDER>
__device__ __forceinline__ double vd_bwd(const double *f, long long q, long long s, double c98, double c24)
{
    if (ORDER == 2) return (f[q] - f[q - s]) * c98;
    return (f[q] - f[q - s]) * c98 + (f[q - 2 * s] - f[q + s]) * c24;
}

template <int ORDER, int TX, int TY>
__global__ void __launch_bounds__(TX *TY)
k_vstress2d(const __grid_constant__ Params2D p)
{
    const int i = blockIdx.x * TX + threadIdx.x + 1;
    const int j = blockIdx.y * TY + threadIdx.y + 1;
    if (i > p.nx || j > p.ny) return;
    const int pitch = p.pitch;
    const long long q = (long long)(j - 1) * pitch + (i - 1);
    const bool in_x = (i <= p.xlo) || (i >= p.xhi);
    const bool in_y = (j <= p.ylo) || (j >= p.yhi);
    const long long qx = in_x ? (long long)(j - 1) * p.sxp + vshell2(i, p.xlo, p.xhi) : 0;
    const long long qy = in_y ? (long long)vshell2(j, p.ylo, p.yhi) * pitch + (i - 1) : 0;
    const double DELTAT = p.deltat;

    if (i <= p.nx - 1 && j >= 2) {                                  // :771-772
        const double lambda_half_x = 0.5 * (p.lambda[q + 1] + p.lambda[q]);
        const double mu_half_x = 0.5 * (p.mu[q + 1] + p.mu[q]);
        const double lambda_plus_mu_half_x = lambda_half_x + mu_half_x;
        const double lambda_plus_two_mu_half_x = lambda_half_x + 2.0 * mu_half_x;
        double value_dvx_dx = vd_fwd<ORDER>(p.vx, q, 1, p.c98x, p.c24x);
        double value_dvy_dy = vd_bwd<ORDER>(p.vy, q, pitch, p.c98y, p.c24y);
        if (in_x) value_dvx_dx = vapply2(p.mx[0], qx, p.cx.b_half[i], p.cx.a_half[i], p.cx.K_half[i], p.cx.rK_half[i], value_dvx_dx);
        if (in_y) value_dvy_dy = vapply2(p.my[0], qy, p.cy.b[j], p.cy.a[j], p.cy.K[j], p.cy.rK[j], value_dvy_dy);

        double sum_of_memory_variables_e1 = 0.0, sum_of_memory_variables_e11 = 0.0;
#pragma unroll
        for (int l = 0; l < 3; l++) {                               // :795-812
            const double e1_old = __ldcs(p.e1[l] + q), e11_old = __ldcs(p.e11[l] + q);
            const double e1_new = (e1_old + (value_dvx_dx + value_dvy_dy) * p.dt_phi1[l] - e1_old * p.half1[l]) * p.mul1[l];
            const double e11_new = (e11_old + 0.5 * (value_dvx_dx - value_dvy_dy) * p.dt_phi2[l] - e11_old * p.half2[l]) * p.mul2[l];
            __stcs(p.e1[l] + q, e1_new);
            __stcs(p.e11[l] + q, e11_new);
            sum_of_memory_variables_e1 = sum_of_memory_variables_e1 + e1_new + e1_old;
            sum_of_memory_variables_e11 = sum_of_memory_variables_e11 + e11_new + e11_old;
        }
        p.sxx[q] = p.sxx[q] + (lambda_plus_two_mu_half_x * value_dvx_dx + lambda_half_x * value_dvy_dy
                               + (0.5 * lambda_plus_mu_half_x * sum_of_memory_variables_e1 + mu_half_x * sum_of_memory_variables_e11)) * DELTAT;
        p.syy[q] = p.syy[q] + (lambda_half_x * value_dvx_dx + lambda_plus_two_mu_half_x * value_dvy_dy
                               + (0.5 * lambda_plus_mu_half_x * sum_of_memory_variables_e1 - mu_half_x * sum_of_memory_variables_e11)) * DELTAT;
    }
    if (i >= 2 && j <= p.ny - 1) {                                  // :832-833
        const double mu_half_y = 0.5 * (p.mu[q + pitch] + p.mu[q]);
        double value_dvy_dx = vd_bwd<ORDER>(p.vy, q, 1, p.c98x, p.c24x);
        double value_dvx_dy = vd_fwd<ORDER>(p.vx, q, pitch, p.c98y, p.c24y);
        if (in_x) value_dvy_dx = vapply2(p.mx[1], qx, p.cx.b[i], p.cx.a[i], p.cx.K[i], p.cx.rK[i], value_dvy_dx);
        if (in_y) value_dvx_dy = vapply2(p.my[1], qy, p.cy.b_half[j], p.cy.a_half[j], p.cy.K_half[j], p.cy.rK_half[j], value_dvx_dy);
        double sum_of_memory_variables_e13 = 0.0;
#pragma unroll
        for (int l = 0; l < 3; l++) {                               // :850-859
            const double e13_old = __ldcs(p.e13[l] + q);
            const double e13_new = (e13_old + (value_dvy_dx + value_dvx_dy) * p.dt_phi2[l] - e13_old * p.half2[l]) * p.mul2[l];
            __stcs(p.e13[l] + q, e13_new);
            sum_of_memory_variables_e13 = sum_of_memory_variables_e13 + e13_new + e13_old;
        }
        p.sxy[q] = p.sxy[q] + mu_half_y * (value_dvy_dx + value_dvx_dy + 0.5 * sum_of_memory_variables_e13) * DELTAT;
    }
}

template <int ORDER, int TX, int TY>
__global__ void __launch_bounds__(TX *TY)
k_vvelocity2d(const __grid_constant__ Params2D p)
{
    const int i = blockIdx.x * TX + threadIdx.x + 1;
    const int j = blockIdx.y * TY + threadIdx.y + 1;
    if (i > p.nx || j > p.ny) return;
    const int pitch = p.pitch;
    const long long q = (long long)(j - 1) * pitch + (i - 1);
    const bool in_x = (i <= p.xlo) || (i >= p.xhi);
    const bool in_y = (j <= p.ylo) || (j >= p.yhi);
    const long long qx = in_x ? (long long)(j - 1) * p.sxp + vshell2(i, p.xlo, p.xhi) : 0;
    const long long qy = in_y ? (long long)vshell2(j, p.ylo, p.yhi) * pitch + (i - 1) : 0;
    const double DELTAT = p.deltat;
    const double rho = p.rho[q];
    const double rho_half_x_half_y = 0.25 * (rho + p.rho[q + 1] + p.rho[q + 1 + pitch] + p.rho[q + pitch]);
    double vx = p.vx[q], vy = p.vy[q];

    if (i >= 2 && j >= 2) {                                         // :879-896
        double value_dsigma_xx_dx = vd_bwd<ORDER>(p.sxx, q, 1, p.c98x, p.c24x);
        double value_dsigma_xy_dy = vd_bwd<ORDER>(p.sxy, q, pitch, p.c98y, p.c24y);
        if (in_x) value_dsigma_xx_dx = vapply2(p.mx[2], qx, p.cx.b[i], p.cx.a[i], p.cx.K[i], p.cx.rK[i], value_dsigma_xx_dx);
        if (in_y) value_dsigma_xy_dy = vapply2(p.my[2], qy, p.cy.b[j], p.cy.a[j], p.cy.K[j], p.cy.rK[j], value_dsigma_xy_dy);
        vx = vx + vdiv_rho((value_dsigma_xx_dx + value_dsigma_xy_dy) * DELTAT, rho, p.rho_exact);
    }
    if (i <= p.nx - 1 && j <= p.ny - 1) {                           // :898-925
        double value_dsigma_xy_dx = vd_fwd<ORDER>(p.sxy, q, 1, p.c98x, p.c24x);
        double value_dsigma_yy_dy = vd_fwd<ORDER>(p.syy, q, pitch, p.c98y, p.c24y);
        if (in_x) value_dsigma_xy_dx = vapply2(p.mx[3], qx, p.cx.b_half[i], p.cx.a_half[i], p.cx.K_half[i], p.cx.rK_half[i], value_dsigma_xy_dx);
        if (in_y) value_dsigma_yy_dy = vapply2(p.my[3], qy, p.cy.b_half[j], p.cy.a_half[j], p.cy.K_half[j], p.cy.rK_half[j], value_dsigma_yy_dy);
        vy = vy + vdiv_rho((value_dsigma_xy_dx + value_dsigma_yy_dy) * DELTAT, rho_half_x_half_y, p.rho_exact);
    }
    if (i == p.isrc && j == p.jsrc) {                               // :971-972
        vx = vx + p.force_x[p.it - 1] * DELTAT / rho;
        vy = vy + p.force_y[p.it - 1] * DELTAT / rho_half_x_half_y;
    }
    if (i == 1 || i == p.nx || j == 1 || j == p.ny) { vx = 0.0; vy = 0.0; }      // :974-985
    p.vx[q] = vx;
    p.vy[q] = vy;
}

// sispressure(it, irec), :1004-1035
__global__ void __launch_bounds__(64) k_vpressure2d(const __grid_constant__ Params2D p, const int *ix_rec, const int *iy_rec,
                                                     int nrec, int nstep, double *sispressure)
{
    for (int r = threadIdx.x; r < nrec; r += 64) {
        const long long q = (long long)(iy_rec[r] - 1) * p.pitch + (ix_rec[r] - 1);
        const double lambda_half_x = 0.5 * (p.lambda[q + 1] + p.lambda[q]);
        const double mu_half_x = 0.5 * (p.mu[q + 1] + p.mu[q]);
        const double epsilon_xx = ((lambda_half_x + 2.0 * mu_half_x) * p.sxx[q] - lambda_half_x * p.syy[q]) /
                                  (4.0 * mu_half_x * (lambda_half_x + mu_half_x));
        const double epsilon_yy = ((lambda_half_x + 2.0 * mu_half_x) * p.syy[q] - lambda_half_x * p.sxx[q]) /
                                  (4.0 * mu_half_x * (lambda_half_x + mu_half_x));
        const double TWO_THIRDS = 2.0 / 3.0;
        sispressure[(long long)r * nstep + (p.it - 1)] = -(lambda_half_x + TWO_THIRDS * mu_half_x) * (epsilon_xx + epsilon_yy);
    }
}

// COMPUTE_ENERGY, :1037-1066: kinetic energy with vy interpolated back to the vx node, potential
// energy with the material parameters interpolated to the stress nodes; per-block partials.
template <int TX, int TY>
__global__ void __launch_bounds__(TX *TY) k_venergy2d(const __grid_constant__ Params2D p)
{
    __shared__ double red[2 * TX * TY / 32];
    const int i = blockIdx.x * TX + threadIdx.x + 1;
    const int j = blockIdx.y * TY + threadIdx.y + 1;
    double ekin = 0.0, epot = 0.0;
    if (i >= p.npml + 1 && i <= p.nx - p.npml && j >= p.npml + 1 && j <= p.ny - p.npml) {
        const int pitch = p.pitch;
        const long long q = (long long)(j - 1) * pitch + (i - 1);
        const double vy_interpolated = 0.25 * (p.vy[q] + p.vy[q - 1] + p.vy[q - 1 - pitch] + p.vy[q - pitch]);
        const double vx = p.vx[q];
        ekin = 0.5 * p.rho[q] * (vx * vx + vy_interpolated * vy_interpolated);
        const double lambda_half_x = 0.5 * (p.lambda[q + 1] + p.lambda[q]);
        const double mu_half_x = 0.5 * (p.mu[q + 1] + p.mu[q]);
        const double mu_half_y = 0.5 * (p.mu[q + pitch] + p.mu[q]);
        const double sxx = p.sxx[q], syy = p.syy[q], sxy = p.sxy[q];
        const double den = 4.0 * mu_half_x * (lambda_half_x + mu_half_x);
        const double epsilon_xx = ((lambda_half_x + 2.0 * mu_half_x) * sxx - lambda_half_x * syy) / den;
        const double epsilon_yy = ((lambda_half_x + 2.0 * mu_half_x) * syy - lambda_half_x * sxx) / den;
        const double epsilon_xy = sxy / (2.0 * mu_half_y);
        epot = 0.5 * (epsilon_xx * sxx + epsilon_yy * syy + 2.0 * epsilon_xy * sxy);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ekin += __shfl_down_sync(0xffffffffu, ekin, o);
        epot += __shfl_down_sync(0xffffffffu, epot, o);
    }
    const int t = threadIdx.y * TX + threadIdx.x, w = t >> 5, l = t & 31;
    constexpr int NW = TX * TY / 32;
    if (l == 0) { red[w] = ekin; red[NW + w] = epot; }
    __syncthreads();
    if (t == 0) {
        double a = 0.0, b = 0.0;
        for (int k = 0; k < NW; k++) { a += red[k]; b += red[NW + k]; }
        const int blk = blockIdx.y * gridDim.x + blockIdx.x;
        p.partials[blk] = a;
        p.partials[p.nblocks + blk] = b;
    }
}

void launch_vstress2d(const Params2D &p, dim3 grid, cudaStream_t s)
{
    if (p.order == 4) k_vstress2d<4, 32, 8><<<grid, dim3(32, 8), 0, s>>>(p);
    else              k_vstress2d<2, 32, 8><<<grid, dim3(32, 8), 0, s>>>(p);
}

void launch_vvelocity2d(const Params2D &p, dim3 grid, cudaStream_t s)
{
    if (p.order == 4) k_vvelocity2d<4, 32, 8><<<grid, dim3(32, 8), 0, s>>>(p);
    else              k_vvelocity2d<2, 32, 8><<<grid, dim3(32, 8), 0, s>>>(p);
}

void launch_vpressure2d(const Params2D &p, const int *ix_rec, const int *iy_rec, int nrec, int nstep, double *sispressure, cudaStream_t s)
{
    if (nrec > 0) k_vpressure2d<<<1, 64, 0, s>>>(p, ix_rec, iy_rec, nrec, nstep, sispressure);
}

void launch_venergy2d(const Params2D &p, dim3 grid, cudaStream_t s) { k_venergy2d<32, 8><<<grid, dim3(32, 8), 0, s>>>(p); }

}  // namespace cpml
