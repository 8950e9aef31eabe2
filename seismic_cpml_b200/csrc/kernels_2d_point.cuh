// Per-point arithmetic of the 2-D isotropic solvers, shared by the pair kernels (kernels_2d.cu) and the TMA-staged
// y-marching kernels (kernels_2d_ws.cu): the C-PML recursion, the staggered material averages, the second / fourth
// order differences and the updates of seismic_CPML_2D_isotropic_{second,fourth}_order.f90 (:556-713 / :557-714), written
// once so that both kernel families produce the same bits.
#pragma once
#include "cpml_internal.h"

namespace cpml {

__device__ __forceinline__ double cpml_apply2(double *__restrict__ mem, long long q,
                                              double b, double a, double K, double rK, double value)
{
    double m = mem[q];
    m = b * m + a * value;
    mem[q] = m;
    return div_exact(value, K, rK) + m;     // value / K + m, rK = RN(1/K)
}

__device__ __forceinline__ int shell_index2(int i, int lo, int hi)
{
    return i <= lo ? i - 1 : lo + (i - hi);
}

// a / rho for a density read from the material arrays: the correctly rounded reciprocal comes from
// __drcp_rn (IEEE round-to-nearest), then div_exact -- the correctly rounded quotient in about half
// the instructions of the generic division.  `exact_ok` is 0 when cpml_set_material_2d found a
// density (or an interpolated density) with an all-ones significand, the one case Markstein's
// theorem excludes; the generic division runs then.
__device__ __forceinline__ double div_rho(double a, double rho, int exact_ok)
{
    return exact_ok ? div_exact(a, rho, __drcp_rn(rho)) : a / rho;
}

__device__ __forceinline__ double2 ld2(const double *f, long long q) { return *reinterpret_cast<const double2 *>(f + q); }
__device__ __forceinline__ void st2(double *f, long long q, double a, double b) { *reinterpret_cast<double2 *>(f + q) = make_double2(a, b); }

// difference of a four-tap (ORDER 4) or two-tap (ORDER 2) stencil: 27 a - 27 b - c + d over den
// (the numerator; the two differences of a nest are divided behind one shared range test, div_exact2)
template <int ORDER>
__device__ __forceinline__ double diffn(double a, double b, double c, double d)
{
    if (ORDER == 2) return a - b;
    return 27.0 * a - 27.0 * b - c + d;
}

template <int ORDER>
__device__ __forceinline__ void stress_point2(const Params2D &p, int i, int j, bool in_x, bool in_y, long long qx, long long qy,
                                              double lam_c, double lam_ip, double mu_c, double mu_ip, double mu_jp,
                                              // vx: i+1, i, i+2, i-1 on row j; j+1, j+2, j-1 at column i
                                              double vx_ip, double vx_c, double vx_ipp, double vx_im, double vx_jp, double vx_jpp, double vx_jm,
                                              // vy: i-1, i+1, i-2 on row j; j-1, j+1, j-2 at column i
                                              double vy_c, double vy_im, double vy_ip, double vy_imm, double vy_jm, double vy_jp, double vy_jmm,
                                              double &sxx, double &syy, double &sxy)
{
    const double DELTAT = p.deltat;
    if (i <= p.nx - 1 && j >= 2) {
        const double lambda_half_x = 0.5 * (lam_ip + lam_c);
        const double mu_half_x = 0.5 * (mu_ip + mu_c);
        const double lambda_plus_two_mu_half_x = lambda_half_x + 2.0 * mu_half_x;
        double value_dvx_dx = diffn<ORDER>(vx_ip, vx_c, vx_ipp, vx_im);
        double value_dvy_dy = diffn<ORDER>(vy_c, vy_jm, vy_jp, vy_jmm);
        div_exact2(value_dvx_dx, value_dvy_dy, p.denx, p.rdenx, p.deny, p.rdeny);
        if (in_x) value_dvx_dx = cpml_apply2(p.mx[0], qx, p.cx.b_half[i], p.cx.a_half[i], p.cx.K_half[i], p.cx.rK_half[i], value_dvx_dx);
        if (in_y) value_dvy_dy = cpml_apply2(p.my[0], qy, p.cy.b[j], p.cy.a[j], p.cy.K[j], p.cy.rK[j], value_dvy_dy);
        sxx = sxx + (lambda_plus_two_mu_half_x * value_dvx_dx + lambda_half_x * value_dvy_dy) * DELTAT;
        syy = syy + (lambda_half_x * value_dvx_dx + lambda_plus_two_mu_half_x * value_dvy_dy) * DELTAT;
    }
    if (i >= 2 && j <= p.ny - 1) {
        const double mu_half_y = 0.5 * (mu_jp + mu_c);
        double value_dvy_dx = diffn<ORDER>(vy_c, vy_im, vy_ip, vy_imm);
        double value_dvx_dy = diffn<ORDER>(vx_jp, vx_c, vx_jpp, vx_jm);
        div_exact2(value_dvy_dx, value_dvx_dy, p.denx, p.rdenx, p.deny, p.rdeny);
        if (in_x) value_dvy_dx = cpml_apply2(p.mx[1], qx, p.cx.b[i], p.cx.a[i], p.cx.K[i], p.cx.rK[i], value_dvy_dx);
        // quirk B3, see k_stress2d
        if (in_y) value_dvx_dy = cpml_apply2(p.my[1], qy, p.cy.b_half[j], p.cy.a_half[j],
                                             ORDER == 4 ? p.cy.K[j] : p.cy.K_half[j],
                                             ORDER == 4 ? p.cy.rK[j] : p.cy.rK_half[j], value_dvx_dy);
        sxy = sxy + mu_half_y * (value_dvy_dx + value_dvx_dy) * DELTAT;
    }
}

// Potential energy of one point from the stresses of this step (2D-2nd :706-711); lives in the STRESS kernel, which
// has lambda, mu and the new stresses in registers, so that the velocity kernel does not stream lambda and mu a
// second time (18 instead of 20 words per point-update).
template <int ORDER>
__device__ __forceinline__ double epot_point2(const Params2D &p, int i, int j, double lam, double mu, double sxx, double syy, double sxy)
{
    const int e0 = ORDER == 4 ? p.npml : p.npml + 1;                       // 2D-2nd :695-704, 2D-4th :696-705
    const int ex1 = ORDER == 4 ? p.nx - p.npml + 1 : p.nx - p.npml;
    const int ey1 = ORDER == 4 ? p.ny - p.npml + 1 : p.ny - p.npml;
    if (!(i >= e0 && i <= ex1 && j >= e0 && j <= ey1)) return 0.0;
    // one division for both 1/(4 mu (lambda + mu)) and 1/(2 mu): the energy is a sum whose order differs from the
    // reference's anyway (tolerance 1e-11, not bitwise)
    const double inv4 = __drcp_rn(4.0 * mu * (lam + mu));
    const double epsilon_xx = ((lam + 2.0 * mu) * sxx - lam * syy) * inv4;
    const double epsilon_yy = ((lam + 2.0 * mu) * syy - lam * sxx) * inv4;
    const double epsilon_xy = sxy * (inv4 * (2.0 * (lam + mu)));
    return 0.5 * (epsilon_xx * sxx + epsilon_yy * syy + 2.0 * epsilon_xy * sxy);
}

template <int ORDER>
__device__ __forceinline__ void velocity_point2(const Params2D &p, int i, int j, bool in_x, bool in_y, long long qx, long long qy,
                                                double rho, double rho_half_x_half_y,
                                                // sxx: i, i-1, i+1, i-2 ; sxy: (j, j-1, j+1, j-2) and (i+1, i, i+2, i-1) ; syy: j+1, j, j+2, j-1
                                                double sxx_c, double sxx_im, double sxx_ip, double sxx_imm,
                                                double sxy_c, double sxy_jm, double sxy_jp, double sxy_jmm,
                                                double sxy_ip, double sxy_ipp, double sxy_im,
                                                double syy_c, double syy_jp, double syy_jpp, double syy_jm,
                                                double &vx, double &vy, double &ekin)
{
    const double DELTAT = p.deltat;
    if (i >= 2 && j >= 2) {
        double value_dsigmaxx_dx = diffn<ORDER>(sxx_c, sxx_im, sxx_ip, sxx_imm);
        double value_dsigmaxy_dy = diffn<ORDER>(sxy_c, sxy_jm, sxy_jp, sxy_jmm);
        div_exact2(value_dsigmaxx_dx, value_dsigmaxy_dy, p.denx, p.rdenx, p.deny, p.rdeny);
        if (in_x) value_dsigmaxx_dx = cpml_apply2(p.mx[2], qx, p.cx.b[i], p.cx.a[i], p.cx.K[i], p.cx.rK[i], value_dsigmaxx_dx);
        if (in_y) value_dsigmaxy_dy = cpml_apply2(p.my[2], qy, p.cy.b[j], p.cy.a[j], p.cy.K[j], p.cy.rK[j], value_dsigmaxy_dy);
        vx = vx + div_rho((value_dsigmaxx_dx + value_dsigmaxy_dy) * DELTAT, rho, p.rho_exact);
    }
    if (i <= p.nx - 1 && j <= p.ny - 1) {
        double value_dsigmaxy_dx = diffn<ORDER>(sxy_ip, sxy_c, sxy_ipp, sxy_im);
        double value_dsigmayy_dy = diffn<ORDER>(syy_jp, syy_c, syy_jpp, syy_jm);
        div_exact2(value_dsigmaxy_dx, value_dsigmayy_dy, p.denx, p.rdenx, p.deny, p.rdeny);
        if (in_x) value_dsigmaxy_dx = cpml_apply2(p.mx[3], qx, p.cx.b_half[i], p.cx.a_half[i], p.cx.K_half[i], p.cx.rK_half[i], value_dsigmaxy_dx);
        if (in_y) value_dsigmayy_dy = cpml_apply2(p.my[3], qy, p.cy.b_half[j], p.cy.a_half[j], p.cy.K_half[j], p.cy.rK_half[j], value_dsigmayy_dy);
        vy = vy + div_rho((value_dsigmaxy_dx + value_dsigmayy_dy) * DELTAT, rho_half_x_half_y, p.rho_exact);
    }
    if (i == p.isrc && j == p.jsrc) {               // 2D-2nd :663-667
        vx = vx + p.force_x[p.it - 1] * DELTAT / rho;
        vy = vy + p.force_y[p.it - 1] * DELTAT / rho_half_x_half_y;
    }
    if (i == 1 || i == p.nx || j == 1 || j == p.ny) { vx = 0.0; vy = 0.0; }   // :669-680

    const int e0 = ORDER == 4 ? p.npml : p.npml + 1;
    const int ex1 = ORDER == 4 ? p.nx - p.npml + 1 : p.nx - p.npml;
    const int ey1 = ORDER == 4 ? p.ny - p.npml + 1 : p.ny - p.npml;
    if (i >= e0 && i <= ex1 && j >= e0 && j <= ey1) ekin += 0.5 * (rho * (vx * vx + vy * vy));     // potential part: k_stress2d_pair
}

}  // namespace cpml
