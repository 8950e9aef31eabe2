/*
 * cpml_oracle_visco2d.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 * See cpml_oracle.h for the contract and how parity is pinned (an execution of the reference source, oracle/f90_exec.py).
 *
 * Restates, loop nest by loop nest and operation by operation, the hot paths of
 *   /root/reference/seismic_CPML_2D_velocity_and_stress_fourth_order_viscoelastic.f90  (2D-visco-4th)
 *   /root/reference/seismic_CPML_2D_velocity_and_stress_second_order_viscoelastic.f90  (2D-visco-2nd)
 * (line numbers below: the fourth-order file).  N_SLS = 3 Zener solids, memory variables in the
 * auxiliary-differential-equation form with the time averaging of Robertsson et al. (1994).
 */
#include "cpml_oracle.h"
#include "oracle_internal.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define PI 3.141592653589793238462643 /* :216 */

/* :931-958 */
void oracle_source_series_ricker(int nstep, double deltat, double f0, double t0, double factor,
                                 double angle_force_deg, double deltax, double deltay,
                                 double *force_x, double *force_y)
{
    const double degrees_to_radians = PI / 180.0; /* :219 */
    for (int it = 1; it <= nstep; it++) {
        double a = PI * PI * f0 * f0;
        double t = (double)(it - 1) * deltat;
        double force_source_term = factor * (1.0 - 2.0 * a * ((t - t0) * (t - t0))) * exp(-a * ((t - t0) * (t - t0)));
        force_source_term = force_source_term / (deltax * deltay);
        force_x[it - 1] = sin(angle_force_deg * degrees_to_radians) * force_source_term;
        force_y[it - 1] = cos(angle_force_deg * degrees_to_radians) * force_source_term;
    }
}

int oracle_run_2d_visco(const oraclev2d_config *cfg,
                        const double *lambda_in, const double *mu_in, const double *rho_in,
                        const double *a_x, const double *b_x, const double *K_x,
                        const double *a_x_half, const double *b_x_half, const double *K_x_half,
                        const double *a_y, const double *b_y, const double *K_y,
                        const double *a_y_half, const double *b_y_half, const double *K_y_half,
                        const double *force_x, const double *force_y,
                        const int *ix_rec, const int *iy_rec,
                        double *sisvx, double *sisvy, double *sispressure,
                        double *energy_kinetic, double *energy_potential,
                        double *fields_final, double *memvar_final, double *velocnorm_final)
{
    const int NX = cfg->nx, NY = cfg->ny, NSTEP = cfg->nstep, NREC = cfg->nrec;
    const int NPOINTS_PML = cfg->npoints_pml;
    const double DELTAX = cfg->deltax, DELTAY = cfg->deltay, DELTAT = cfg->deltat;
    const int fourth = (cfg->order == 4);
    enum { N_SLS = 3 };
    if (cfg->order != 2 && cfg->order != 4) return 1;

    /* :208-213 (second-order file :208-210) */
    const double ONE_OVER_DELTAX = 1.0 / DELTAX, ONE_OVER_DELTAY = 1.0 / DELTAY;
    const double NINE_OVER_8_DELTAX = 9.0 / (8.0 * DELTAX), NINE_OVER_8_DELTAY = 9.0 / (8.0 * DELTAY);
    const double ONE_OVER_24_DELTAX = 1.0 / (24.0 * DELTAX), ONE_OVER_24_DELTAY = 1.0 / (24.0 * DELTAY);
    const double TWO_THIRDS = 2.0 / 3.0;                      /* :224 */

    /* attenuation constants, :386-399 */
    double one_over_tau_sigma_nu1[N_SLS], one_over_tau_sigma_nu2[N_SLS];
    double HALF_DELTAT_over_tau_sigma_nu1[N_SLS], HALF_DELTAT_over_tau_sigma_nu2[N_SLS];
    double multiplication_factor_tau_sigma_nu1[N_SLS], multiplication_factor_tau_sigma_nu2[N_SLS];
    double DELTAT_phi_nu1[N_SLS], DELTAT_phi_nu2[N_SLS];
    double te1[N_SLS], ts1[N_SLS], te2[N_SLS], ts2[N_SLS];
    for (int l = 0; l < N_SLS; l++) {
        if (cfg->viscoelastic_attenuation) {
            te1[l] = cfg->tau_epsilon_nu1[l]; ts1[l] = cfg->tau_sigma_nu1[l];
            te2[l] = cfg->tau_epsilon_nu2[l]; ts2[l] = cfg->tau_sigma_nu2[l];
        } else {                                              /* :374-380 */
            te1[l] = ts1[l] = te2[l] = ts2[l] = 1.0;
        }
    }
    double sum1 = 0.0, sum2 = 0.0;                            /* sum(tau_epsilon/tau_sigma), :398-399 */
    for (int l = 0; l < N_SLS; l++) { sum1 = sum1 + te1[l] / ts1[l]; sum2 = sum2 + te2[l] / ts2[l]; }
    for (int l = 0; l < N_SLS; l++) {
        one_over_tau_sigma_nu1[l] = 1.0 / ts1[l];
        one_over_tau_sigma_nu2[l] = 1.0 / ts2[l];
        HALF_DELTAT_over_tau_sigma_nu1[l] = 0.5 * DELTAT / ts1[l];
        HALF_DELTAT_over_tau_sigma_nu2[l] = 0.5 * DELTAT / ts2[l];
        multiplication_factor_tau_sigma_nu1[l] = 1.0 / (1.0 + 0.5 * DELTAT * one_over_tau_sigma_nu1[l]);
        multiplication_factor_tau_sigma_nu2[l] = 1.0 / (1.0 + 0.5 * DELTAT * one_over_tau_sigma_nu2[l]);
        DELTAT_phi_nu1[l] = DELTAT * (1.0 - te1[l] / ts1[l]) / ts1[l] / sum1;
        DELTAT_phi_nu2[l] = DELTAT * (1.0 - te2[l] / ts2[l]) / ts2[l] / sum2;
    }

    /* main arrays carry the (0:NX+1,0:NY+1) ring of the fourth-order file (:235); the second-order
     * loops never touch it.  Memory arrays are (NX,NY) / (NX,NY,N_SLS) (:246-262, :340-342). */
    const size_t LD = (size_t)NX + 2;
    const size_t N = LD * ((size_t)NY + 2);
    const size_t NM = (size_t)NX * NY;
#define A2(arr, i, j) arr[(size_t)(i) + LD * (size_t)(j)]
#define M2(arr, i, j) arr[(size_t)((i) - 1) + (size_t)NX * (size_t)((j) - 1)]
#define M3(arr, i, j, l) arr[(size_t)((i) - 1) + (size_t)NX * ((size_t)((j) - 1) + (size_t)NY * (size_t)(l))]
#define P1(arr, i) arr[(i) - 1]
    double *vx = calloc(N, sizeof(double)), *vy = calloc(N, sizeof(double));
    double *sigma_xx = calloc(N, sizeof(double)), *sigma_yy = calloc(N, sizeof(double)), *sigma_xy = calloc(N, sizeof(double));
    double *lambda_unrelaxed = calloc(N, sizeof(double)), *mu_unrelaxed = calloc(N, sizeof(double)), *rho = calloc(N, sizeof(double));
    double *memory_dvx_dx = calloc(NM, sizeof(double)), *memory_dvx_dy = calloc(NM, sizeof(double));
    double *memory_dvy_dx = calloc(NM, sizeof(double)), *memory_dvy_dy = calloc(NM, sizeof(double));
    double *memory_dsigma_xx_dx = calloc(NM, sizeof(double)), *memory_dsigma_yy_dy = calloc(NM, sizeof(double));
    double *memory_dsigma_xy_dx = calloc(NM, sizeof(double)), *memory_dsigma_xy_dy = calloc(NM, sizeof(double));
    double *e1 = calloc(NM * N_SLS, sizeof(double)), *e1_old = calloc(NM * N_SLS, sizeof(double));
    double *e11 = calloc(NM * N_SLS, sizeof(double)), *e11_old = calloc(NM * N_SLS, sizeof(double));
    double *e13 = calloc(NM * N_SLS, sizeof(double)), *e13_old = calloc(NM * N_SLS, sizeof(double));

    for (int j = 1; j <= NY; j++)
        for (int i = 1; i <= NX; i++) {            /* :596-602 */
            size_t s = (size_t)(i - 1) + (size_t)NX * (size_t)(j - 1);
            A2(rho, i, j) = rho_in[s];
            A2(mu_unrelaxed, i, j) = mu_in[s];
            A2(lambda_unrelaxed, i, j) = lambda_in[s];
        }
    memset(sisvx, 0, sizeof(double) * (size_t)NSTEP * NREC);   /* :684-692 */
    memset(sisvy, 0, sizeof(double) * (size_t)NSTEP * NREC);
    memset(sispressure, 0, sizeof(double) * (size_t)NSTEP * NREC);
    memset(energy_kinetic, 0, sizeof(double) * (size_t)NSTEP);
    memset(energy_potential, 0, sizeof(double) * (size_t)NSTEP);

    /* difference operators of the two files: :724-725 vs second-order :718-719 */
#define DX_FWD(f, i, j) (fourth ? (A2(f, (i) + 1, j) - A2(f, i, j)) * NINE_OVER_8_DELTAX + (A2(f, (i) - 1, j) - A2(f, (i) + 2, j)) * ONE_OVER_24_DELTAX \
                                : (A2(f, (i) + 1, j) - A2(f, i, j)) * ONE_OVER_DELTAX)
#define DX_BWD(f, i, j) (fourth ? (A2(f, i, j) - A2(f, (i) - 1, j)) * NINE_OVER_8_DELTAX + (A2(f, (i) - 2, j) - A2(f, (i) + 1, j)) * ONE_OVER_24_DELTAX \
                                : (A2(f, i, j) - A2(f, (i) - 1, j)) * ONE_OVER_DELTAX)
#define DY_FWD(f, i, j) (fourth ? (A2(f, i, (j) + 1) - A2(f, i, j)) * NINE_OVER_8_DELTAY + (A2(f, i, (j) - 1) - A2(f, i, (j) + 2)) * ONE_OVER_24_DELTAY \
                                : (A2(f, i, (j) + 1) - A2(f, i, j)) * ONE_OVER_DELTAY)
#define DY_BWD(f, i, j) (fourth ? (A2(f, i, j) - A2(f, i, (j) - 1)) * NINE_OVER_8_DELTAY + (A2(f, i, (j) - 2) - A2(f, i, (j) + 1)) * ONE_OVER_24_DELTAY \
                                : (A2(f, i, j) - A2(f, i, (j) - 1)) * ONE_OVER_DELTAY)

    double t_loop_start = oracle_now_seconds();
    for (int it = 1; it <= NSTEP; it++) {          /* :705 */
        if (it == oracle_g_warmup_steps + 1) t_loop_start = oracle_now_seconds();

        if (!cfg->viscoelastic_attenuation) {      /* :713-760 */
            for (int j = 2; j <= NY; j++)
                for (int i = 1; i <= NX - 1; i++) {
                    double lambda_half_x = 0.5 * (A2(lambda_unrelaxed, i + 1, j) + A2(lambda_unrelaxed, i, j));
                    double mu_half_x = 0.5 * (A2(mu_unrelaxed, i + 1, j) + A2(mu_unrelaxed, i, j));
                    double lambda_plus_two_mu_half_x = lambda_half_x + 2.0 * mu_half_x;
                    double value_dvx_dx = DX_FWD(vx, i, j);
                    double value_dvy_dy = DY_BWD(vy, i, j);
                    M2(memory_dvx_dx, i, j) = P1(b_x_half, i) * M2(memory_dvx_dx, i, j) + P1(a_x_half, i) * value_dvx_dx;
                    M2(memory_dvy_dy, i, j) = P1(b_y, j) * M2(memory_dvy_dy, i, j) + P1(a_y, j) * value_dvy_dy;
                    value_dvx_dx = value_dvx_dx / P1(K_x_half, i) + M2(memory_dvx_dx, i, j);
                    value_dvy_dy = value_dvy_dy / P1(K_y, j) + M2(memory_dvy_dy, i, j);
                    A2(sigma_xx, i, j) = A2(sigma_xx, i, j) + (lambda_plus_two_mu_half_x * value_dvx_dx + lambda_half_x * value_dvy_dy) * DELTAT;
                    A2(sigma_yy, i, j) = A2(sigma_yy, i, j) + (lambda_half_x * value_dvx_dx + lambda_plus_two_mu_half_x * value_dvy_dy) * DELTAT;
                }
            for (int j = 1; j <= NY - 1; j++)
                for (int i = 2; i <= NX; i++) {
                    double mu_half_y = 0.5 * (A2(mu_unrelaxed, i, j + 1) + A2(mu_unrelaxed, i, j));
                    double value_dvy_dx = DX_BWD(vy, i, j);
                    double value_dvx_dy = DY_FWD(vx, i, j);
                    M2(memory_dvy_dx, i, j) = P1(b_x, i) * M2(memory_dvy_dx, i, j) + P1(a_x, i) * value_dvy_dx;
                    M2(memory_dvx_dy, i, j) = P1(b_y_half, j) * M2(memory_dvx_dy, i, j) + P1(a_y_half, j) * value_dvx_dy;
                    value_dvy_dx = value_dvy_dx / P1(K_x, i) + M2(memory_dvy_dx, i, j);
                    value_dvx_dy = value_dvx_dy / P1(K_y_half, j) + M2(memory_dvx_dy, i, j);
                    A2(sigma_xy, i, j) = A2(sigma_xy, i, j) + mu_half_y * (value_dvy_dx + value_dvx_dy) * DELTAT;
                }
        } else {
            /* the present becomes the past for the memory variables, :767-769 */
            memcpy(e1_old, e1, NM * N_SLS * sizeof(double));
            memcpy(e11_old, e11, NM * N_SLS * sizeof(double));
            memcpy(e13_old, e13, NM * N_SLS * sizeof(double));

            for (int j = 2; j <= NY; j++)          /* :771-830 */
                for (int i = 1; i <= NX - 1; i++) {
                    double lambda_half_x = 0.5 * (A2(lambda_unrelaxed, i + 1, j) + A2(lambda_unrelaxed, i, j));
                    double mu_half_x = 0.5 * (A2(mu_unrelaxed, i + 1, j) + A2(mu_unrelaxed, i, j));
                    double lambda_plus_mu_half_x = lambda_half_x + mu_half_x;
                    double lambda_plus_two_mu_half_x = lambda_half_x + 2.0 * mu_half_x;

                    double value_dvx_dx = DX_FWD(vx, i, j);
                    double value_dvy_dy = DY_BWD(vy, i, j);

                    M2(memory_dvx_dx, i, j) = P1(b_x_half, i) * M2(memory_dvx_dx, i, j) + P1(a_x_half, i) * value_dvx_dx;
                    M2(memory_dvy_dy, i, j) = P1(b_y, j) * M2(memory_dvy_dy, i, j) + P1(a_y, j) * value_dvy_dy;

                    value_dvx_dx = value_dvx_dx / P1(K_x_half, i) + M2(memory_dvx_dx, i, j);
                    value_dvy_dy = value_dvy_dy / P1(K_y, j) + M2(memory_dvy_dy, i, j);

                    double sum_of_memory_variables_e1 = 0.0;
                    double sum_of_memory_variables_e11 = 0.0;
                    for (int l = 0; l < N_SLS; l++) {
                        M3(e1, i, j, l) = (M3(e1_old, i, j, l) +
                                           (value_dvx_dx + value_dvy_dy) * DELTAT_phi_nu1[l] -
                                           M3(e1_old, i, j, l) * HALF_DELTAT_over_tau_sigma_nu1[l])
                                          * multiplication_factor_tau_sigma_nu1[l];
                        M3(e11, i, j, l) = (M3(e11_old, i, j, l) +
                                            0.5 * (value_dvx_dx - value_dvy_dy) * DELTAT_phi_nu2[l] -
                                            M3(e11_old, i, j, l) * HALF_DELTAT_over_tau_sigma_nu2[l])
                                           * multiplication_factor_tau_sigma_nu2[l];
                        sum_of_memory_variables_e1 = sum_of_memory_variables_e1 + M3(e1, i, j, l) + M3(e1_old, i, j, l);
                        sum_of_memory_variables_e11 = sum_of_memory_variables_e11 + M3(e11, i, j, l) + M3(e11_old, i, j, l);
                    }

                    A2(sigma_xx, i, j) = A2(sigma_xx, i, j) +
                        (lambda_plus_two_mu_half_x * value_dvx_dx + lambda_half_x * value_dvy_dy
                         + (0.5 * lambda_plus_mu_half_x * sum_of_memory_variables_e1 + mu_half_x * sum_of_memory_variables_e11)) * DELTAT;
                    A2(sigma_yy, i, j) = A2(sigma_yy, i, j) +
                        (lambda_half_x * value_dvx_dx + lambda_plus_two_mu_half_x * value_dvy_dy
                         + (0.5 * lambda_plus_mu_half_x * sum_of_memory_variables_e1 - mu_half_x * sum_of_memory_variables_e11)) * DELTAT;
                }

            for (int j = 1; j <= NY - 1; j++)      /* :832-871 */
                for (int i = 2; i <= NX; i++) {
                    double mu_half_y = 0.5 * (A2(mu_unrelaxed, i, j + 1) + A2(mu_unrelaxed, i, j));
                    double value_dvy_dx = DX_BWD(vy, i, j);
                    double value_dvx_dy = DY_FWD(vx, i, j);

                    M2(memory_dvy_dx, i, j) = P1(b_x, i) * M2(memory_dvy_dx, i, j) + P1(a_x, i) * value_dvy_dx;
                    M2(memory_dvx_dy, i, j) = P1(b_y_half, j) * M2(memory_dvx_dy, i, j) + P1(a_y_half, j) * value_dvx_dy;

                    value_dvy_dx = value_dvy_dx / P1(K_x, i) + M2(memory_dvy_dx, i, j);
                    value_dvx_dy = value_dvx_dy / P1(K_y_half, j) + M2(memory_dvx_dy, i, j);

                    double sum_of_memory_variables_e13 = 0.0;
                    for (int l = 0; l < N_SLS; l++) {
                        M3(e13, i, j, l) = (M3(e13_old, i, j, l) +
                                            (value_dvy_dx + value_dvx_dy) * DELTAT_phi_nu2[l] -
                                            M3(e13_old, i, j, l) * HALF_DELTAT_over_tau_sigma_nu2[l])
                                           * multiplication_factor_tau_sigma_nu2[l];
                        sum_of_memory_variables_e13 = sum_of_memory_variables_e13 + M3(e13, i, j, l) + M3(e13_old, i, j, l);
                    }
                    A2(sigma_xy, i, j) = A2(sigma_xy, i, j) + mu_half_y * (value_dvy_dx + value_dvx_dy
                                                                          + 0.5 * sum_of_memory_variables_e13) * DELTAT;
                }
        }

        /* ---- velocity : :879-925 */
        for (int j = 2; j <= NY; j++)
            for (int i = 2; i <= NX; i++) {
                double value_dsigma_xx_dx = DX_BWD(sigma_xx, i, j);
                double value_dsigma_xy_dy = DY_BWD(sigma_xy, i, j);
                M2(memory_dsigma_xx_dx, i, j) = P1(b_x, i) * M2(memory_dsigma_xx_dx, i, j) + P1(a_x, i) * value_dsigma_xx_dx;
                M2(memory_dsigma_xy_dy, i, j) = P1(b_y, j) * M2(memory_dsigma_xy_dy, i, j) + P1(a_y, j) * value_dsigma_xy_dy;
                value_dsigma_xx_dx = value_dsigma_xx_dx / P1(K_x, i) + M2(memory_dsigma_xx_dx, i, j);
                value_dsigma_xy_dy = value_dsigma_xy_dy / P1(K_y, j) + M2(memory_dsigma_xy_dy, i, j);
                A2(vx, i, j) = A2(vx, i, j) + (value_dsigma_xx_dx + value_dsigma_xy_dy) * DELTAT / A2(rho, i, j);
            }
        for (int j = 1; j <= NY - 1; j++)
            for (int i = 1; i <= NX - 1; i++) {
                double rho_half_x_half_y = 0.25 * (A2(rho, i, j) + A2(rho, i + 1, j) + A2(rho, i + 1, j + 1) + A2(rho, i, j + 1));
                double value_dsigma_xy_dx = DX_FWD(sigma_xy, i, j);
                double value_dsigma_yy_dy = DY_FWD(sigma_yy, i, j);
                M2(memory_dsigma_xy_dx, i, j) = P1(b_x_half, i) * M2(memory_dsigma_xy_dx, i, j) + P1(a_x_half, i) * value_dsigma_xy_dx;
                M2(memory_dsigma_yy_dy, i, j) = P1(b_y_half, j) * M2(memory_dsigma_yy_dy, i, j) + P1(a_y_half, j) * value_dsigma_yy_dy;
                value_dsigma_xy_dx = value_dsigma_xy_dx / P1(K_x_half, i) + M2(memory_dsigma_xy_dx, i, j);
                value_dsigma_yy_dy = value_dsigma_yy_dy / P1(K_y_half, j) + M2(memory_dsigma_yy_dy, i, j);
                A2(vy, i, j) = A2(vy, i, j) + (value_dsigma_xy_dx + value_dsigma_yy_dy) * DELTAT / rho_half_x_half_y;
            }

        /* ---- source : :927-972 (force series evaluated by the driver) */
        {
            int i = cfg->isource, j = cfg->jsource;
            double rho_half_x_half_y = 0.25 * (A2(rho, i, j) + A2(rho, i + 1, j) + A2(rho, i + 1, j + 1) + A2(rho, i, j + 1));
            A2(vx, i, j) = A2(vx, i, j) + force_x[it - 1] * DELTAT / A2(rho, i, j);
            A2(vy, i, j) = A2(vy, i, j) + force_y[it - 1] * DELTAT / rho_half_x_half_y;
        }

        /* ---- Dirichlet : :974-985 ; the (:) sections span the ghost ring too */
        for (int j = 0; j <= NY + 1; j++) { A2(vx, 1, j) = 0.0; A2(vx, NX, j) = 0.0; A2(vy, 1, j) = 0.0; A2(vy, NX, j) = 0.0; }
        for (int i = 0; i <= NX + 1; i++) { A2(vx, i, 1) = 0.0; A2(vx, i, NY) = 0.0; A2(vy, i, 1) = 0.0; A2(vy, i, NY) = 0.0; }

        /* ---- seismograms : :987-1035 */
        for (int irec = 1; irec <= NREC; irec++) {
            int i = ix_rec[irec - 1], j = iy_rec[irec - 1];
            sisvx[(size_t)(it - 1) + (size_t)NSTEP * (irec - 1)] = A2(vx, i, j);
            sisvy[(size_t)(it - 1) + (size_t)NSTEP * (irec - 1)] = A2(vy, i, j);
            double lambda_half_x = 0.5 * (A2(lambda_unrelaxed, i + 1, j) + A2(lambda_unrelaxed, i, j));
            double mu_half_x = 0.5 * (A2(mu_unrelaxed, i + 1, j) + A2(mu_unrelaxed, i, j));
            double epsilon_xx = ((lambda_half_x + 2.0 * mu_half_x) * A2(sigma_xx, i, j) - lambda_half_x * A2(sigma_yy, i, j))
                                / (4.0 * mu_half_x * (lambda_half_x + mu_half_x));
            double epsilon_yy = ((lambda_half_x + 2.0 * mu_half_x) * A2(sigma_yy, i, j) - lambda_half_x * A2(sigma_xx, i, j))
                                / (4.0 * mu_half_x * (lambda_half_x + mu_half_x));
            sispressure[(size_t)(it - 1) + (size_t)NSTEP * (irec - 1)] = -(lambda_half_x + TWO_THIRDS * mu_half_x) * (epsilon_xx + epsilon_yy);
        }

        /* ---- energy : :1037-1066 (COMPUTE_ENERGY) */
        if (cfg->compute_energy) {
            double ek = 0.0, ep = 0.0;
            for (int j = NPOINTS_PML + 1; j <= NY - NPOINTS_PML; j++)
                for (int i = NPOINTS_PML + 1; i <= NX - NPOINTS_PML; i++) {
                    double vy_interpolated = 0.25 * (A2(vy, i, j) + A2(vy, i - 1, j) + A2(vy, i - 1, j - 1) + A2(vy, i, j - 1));
                    ek = ek + 0.5 * A2(rho, i, j) * (A2(vx, i, j) * A2(vx, i, j) + vy_interpolated * vy_interpolated);
                }
            for (int j = NPOINTS_PML + 1; j <= NY - NPOINTS_PML; j++)
                for (int i = NPOINTS_PML + 1; i <= NX - NPOINTS_PML; i++) {
                    double lambda_half_x = 0.5 * (A2(lambda_unrelaxed, i + 1, j) + A2(lambda_unrelaxed, i, j));
                    double mu_half_x = 0.5 * (A2(mu_unrelaxed, i + 1, j) + A2(mu_unrelaxed, i, j));
                    double mu_half_y = 0.5 * (A2(mu_unrelaxed, i, j + 1) + A2(mu_unrelaxed, i, j));
                    double epsilon_xx = ((lambda_half_x + 2.0 * mu_half_x) * A2(sigma_xx, i, j) - lambda_half_x * A2(sigma_yy, i, j))
                                        / (4.0 * mu_half_x * (lambda_half_x + mu_half_x));
                    double epsilon_yy = ((lambda_half_x + 2.0 * mu_half_x) * A2(sigma_yy, i, j) - lambda_half_x * A2(sigma_xx, i, j))
                                        / (4.0 * mu_half_x * (lambda_half_x + mu_half_x));
                    double epsilon_xy = A2(sigma_xy, i, j) / (2.0 * mu_half_y);
                    ep = ep + 0.5 * (epsilon_xx * A2(sigma_xx, i, j) + epsilon_yy * A2(sigma_yy, i, j) + 2.0 * epsilon_xy * A2(sigma_xy, i, j));
                }
            energy_kinetic[it - 1] = ek;
            energy_potential[it - 1] = ep;
        }
    }
    oracle_g_loop_seconds = oracle_now_seconds() - t_loop_start;

    if (velocnorm_final) {                         /* :1072 */
        double vmax = 0.0;
        for (size_t s = 0; s < N; s++) {
            double v = sqrt(vx[s] * vx[s] + vy[s] * vy[s]);
            if (v > vmax) vmax = v;
        }
        *velocnorm_final = vmax;
    }
    if (fields_final) {
        double *src[5] = {vx, vy, sigma_xx, sigma_yy, sigma_xy};
        for (int q = 0; q < 5; q++)
            for (int j = 1; j <= NY; j++)
                memcpy(fields_final + (size_t)q * NM + (size_t)NX * (size_t)(j - 1), &A2(src[q], 1, j), (size_t)NX * sizeof(double));
    }
    if (memvar_final) {
        memcpy(memvar_final, e1, NM * N_SLS * sizeof(double));
        memcpy(memvar_final + NM * N_SLS, e11, NM * N_SLS * sizeof(double));
        memcpy(memvar_final + 2 * NM * N_SLS, e13, NM * N_SLS * sizeof(double));
    }
    free(vx); free(vy); free(sigma_xx); free(sigma_yy); free(sigma_xy);
    free(lambda_unrelaxed); free(mu_unrelaxed); free(rho);
    free(memory_dvx_dx); free(memory_dvx_dy); free(memory_dvy_dx); free(memory_dvy_dy);
    free(memory_dsigma_xx_dx); free(memory_dsigma_yy_dy); free(memory_dsigma_xy_dx); free(memory_dsigma_xy_dy);
    free(e1); free(e1_old); free(e11); free(e11_old); free(e13); free(e13_old);
#undef A2
#undef M2
#undef M3
#undef P1
#undef DX_FWD
#undef DX_BWD
#undef DY_FWD
#undef DY_BWD
    return 0;
}
