/*
 * cpml_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 * See cpml_oracle.h for the contract and how parity is pinned (an execution of the reference source, oracle/f90_exec.py).
 *
 * Restates, loop nest by loop nest and operation by operation, the hot paths of
 *   /root/reference/seismic_CPML_2D_isotropic_second_order.f90   (2D-2nd)
 *   /root/reference/seismic_CPML_2D_isotropic_fourth_order.f90   (2D-4th)
 *   /root/reference/seismic_CPML_3D_isotropic_MPI_OpenMP.f90     (3D-iso)
 * Golden build:  gcc -O2 -ffp-contract=off  (no FMA contraction: the reference
 * Makefile:36 builds with plain -O3 for baseline x86-64, which has no FMA).
 * Timed build:   gcc -O3 -march=x86-64-v3 -fopenmp (cpu_baseline in bench.py).
 */
#define _POSIX_C_SOURCE 199309L
#include "cpml_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#if defined(__x86_64__)
#include <xmmintrin.h>
#include <pmmintrin.h>
#endif

#define PI 3.141592653589793238462643 /* 3D-iso :199 */

#include <time.h>
/* shared with cpml_oracle_visco.c (oracle_internal.h) */
int oracle_g_warmup_steps = 0;
double oracle_g_loop_seconds = 0.0;
double oracle_now_seconds(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
void oracle_set_warmup_steps(int w) { oracle_g_warmup_steps = w > 0 ? w : 0; }
double oracle_last_loop_seconds(void) { return oracle_g_loop_seconds; }

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm of bench.py asks for the
 * host's cores explicitly. */
void oracle_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

void oracle_set_ftz(int on)
{
#if defined(__x86_64__) && defined(_OPENMP)
#pragma omp parallel
    {
        _MM_SET_FLUSH_ZERO_MODE(on ? _MM_FLUSH_ZERO_ON : _MM_FLUSH_ZERO_OFF);
        _MM_SET_DENORMALS_ZERO_MODE(on ? _MM_DENORMALS_ZERO_ON : _MM_DENORMALS_ZERO_OFF);
    }
#elif defined(__x86_64__)
    _MM_SET_FLUSH_ZERO_MODE(on ? _MM_FLUSH_ZERO_ON : _MM_FLUSH_ZERO_OFF);
    _MM_SET_DENORMALS_ZERO_MODE(on ? _MM_DENORMALS_ZERO_ON : _MM_DENORMALS_ZERO_OFF);
#else
    (void)on;
#endif
}

/* ------------------------------------------------------------------ setup */

/* 3D-iso :399-667 (x axis :452-525, y :527-596, z :598-667). */
void oracle_pml_profile(int n, double delta, double deltat, int npoints_pml,
                        int use_pml_min, int use_pml_max,
                        double cp, double rcoef, double npower,
                        double k_max_pml, double alpha_max_pml,
                        int origin_top_uses_n, int clamp_alpha,
                        double *a, double *b, double *K,
                        double *a_half, double *b_half, double *K_half)
{
    oracle_pml_profile_visco(n, delta, deltat, npoints_pml, use_pml_min, use_pml_max, cp, 1.0, rcoef,
                             npower, k_max_pml, alpha_max_pml, origin_top_uses_n, clamp_alpha,
                             a, b, K, a_half, b_half, K_half);
}

/* The same text with the viscoelastic program's d0 (3D-visco :547-549):
 *   d0 = -(NPOWER+1) * cp * dsqrt(taumax) * log(Rcoef) / (2 * thickness)
 * sqrt_taumax == 1.0 gives the isotropic formula bit for bit (x * 1.0 == x). */
void oracle_pml_profile_visco(int n, double delta, double deltat, int npoints_pml,
                              int use_pml_min, int use_pml_max,
                              double cp, double sqrt_taumax, double rcoef, double npower,
                              double k_max_pml, double alpha_max_pml,
                              int origin_top_uses_n, int clamp_alpha,
                              double *a, double *b, double *K,
                              double *a_half, double *b_half, double *K_half)
{
    /* :402, :413 ; 3D-visco :547 */
    double thickness = npoints_pml * delta;
    double d0 = -(npower + 1) * cp * sqrt_taumax * log(rcoef) / (2.0 * thickness);
    /* :455-456 (2D-4th :401 for the quirk) */
    double originleft = thickness;
    double originright = (origin_top_uses_n ? n : (n - 1)) * delta - thickness;

    for (int i = 1; i <= n; i++) {
        double d = 0.0, d_half = 0.0;          /* :425-432 */
        double Kv = 1.0, Kv_half = 1.0;
        double alpha = 0.0, alpha_half = 0.0;
        double av = 0.0, av_half = 0.0;
        double val = delta * (double)(i - 1);  /* :461 */
        double abscissa_in_PML, abscissa_normalized;

        if (use_pml_min) {                     /* :464-486 */
            abscissa_in_PML = originleft - val;
            if (abscissa_in_PML >= 0.0) {
                abscissa_normalized = abscissa_in_PML / thickness;
                d = d0 * pow(abscissa_normalized, npower);
                Kv = 1.0 + (k_max_pml - 1.0) * pow(abscissa_normalized, npower);
                alpha = alpha_max_pml * (1.0 - abscissa_normalized);
            }
            abscissa_in_PML = originleft - (val + delta / 2.0);
            if (abscissa_in_PML >= 0.0) {
                abscissa_normalized = abscissa_in_PML / thickness;
                d_half = d0 * pow(abscissa_normalized, npower);
                Kv_half = 1.0 + (k_max_pml - 1.0) * pow(abscissa_normalized, npower);
                alpha_half = alpha_max_pml * (1.0 - abscissa_normalized);
            }
        }
        if (use_pml_max) {                     /* :489-511 */
            abscissa_in_PML = val - originright;
            if (abscissa_in_PML >= 0.0) {
                abscissa_normalized = abscissa_in_PML / thickness;
                d = d0 * pow(abscissa_normalized, npower);
                Kv = 1.0 + (k_max_pml - 1.0) * pow(abscissa_normalized, npower);
                alpha = alpha_max_pml * (1.0 - abscissa_normalized);
            }
            abscissa_in_PML = val + delta / 2.0 - originright;
            if (abscissa_in_PML >= 0.0) {
                abscissa_normalized = abscissa_in_PML / thickness;
                d_half = d0 * pow(abscissa_normalized, npower);
                Kv_half = 1.0 + (k_max_pml - 1.0) * pow(abscissa_normalized, npower);
                alpha_half = alpha_max_pml * (1.0 - abscissa_normalized);
            }
        }
        if (clamp_alpha) {                     /* :514-515 (x axis only) */
            if (alpha < 0.0) alpha = 0.0;
            if (alpha_half < 0.0) alpha_half = 0.0;
        }
        /* :517-518 */
        double bv = exp(-(d / Kv + alpha) * deltat);
        double bv_half = exp(-(d_half / Kv_half + alpha_half) * deltat);
        /* :521-523 */
        if (fabs(d) > 1.e-6) av = d * (bv - 1.0) / (Kv * (d + Kv * alpha));
        if (fabs(d_half) > 1.e-6)
            av_half = d_half * (bv_half - 1.0) / (Kv_half * (d_half + Kv_half * alpha_half));

        a[i - 1] = av; b[i - 1] = bv; K[i - 1] = Kv;
        a_half[i - 1] = av_half; b_half[i - 1] = bv_half; K_half[i - 1] = Kv_half;
    }
}

/* 3D-iso :1058-1071 */
void oracle_source_series(int nstep, double deltat, double f0, double t0,
                          double factor, double angle_force_deg,
                          double *force_x, double *force_y)
{
    const double degrees_to_radians = PI / 180.0; /* :202 */
    for (int it = 1; it <= nstep; it++) {
        double a = PI * PI * f0 * f0;
        double t = (double)(it - 1) * deltat;
        double source_term = -factor * 2.0 * a * (t - t0) * exp(-a * ((t - t0) * (t - t0)));
        force_x[it - 1] = sin(angle_force_deg * degrees_to_radians) * source_term;
        force_y[it - 1] = cos(angle_force_deg * degrees_to_radians) * source_term;
    }
}

/* 3D-iso :683-706 */
void oracle_find_receivers(int nx, int ny, double deltax, double deltay, int nrec,
                           double xdeb, double ydeb, double xfin, double yfin,
                           int *ix_rec, int *iy_rec, double *dist_rec)
{
    const double HUGEVAL = 1.e+30; /* :208 */
    /* nrec == 1 would divide by zero in the reference (:683); keep spacing 0 then */
    double xspacerec = nrec > 1 ? (xfin - xdeb) / (double)(nrec - 1) : 0.0;
    double yspacerec = nrec > 1 ? (yfin - ydeb) / (double)(nrec - 1) : 0.0;
    for (int irec = 1; irec <= nrec; irec++) {
        double xrec = xdeb + (double)(irec - 1) * xspacerec;
        double yrec = ydeb + (double)(irec - 1) * yspacerec;
        double dist = HUGEVAL;
        for (int j = 1; j <= ny; j++) {
            for (int i = 1; i <= nx; i++) {
                double dx = deltax * (double)(i - 1) - xrec;
                double dy = deltay * (double)(j - 1) - yrec;
                double distval = sqrt(dx * dx + dy * dy);
                if (distval < dist) {
                    dist = distval;
                    ix_rec[irec - 1] = i;
                    iy_rec[irec - 1] = j;
                }
            }
        }
        if (dist_rec) dist_rec[irec - 1] = dist;
    }
}

/* ---------------------------------------------------------------- 2-D iso */

int oracle_run_2d(const oracle2d_config *cfg,
                  const double *lambda_in, const double *mu_in, const double *rho_in,
                  const double *a_x, const double *b_x, const double *K_x,
                  const double *a_x_half, const double *b_x_half, const double *K_x_half,
                  const double *a_y, const double *b_y, const double *K_y,
                  const double *a_y_half, const double *b_y_half, const double *K_y_half,
                  const double *force_x, const double *force_y,
                  const int *ix_rec, const int *iy_rec,
                  double *sisvx, double *sisvy,
                  double *energy_kinetic, double *energy_potential,
                  double *vx_final, double *vy_final,
                  double *sigmaxx_final, double *sigmayy_final, double *sigmaxy_final,
                  double *velocnorm_final)
{
    const int NX = cfg->nx, NY = cfg->ny, NSTEP = cfg->nstep, NREC = cfg->nrec;
    const int NPOINTS_PML = cfg->npoints_pml;
    const double DELTAX = cfg->deltax, DELTAY = cfg->deltay, DELTAT = cfg->deltat;
    const int fourth = (cfg->order == 4);
    if (cfg->order != 2 && cfg->order != 4) return 1;

    /* All arrays carry the (0:NX+1,0:NY+1) ghost ring of 2D-4th :205,223; the
     * second-order loops never touch it, so one layout serves both programs. */
    const size_t LD = (size_t)NX + 2;
    const size_t N = LD * ((size_t)NY + 2);
#define A2(arr, i, j) arr[(size_t)(i) + LD * (size_t)(j)]
    double *vx = calloc(N, sizeof(double)), *vy = calloc(N, sizeof(double));
    double *sigmaxx = calloc(N, sizeof(double)), *sigmayy = calloc(N, sizeof(double));
    double *sigmaxy = calloc(N, sizeof(double));
    double *lambda = calloc(N, sizeof(double)), *mu = calloc(N, sizeof(double));
    double *rho = calloc(N, sizeof(double));
    double *memory_dvx_dx = calloc(N, sizeof(double)), *memory_dvx_dy = calloc(N, sizeof(double));
    double *memory_dvy_dx = calloc(N, sizeof(double)), *memory_dvy_dy = calloc(N, sizeof(double));
    double *memory_dsigmaxx_dx = calloc(N, sizeof(double));
    double *memory_dsigmayy_dy = calloc(N, sizeof(double));
    double *memory_dsigmaxy_dx = calloc(N, sizeof(double));
    double *memory_dsigmaxy_dy = calloc(N, sizeof(double));

    /* 1-based views of the profiles */
#define P1(arr, i) arr[(i) - 1]

    for (int j = 1; j <= NY; j++)
        for (int i = 1; i <= NX; i++) {            /* 2D-2nd :468-474 */
            size_t s = (size_t)(i - 1) + (size_t)NX * (size_t)(j - 1);
            A2(rho, i, j) = rho_in[s];
            A2(mu, i, j) = mu_in[s];
            A2(lambda, i, j) = lambda_in[s];
        }
    memset(sisvx, 0, sizeof(double) * (size_t)NSTEP * NREC);   /* :539-544 */
    memset(sisvy, 0, sizeof(double) * (size_t)NSTEP * NREC);
    memset(energy_kinetic, 0, sizeof(double) * (size_t)NSTEP);
    memset(energy_potential, 0, sizeof(double) * (size_t)NSTEP);

    /* energy box: 2D-2nd :695-704 (NPML+1..N-NPML); 2D-4th :696-705 (NPML..N-NPML+1) */
    const int ebx0 = fourth ? NPOINTS_PML : NPOINTS_PML + 1;
    const int ebx1 = fourth ? NX - NPOINTS_PML + 1 : NX - NPOINTS_PML;
    const int eby0 = fourth ? NPOINTS_PML : NPOINTS_PML + 1;
    const int eby1 = fourth ? NY - NPOINTS_PML + 1 : NY - NPOINTS_PML;

    double t_loop_start = oracle_now_seconds();
    for (int it = 1; it <= NSTEP; it++) {          /* 2D-2nd :550 */
        if (it == oracle_g_warmup_steps + 1) t_loop_start = oracle_now_seconds();

        /* ---- sigma_xx, sigma_yy : 2D-2nd :556-580 ; 2D-4th :557-581 */
        for (int j = 2; j <= NY; j++) {
            for (int i = 1; i <= NX - 1; i++) {
                double lambda_half_x = 0.5 * (A2(lambda, i + 1, j) + A2(lambda, i, j));
                double mu_half_x = 0.5 * (A2(mu, i + 1, j) + A2(mu, i, j));
                double lambda_plus_two_mu_half_x = lambda_half_x + 2.0 * mu_half_x;
                double value_dvx_dx, value_dvy_dy;
                if (!fourth) {
                    value_dvx_dx = (A2(vx, i + 1, j) - A2(vx, i, j)) / DELTAX;
                    value_dvy_dy = (A2(vy, i, j) - A2(vy, i, j - 1)) / DELTAY;
                } else {                            /* 2D-4th :565-566 */
                    value_dvx_dx = (27.0 * A2(vx, i + 1, j) - 27.0 * A2(vx, i, j)
                                    - A2(vx, i + 2, j) + A2(vx, i - 1, j)) / (24.0 * DELTAX);
                    value_dvy_dy = (27.0 * A2(vy, i, j) - 27.0 * A2(vy, i, j - 1)
                                    - A2(vy, i, j + 1) + A2(vy, i, j - 2)) / (24.0 * DELTAY);
                }
                A2(memory_dvx_dx, i, j) = P1(b_x_half, i) * A2(memory_dvx_dx, i, j)
                                          + P1(a_x_half, i) * value_dvx_dx;
                A2(memory_dvy_dy, i, j) = P1(b_y, j) * A2(memory_dvy_dy, i, j)
                                          + P1(a_y, j) * value_dvy_dy;
                value_dvx_dx = value_dvx_dx / P1(K_x_half, i) + A2(memory_dvx_dx, i, j);
                value_dvy_dy = value_dvy_dy / P1(K_y, j) + A2(memory_dvy_dy, i, j);
                A2(sigmaxx, i, j) = A2(sigmaxx, i, j)
                    + (lambda_plus_two_mu_half_x * value_dvx_dx + lambda_half_x * value_dvy_dy) * DELTAT;
                A2(sigmayy, i, j) = A2(sigmayy, i, j)
                    + (lambda_half_x * value_dvx_dx + lambda_plus_two_mu_half_x * value_dvy_dy) * DELTAT;
            }
        }

        /* ---- sigma_xy : 2D-2nd :582-600 ; 2D-4th :583-601 */
        for (int j = 1; j <= NY - 1; j++) {
            for (int i = 2; i <= NX; i++) {
                double mu_half_y = 0.5 * (A2(mu, i, j + 1) + A2(mu, i, j));
                double value_dvy_dx, value_dvx_dy;
                if (!fourth) {
                    value_dvy_dx = (A2(vy, i, j) - A2(vy, i - 1, j)) / DELTAX;
                    value_dvx_dy = (A2(vx, i, j + 1) - A2(vx, i, j)) / DELTAY;
                } else {                            /* 2D-4th :589-590 */
                    value_dvy_dx = (27.0 * A2(vy, i, j) - 27.0 * A2(vy, i - 1, j)
                                    - A2(vy, i + 1, j) + A2(vy, i - 2, j)) / (24.0 * DELTAX);
                    value_dvx_dy = (27.0 * A2(vx, i, j + 1) - 27.0 * A2(vx, i, j)
                                    - A2(vx, i, j + 2) + A2(vx, i, j - 1)) / (24.0 * DELTAY);
                }
                A2(memory_dvy_dx, i, j) = P1(b_x, i) * A2(memory_dvy_dx, i, j)
                                          + P1(a_x, i) * value_dvy_dx;
                A2(memory_dvx_dy, i, j) = P1(b_y_half, j) * A2(memory_dvx_dy, i, j)
                                          + P1(a_y_half, j) * value_dvx_dy;
                value_dvy_dx = value_dvy_dx / P1(K_x, i) + A2(memory_dvy_dx, i, j);
                /* quirk B3: 2D-4th :596 divides by K_y(j), 2D-2nd :595 by K_y_half(j) */
                if (!fourth)
                    value_dvx_dy = value_dvx_dy / P1(K_y_half, j) + A2(memory_dvx_dy, i, j);
                else
                    value_dvx_dy = value_dvx_dy / P1(K_y, j) + A2(memory_dvx_dy, i, j);
                A2(sigmaxy, i, j) = A2(sigmaxy, i, j)
                    + mu_half_y * (value_dvy_dx + value_dvx_dy) * DELTAT;
            }
        }

        /* ---- vx : 2D-2nd :606-621 ; 2D-4th :607-622 */
        for (int j = 2; j <= NY; j++) {
            for (int i = 2; i <= NX; i++) {
                double value_dsigmaxx_dx, value_dsigmaxy_dy;
                if (!fourth) {
                    value_dsigmaxx_dx = (A2(sigmaxx, i, j) - A2(sigmaxx, i - 1, j)) / DELTAX;
                    value_dsigmaxy_dy = (A2(sigmaxy, i, j) - A2(sigmaxy, i, j - 1)) / DELTAY;
                } else {                            /* 2D-4th :610-611 */
                    value_dsigmaxx_dx = (27.0 * A2(sigmaxx, i, j) - 27.0 * A2(sigmaxx, i - 1, j)
                                         - A2(sigmaxx, i + 1, j) + A2(sigmaxx, i - 2, j)) / (24.0 * DELTAX);
                    value_dsigmaxy_dy = (27.0 * A2(sigmaxy, i, j) - 27.0 * A2(sigmaxy, i, j - 1)
                                         - A2(sigmaxy, i, j + 1) + A2(sigmaxy, i, j - 2)) / (24.0 * DELTAY);
                }
                A2(memory_dsigmaxx_dx, i, j) = P1(b_x, i) * A2(memory_dsigmaxx_dx, i, j)
                                               + P1(a_x, i) * value_dsigmaxx_dx;
                A2(memory_dsigmaxy_dy, i, j) = P1(b_y, j) * A2(memory_dsigmaxy_dy, i, j)
                                               + P1(a_y, j) * value_dsigmaxy_dy;
                value_dsigmaxx_dx = value_dsigmaxx_dx / P1(K_x, i) + A2(memory_dsigmaxx_dx, i, j);
                value_dsigmaxy_dy = value_dsigmaxy_dy / P1(K_y, j) + A2(memory_dsigmaxy_dy, i, j);
                A2(vx, i, j) = A2(vx, i, j)
                    + (value_dsigmaxx_dx + value_dsigmaxy_dy) * DELTAT / A2(rho, i, j);
            }
        }

        /* ---- vy : 2D-2nd :623-641 ; 2D-4th :624-642 */
        for (int j = 1; j <= NY - 1; j++) {
            for (int i = 1; i <= NX - 1; i++) {
                double rho_half_x_half_y = 0.25 * (A2(rho, i, j) + A2(rho, i + 1, j)
                                                   + A2(rho, i + 1, j + 1) + A2(rho, i, j + 1));
                double value_dsigmaxy_dx, value_dsigmayy_dy;
                if (!fourth) {
                    value_dsigmaxy_dx = (A2(sigmaxy, i + 1, j) - A2(sigmaxy, i, j)) / DELTAX;
                    value_dsigmayy_dy = (A2(sigmayy, i, j + 1) - A2(sigmayy, i, j)) / DELTAY;
                } else {                            /* 2D-4th :631-632 */
                    value_dsigmaxy_dx = (27.0 * A2(sigmaxy, i + 1, j) - 27.0 * A2(sigmaxy, i, j)
                                         - A2(sigmaxy, i + 2, j) + A2(sigmaxy, i - 1, j)) / (24.0 * DELTAX);
                    value_dsigmayy_dy = (27.0 * A2(sigmayy, i, j + 1) - 27.0 * A2(sigmayy, i, j)
                                         - A2(sigmayy, i, j + 2) + A2(sigmayy, i, j - 1)) / (24.0 * DELTAY);
                }
                A2(memory_dsigmaxy_dx, i, j) = P1(b_x_half, i) * A2(memory_dsigmaxy_dx, i, j)
                                               + P1(a_x_half, i) * value_dsigmaxy_dx;
                A2(memory_dsigmayy_dy, i, j) = P1(b_y_half, j) * A2(memory_dsigmayy_dy, i, j)
                                               + P1(a_y_half, j) * value_dsigmayy_dy;
                value_dsigmaxy_dx = value_dsigmaxy_dx / P1(K_x_half, i) + A2(memory_dsigmaxy_dx, i, j);
                value_dsigmayy_dy = value_dsigmayy_dy / P1(K_y_half, j) + A2(memory_dsigmayy_dy, i, j);
                A2(vy, i, j) = A2(vy, i, j)
                    + (value_dsigmaxy_dx + value_dsigmayy_dy) * DELTAT / rho_half_x_half_y;
            }
        }

        /* ---- source : 2D-2nd :643-667 (force series precomputed by the driver) */
        {
            int i = cfg->isource, j = cfg->jsource;
            double rho_half_x_half_y = 0.25 * (A2(rho, i, j) + A2(rho, i + 1, j)
                                               + A2(rho, i + 1, j + 1) + A2(rho, i, j + 1));
            A2(vx, i, j) = A2(vx, i, j) + force_x[it - 1] * DELTAT / A2(rho, i, j);
            A2(vy, i, j) = A2(vy, i, j) + force_y[it - 1] * DELTAT / rho_half_x_half_y;
        }

        /* ---- Dirichlet : 2D-2nd :669-680 ; 2D-4th :670-681.  In the fourth-order
         * program "vx(1,:)" spans j = 0..NY+1 and "vx(:,1)" spans i = 0..NX+1; the
         * ghost ring is zero anyway. */
        for (int j = 0; j <= NY + 1; j++) {
            A2(vx, 1, j) = 0.0; A2(vx, NX, j) = 0.0;
            A2(vy, 1, j) = 0.0; A2(vy, NX, j) = 0.0;
        }
        for (int i = 0; i <= NX + 1; i++) {
            A2(vx, i, 1) = 0.0; A2(vx, i, NY) = 0.0;
            A2(vy, i, 1) = 0.0; A2(vy, i, NY) = 0.0;
        }

        /* ---- seismograms : 2D-2nd :682-686 */
        for (int irec = 1; irec <= NREC; irec++) {
            sisvx[(size_t)(it - 1) + (size_t)NSTEP * (irec - 1)] = A2(vx, ix_rec[irec - 1], iy_rec[irec - 1]);
            sisvy[(size_t)(it - 1) + (size_t)NSTEP * (irec - 1)] = A2(vy, ix_rec[irec - 1], iy_rec[irec - 1]);
        }

        /* ---- energy : 2D-2nd :688-713.  Fortran SUM over the array section is
         * evaluated here in array-element order (i fastest). */
        {
            double ek = 0.0, ep = 0.0;
            for (int j = eby0; j <= eby1; j++)
                for (int i = ebx0; i <= ebx1; i++)
                    ek += A2(rho, i, j) * (A2(vx, i, j) * A2(vx, i, j) + A2(vy, i, j) * A2(vy, i, j));
            energy_kinetic[it - 1] = 0.5 * ek;
            for (int j = eby0; j <= eby1; j++) {
                for (int i = ebx0; i <= ebx1; i++) {
                    double l = A2(lambda, i, j), m = A2(mu, i, j);
                    double epsilon_xx = ((l + 2.0 * m) * A2(sigmaxx, i, j) - l * A2(sigmayy, i, j))
                                        / (4.0 * m * (l + m));
                    double epsilon_yy = ((l + 2.0 * m) * A2(sigmayy, i, j) - l * A2(sigmaxx, i, j))
                                        / (4.0 * m * (l + m));
                    double epsilon_xy = A2(sigmaxy, i, j) / (2.0 * m);
                    ep = ep + 0.5 * (epsilon_xx * A2(sigmaxx, i, j) + epsilon_yy * A2(sigmayy, i, j)
                                     + 2.0 * epsilon_xy * A2(sigmaxy, i, j));
                }
            }
            energy_potential[it - 1] = ep;
        }
    }

    oracle_g_loop_seconds = oracle_now_seconds() - t_loop_start;

    if (velocnorm_final) {                          /* 2D-2nd :719 */
        double vmax = 0.0;
        for (size_t s = 0; s < N; s++) {
            double v = sqrt(vx[s] * vx[s] + vy[s] * vy[s]);
            if (v > vmax) vmax = v;
        }
        *velocnorm_final = vmax;
    }
    for (int j = 1; j <= NY; j++)
        for (int i = 1; i <= NX; i++) {
            size_t s = (size_t)(i - 1) + (size_t)NX * (size_t)(j - 1);
            if (vx_final) vx_final[s] = A2(vx, i, j);
            if (vy_final) vy_final[s] = A2(vy, i, j);
            if (sigmaxx_final) sigmaxx_final[s] = A2(sigmaxx, i, j);
            if (sigmayy_final) sigmayy_final[s] = A2(sigmayy, i, j);
            if (sigmaxy_final) sigmaxy_final[s] = A2(sigmaxy, i, j);
        }

    free(vx); free(vy); free(sigmaxx); free(sigmayy); free(sigmaxy);
    free(lambda); free(mu); free(rho);
    free(memory_dvx_dx); free(memory_dvx_dy); free(memory_dvy_dx); free(memory_dvy_dy);
    free(memory_dsigmaxx_dx); free(memory_dsigmayy_dy);
    free(memory_dsigmaxy_dx); free(memory_dsigmaxy_dy);
#undef A2
    return 0;
}

/* ---------------------------------------------------------------- 3-D iso */

/* One emulated MPI rank: the arrays declared at 3D-iso :222-240 and :273. */
typedef struct {
    double *vx, *vy, *vz, *sigmaxx, *sigmayy, *sigmazz, *sigmaxy, *sigmaxz, *sigmayz;
    double *memory_dvx_dx, *memory_dvx_dy, *memory_dvx_dz;
    double *memory_dvy_dx, *memory_dvy_dy, *memory_dvy_dz;
    double *memory_dvz_dx, *memory_dvz_dy, *memory_dvz_dz;
    double *memory_dsigmaxx_dx, *memory_dsigmayy_dy, *memory_dsigmazz_dz;
    double *memory_dsigmaxy_dx, *memory_dsigmaxy_dy;
    double *memory_dsigmaxz_dx, *memory_dsigmaxz_dz;
    double *memory_dsigmayz_dy, *memory_dsigmayz_dz;
} slab_t;

static double *zalloc_par(size_t n)
{
    double *p = malloc(n * sizeof(double));
    if (!p) return NULL;
    /* parallel first touch so that the timed build places pages near the threads */
#pragma omp parallel for schedule(static)
    for (long long s = 0; s < (long long)n; s++) p[s] = 0.0;
    return p;
}

int oracle_run_3d_iso(const oracle3d_config *cfg,
                      const double *a_x, const double *b_x, const double *K_x,
                      const double *a_x_half, const double *b_x_half, const double *K_x_half,
                      const double *a_y, const double *b_y, const double *K_y,
                      const double *a_y_half, const double *b_y_half, const double *K_y_half,
                      const double *a_z, const double *b_z, const double *K_z,
                      const double *a_z_half, const double *b_z_half, const double *K_z_half,
                      const double *force_x, const double *force_y,
                      const int *ix_rec, const int *iy_rec,
                      double *sisvx, double *sisvy, double *total_energy,
                      double *plane_vx, double *plane_vy,
                      double *fields_final, double *vnorm_final)
{
    const int NX = cfg->nx, NY = cfg->ny, NZ = cfg->nz, NPROC = cfg->nproc;
    const int NSTEP = cfg->nstep, NREC = cfg->nrec, NPOINTS_PML = cfg->npoints_pml;
    /* topology checks, 3D-iso :381-394 (evenness relaxed for nproc == 1, which the
     * reference cannot run; the source then sits at global k = NZ/2 all the same) */
    if (NPROC < 1) return 1;
    if (NPROC > 1 && NPROC % 2 != 0) return 2;
    if (NZ % NPROC != 0) return 3;
    const int NZ_LOCAL = NZ / NPROC;
    if (NZ_LOCAL < NPOINTS_PML) return 4;
    if (NZ % 2 != 0) return 5;                     /* :126 "even number" */

    const double ONE_OVER_DELTAX = 1.0 / cfg->deltax;   /* :134-136 */
    const double ONE_OVER_DELTAY = 1.0 / cfg->deltay;
    const double ONE_OVER_DELTAZ = 1.0 / cfg->deltaz;
    const double lambda = cfg->lambda, mu = cfg->mu, rho = cfg->rho;
    const double DELTAT = cfg->deltat;
    const double lambdaplustwomu = cfg->lambdaplustwomu;   /* rho*cp*cp, :144 */
    const double DELTAT_lambda = DELTAT * lambda;          /* :296-300 */
    const double DELTAT_mu = DELTAT * mu;
    const double DELTAT_lambdaplus2mu = DELTAT * lambdaplustwomu;
    const double DELTAT_over_rho = DELTAT / rho;

    const size_t PLANE = (size_t)NX * NY;
    const size_t NF = PLANE * ((size_t)NZ_LOCAL + 2);      /* (NX,NY,0:NZ_LOCAL+1) */
    const size_t NM = PLANE * (size_t)NZ_LOCAL;            /* (NX,NY,NZ_LOCAL)     */
#define F3(arr, i, j, k) arr[(size_t)((i) - 1) + (size_t)NX * ((size_t)((j) - 1) + (size_t)NY * (size_t)(k))]
#define M3(arr, i, j, k) arr[(size_t)((i) - 1) + (size_t)NX * ((size_t)((j) - 1) + (size_t)NY * (size_t)((k) - 1))]
#define P1(arr, i) arr[(i) - 1]

    slab_t *S = calloc((size_t)NPROC, sizeof(slab_t));
    for (int r = 0; r < NPROC; r++) {
        /* :720-749 (quirk B1: sigmaxx is never zeroed by the reference; static storage is zero) */
        double **f = (double **)&S[r];
        for (int q = 0; q < 9; q++) f[q] = zalloc_par(NF);
        for (int q = 9; q < 27; q++) f[q] = zalloc_par(NM);
        for (int q = 0; q < 27; q++) if (!f[q]) return 6;
    }
    memset(sisvx, 0, sizeof(double) * (size_t)NSTEP * NREC);   /* :751-756 */
    memset(sisvy, 0, sizeof(double) * (size_t)NSTEP * NREC);
    memset(total_energy, 0, sizeof(double) * (size_t)NSTEP);

    /* :346 ; with nproc == 1 the cut plane is the middle of the only slab */
    const int rank_cut_plane = NPROC / 2 - 1;
    const int src_rank = NPROC > 1 ? rank_cut_plane : 0;
    const int src_klocal = NPROC > 1 ? NZ_LOCAL : NZ / 2;

    double t_loop_start = oracle_now_seconds();
    for (int it = 1; it <= NSTEP; it++) {          /* :802 */
        if (it == oracle_g_warmup_steps + 1) t_loop_start = oracle_now_seconds();

        /* ---- halo exchange of v : :810-823.  MPI_SENDRECV with MPI_PROC_NULL at the
         * ends leaves the end halos untouched (zero). */
        for (int r = 0; r < NPROC; r++) {
            if (r + 1 < NPROC) {
                /* vx(:,:,1) of rank r+1 -> vx(:,:,NZ_LOCAL+1) of rank r  (left shift) */
                memcpy(&F3(S[r].vx, 1, 1, NZ_LOCAL + 1), &F3(S[r + 1].vx, 1, 1, 1), PLANE * sizeof(double));
                memcpy(&F3(S[r].vy, 1, 1, NZ_LOCAL + 1), &F3(S[r + 1].vy, 1, 1, 1), PLANE * sizeof(double));
                /* vz(:,:,NZ_LOCAL) of rank r -> vz(:,:,0) of rank r+1  (right shift) */
                memcpy(&F3(S[r + 1].vz, 1, 1, 0), &F3(S[r].vz, 1, 1, NZ_LOCAL), PLANE * sizeof(double));
            }
        }

        for (int r = 0; r < NPROC; r++) {
            slab_t *s = &S[r];
            const int offset_k = r * NZ_LOCAL;                         /* :397 */
            const int k2begin = (r == 0) ? 2 : 1;                      /* :792-793 */
            const int kminus1end = (r == NPROC - 1) ? NZ_LOCAL - 1 : NZ_LOCAL; /* :795-796 */

            /* ---- sigmaxx, sigmayy, sigmazz : :836-863 */
#pragma omp parallel for schedule(static)
            for (int k = k2begin; k <= NZ_LOCAL; k++) {
                int kglobal = k + offset_k;
                for (int j = 2; j <= NY; j++) {
                    for (int i = 1; i <= NX - 1; i++) {
                        double value_dvx_dx = (F3(s->vx, i + 1, j, k) - F3(s->vx, i, j, k)) * ONE_OVER_DELTAX;
                        double value_dvy_dy = (F3(s->vy, i, j, k) - F3(s->vy, i, j - 1, k)) * ONE_OVER_DELTAY;
                        double value_dvz_dz = (F3(s->vz, i, j, k) - F3(s->vz, i, j, k - 1)) * ONE_OVER_DELTAZ;

                        M3(s->memory_dvx_dx, i, j, k) = P1(b_x_half, i) * M3(s->memory_dvx_dx, i, j, k) + P1(a_x_half, i) * value_dvx_dx;
                        M3(s->memory_dvy_dy, i, j, k) = P1(b_y, j) * M3(s->memory_dvy_dy, i, j, k) + P1(a_y, j) * value_dvy_dy;
                        M3(s->memory_dvz_dz, i, j, k) = P1(b_z, kglobal) * M3(s->memory_dvz_dz, i, j, k) + P1(a_z, kglobal) * value_dvz_dz;

                        value_dvx_dx = value_dvx_dx / P1(K_x_half, i) + M3(s->memory_dvx_dx, i, j, k);
                        value_dvy_dy = value_dvy_dy / P1(K_y, j) + M3(s->memory_dvy_dy, i, j, k);
                        value_dvz_dz = value_dvz_dz / P1(K_z, kglobal) + M3(s->memory_dvz_dz, i, j, k);

                        F3(s->sigmaxx, i, j, k) = DELTAT_lambdaplus2mu * value_dvx_dx
                            + DELTAT_lambda * (value_dvy_dy + value_dvz_dz) + F3(s->sigmaxx, i, j, k);
                        F3(s->sigmayy, i, j, k) = DELTAT_lambda * (value_dvx_dx + value_dvz_dz)
                            + DELTAT_lambdaplus2mu * value_dvy_dy + F3(s->sigmayy, i, j, k);
                        F3(s->sigmazz, i, j, k) = DELTAT_lambda * (value_dvx_dx + value_dvy_dy)
                            + DELTAT_lambdaplus2mu * value_dvz_dz + F3(s->sigmazz, i, j, k);
                    }
                }
            }

            /* ---- sigmaxy : :877-894 */
#pragma omp parallel for schedule(static)
            for (int k = 1; k <= NZ_LOCAL; k++) {
                for (int j = 1; j <= NY - 1; j++) {
                    for (int i = 2; i <= NX; i++) {
                        double value_dvy_dx = (F3(s->vy, i, j, k) - F3(s->vy, i - 1, j, k)) * ONE_OVER_DELTAX;
                        double value_dvx_dy = (F3(s->vx, i, j + 1, k) - F3(s->vx, i, j, k)) * ONE_OVER_DELTAY;

                        M3(s->memory_dvy_dx, i, j, k) = P1(b_x, i) * M3(s->memory_dvy_dx, i, j, k) + P1(a_x, i) * value_dvy_dx;
                        M3(s->memory_dvx_dy, i, j, k) = P1(b_y_half, j) * M3(s->memory_dvx_dy, i, j, k) + P1(a_y_half, j) * value_dvx_dy;

                        value_dvy_dx = value_dvy_dx / P1(K_x, i) + M3(s->memory_dvy_dx, i, j, k);
                        value_dvx_dy = value_dvx_dy / P1(K_y_half, j) + M3(s->memory_dvx_dy, i, j, k);

                        F3(s->sigmaxy, i, j, k) = DELTAT_mu * (value_dvy_dx + value_dvx_dy) + F3(s->sigmaxy, i, j, k);
                    }
                }
            }

            /* ---- sigmaxz, sigmayz : :908-943 */
#pragma omp parallel for schedule(static)
            for (int k = 1; k <= kminus1end; k++) {
                int kglobal = k + offset_k;
                for (int j = 1; j <= NY; j++) {
                    for (int i = 2; i <= NX; i++) {
                        double value_dvz_dx = (F3(s->vz, i, j, k) - F3(s->vz, i - 1, j, k)) * ONE_OVER_DELTAX;
                        double value_dvx_dz = (F3(s->vx, i, j, k + 1) - F3(s->vx, i, j, k)) * ONE_OVER_DELTAZ;

                        M3(s->memory_dvz_dx, i, j, k) = P1(b_x, i) * M3(s->memory_dvz_dx, i, j, k) + P1(a_x, i) * value_dvz_dx;
                        M3(s->memory_dvx_dz, i, j, k) = P1(b_z_half, kglobal) * M3(s->memory_dvx_dz, i, j, k) + P1(a_z_half, kglobal) * value_dvx_dz;

                        value_dvz_dx = value_dvz_dx / P1(K_x, i) + M3(s->memory_dvz_dx, i, j, k);
                        value_dvx_dz = value_dvx_dz / P1(K_z_half, kglobal) + M3(s->memory_dvx_dz, i, j, k);

                        F3(s->sigmaxz, i, j, k) = DELTAT_mu * (value_dvz_dx + value_dvx_dz) + F3(s->sigmaxz, i, j, k);
                    }
                }
                for (int j = 1; j <= NY - 1; j++) {
                    for (int i = 1; i <= NX; i++) {
                        double value_dvz_dy = (F3(s->vz, i, j + 1, k) - F3(s->vz, i, j, k)) * ONE_OVER_DELTAY;
                        double value_dvy_dz = (F3(s->vy, i, j, k + 1) - F3(s->vy, i, j, k)) * ONE_OVER_DELTAZ;

                        M3(s->memory_dvz_dy, i, j, k) = P1(b_y_half, j) * M3(s->memory_dvz_dy, i, j, k) + P1(a_y_half, j) * value_dvz_dy;
                        M3(s->memory_dvy_dz, i, j, k) = P1(b_z_half, kglobal) * M3(s->memory_dvy_dz, i, j, k) + P1(a_z_half, kglobal) * value_dvy_dz;

                        value_dvz_dy = value_dvz_dy / P1(K_y_half, j) + M3(s->memory_dvz_dy, i, j, k);
                        value_dvy_dz = value_dvy_dz / P1(K_z_half, kglobal) + M3(s->memory_dvy_dz, i, j, k);

                        F3(s->sigmayz, i, j, k) = DELTAT_mu * (value_dvz_dy + value_dvy_dz) + F3(s->sigmayz, i, j, k);
                    }
                }
            }
        }

        /* ---- halo exchange of sigma : :950-963 */
        for (int r = 0; r < NPROC; r++) {
            if (r + 1 < NPROC) {
                memcpy(&F3(S[r].sigmazz, 1, 1, NZ_LOCAL + 1), &F3(S[r + 1].sigmazz, 1, 1, 1), PLANE * sizeof(double));
                memcpy(&F3(S[r + 1].sigmayz, 1, 1, 0), &F3(S[r].sigmayz, 1, 1, NZ_LOCAL), PLANE * sizeof(double));
                memcpy(&F3(S[r + 1].sigmaxz, 1, 1, 0), &F3(S[r].sigmaxz, 1, 1, NZ_LOCAL), PLANE * sizeof(double));
            }
        }

        double energy_sum = 0.0;

        for (int r = 0; r < NPROC; r++) {
            slab_t *s = &S[r];
            const int offset_k = r * NZ_LOCAL;
            const int k2begin = (r == 0) ? 2 : 1;
            const int kminus1end = (r == NPROC - 1) ? NZ_LOCAL - 1 : NZ_LOCAL;

            /* ---- vx, vy : :976-1017 */
#pragma omp parallel for schedule(static)
            for (int k = k2begin; k <= NZ_LOCAL; k++) {
                int kglobal = k + offset_k;
                for (int j = 2; j <= NY; j++) {
                    for (int i = 2; i <= NX; i++) {
                        double value_dsigmaxx_dx = (F3(s->sigmaxx, i, j, k) - F3(s->sigmaxx, i - 1, j, k)) * ONE_OVER_DELTAX;
                        double value_dsigmaxy_dy = (F3(s->sigmaxy, i, j, k) - F3(s->sigmaxy, i, j - 1, k)) * ONE_OVER_DELTAY;
                        double value_dsigmaxz_dz = (F3(s->sigmaxz, i, j, k) - F3(s->sigmaxz, i, j, k - 1)) * ONE_OVER_DELTAZ;

                        M3(s->memory_dsigmaxx_dx, i, j, k) = P1(b_x, i) * M3(s->memory_dsigmaxx_dx, i, j, k) + P1(a_x, i) * value_dsigmaxx_dx;
                        M3(s->memory_dsigmaxy_dy, i, j, k) = P1(b_y, j) * M3(s->memory_dsigmaxy_dy, i, j, k) + P1(a_y, j) * value_dsigmaxy_dy;
                        M3(s->memory_dsigmaxz_dz, i, j, k) = P1(b_z, kglobal) * M3(s->memory_dsigmaxz_dz, i, j, k) + P1(a_z, kglobal) * value_dsigmaxz_dz;

                        value_dsigmaxx_dx = value_dsigmaxx_dx / P1(K_x, i) + M3(s->memory_dsigmaxx_dx, i, j, k);
                        value_dsigmaxy_dy = value_dsigmaxy_dy / P1(K_y, j) + M3(s->memory_dsigmaxy_dy, i, j, k);
                        value_dsigmaxz_dz = value_dsigmaxz_dz / P1(K_z, kglobal) + M3(s->memory_dsigmaxz_dz, i, j, k);

                        F3(s->vx, i, j, k) = DELTAT_over_rho * (value_dsigmaxx_dx + value_dsigmaxy_dy + value_dsigmaxz_dz) + F3(s->vx, i, j, k);
                    }
                }
                for (int j = 1; j <= NY - 1; j++) {
                    for (int i = 1; i <= NX - 1; i++) {
                        double value_dsigmaxy_dx = (F3(s->sigmaxy, i + 1, j, k) - F3(s->sigmaxy, i, j, k)) * ONE_OVER_DELTAX;
                        double value_dsigmayy_dy = (F3(s->sigmayy, i, j + 1, k) - F3(s->sigmayy, i, j, k)) * ONE_OVER_DELTAY;
                        double value_dsigmayz_dz = (F3(s->sigmayz, i, j, k) - F3(s->sigmayz, i, j, k - 1)) * ONE_OVER_DELTAZ;

                        M3(s->memory_dsigmaxy_dx, i, j, k) = P1(b_x_half, i) * M3(s->memory_dsigmaxy_dx, i, j, k) + P1(a_x_half, i) * value_dsigmaxy_dx;
                        M3(s->memory_dsigmayy_dy, i, j, k) = P1(b_y_half, j) * M3(s->memory_dsigmayy_dy, i, j, k) + P1(a_y_half, j) * value_dsigmayy_dy;
                        M3(s->memory_dsigmayz_dz, i, j, k) = P1(b_z, kglobal) * M3(s->memory_dsigmayz_dz, i, j, k) + P1(a_z, kglobal) * value_dsigmayz_dz;

                        value_dsigmaxy_dx = value_dsigmaxy_dx / P1(K_x_half, i) + M3(s->memory_dsigmaxy_dx, i, j, k);
                        value_dsigmayy_dy = value_dsigmayy_dy / P1(K_y_half, j) + M3(s->memory_dsigmayy_dy, i, j, k);
                        value_dsigmayz_dz = value_dsigmayz_dz / P1(K_z, kglobal) + M3(s->memory_dsigmayz_dz, i, j, k);

                        F3(s->vy, i, j, k) = DELTAT_over_rho * (value_dsigmaxy_dx + value_dsigmayy_dy + value_dsigmayz_dz) + F3(s->vy, i, j, k);
                    }
                }
            }

            /* ---- vz : :1031-1052 */
#pragma omp parallel for schedule(static)
            for (int k = 1; k <= kminus1end; k++) {
                int kglobal = k + offset_k;
                for (int j = 2; j <= NY; j++) {
                    for (int i = 1; i <= NX - 1; i++) {
                        double value_dsigmaxz_dx = (F3(s->sigmaxz, i + 1, j, k) - F3(s->sigmaxz, i, j, k)) * ONE_OVER_DELTAX;
                        double value_dsigmayz_dy = (F3(s->sigmayz, i, j, k) - F3(s->sigmayz, i, j - 1, k)) * ONE_OVER_DELTAY;
                        double value_dsigmazz_dz = (F3(s->sigmazz, i, j, k + 1) - F3(s->sigmazz, i, j, k)) * ONE_OVER_DELTAZ;

                        M3(s->memory_dsigmaxz_dx, i, j, k) = P1(b_x_half, i) * M3(s->memory_dsigmaxz_dx, i, j, k) + P1(a_x_half, i) * value_dsigmaxz_dx;
                        M3(s->memory_dsigmayz_dy, i, j, k) = P1(b_y, j) * M3(s->memory_dsigmayz_dy, i, j, k) + P1(a_y, j) * value_dsigmayz_dy;
                        M3(s->memory_dsigmazz_dz, i, j, k) = P1(b_z_half, kglobal) * M3(s->memory_dsigmazz_dz, i, j, k) + P1(a_z_half, kglobal) * value_dsigmazz_dz;

                        value_dsigmaxz_dx = value_dsigmaxz_dx / P1(K_x_half, i) + M3(s->memory_dsigmaxz_dx, i, j, k);
                        value_dsigmayz_dy = value_dsigmayz_dy / P1(K_y, j) + M3(s->memory_dsigmayz_dy, i, j, k);
                        value_dsigmazz_dz = value_dsigmazz_dz / P1(K_z_half, kglobal) + M3(s->memory_dsigmazz_dz, i, j, k);

                        F3(s->vz, i, j, k) = DELTAT_over_rho * (value_dsigmaxz_dx + value_dsigmayz_dy + value_dsigmazz_dz) + F3(s->vz, i, j, k);
                    }
                }
            }

            /* ---- source : :1055-1083 */
            if (r == src_rank) {
                int i = cfg->isource, j = cfg->jsource;
                F3(s->vx, i, j, src_klocal) = F3(s->vx, i, j, src_klocal) + force_x[it - 1] * DELTAT / rho;
                F3(s->vy, i, j, src_klocal) = F3(s->vy, i, j, src_klocal) + force_y[it - 1] * DELTAT / rho;
            }

            /* ---- Dirichlet : :1087-1121 ; the (:,:) sections span k = 0..NZ_LOCAL+1 */
#pragma omp parallel for schedule(static)
            for (int k = 0; k <= NZ_LOCAL + 1; k++) {
                for (int j = 1; j <= NY; j++) {
                    F3(s->vx, 1, j, k) = 0.0; F3(s->vy, 1, j, k) = 0.0; F3(s->vz, 1, j, k) = 0.0;
                    F3(s->vx, NX, j, k) = 0.0; F3(s->vy, NX, j, k) = 0.0; F3(s->vz, NX, j, k) = 0.0;
                }
                for (int i = 1; i <= NX; i++) {
                    F3(s->vx, i, 1, k) = 0.0; F3(s->vy, i, 1, k) = 0.0; F3(s->vz, i, 1, k) = 0.0;
                    F3(s->vx, i, NY, k) = 0.0; F3(s->vy, i, NY, k) = 0.0; F3(s->vz, i, NY, k) = 0.0;
                }
            }
            if (r == 0)
                for (size_t q = 0; q < PLANE; q++) {
                    (&F3(s->vx, 1, 1, 1))[q] = 0.0; (&F3(s->vy, 1, 1, 1))[q] = 0.0; (&F3(s->vz, 1, 1, 1))[q] = 0.0;
                }
            if (r == NPROC - 1)
                for (size_t q = 0; q < PLANE; q++) {
                    (&F3(s->vx, 1, 1, NZ_LOCAL))[q] = 0.0; (&F3(s->vy, 1, 1, NZ_LOCAL))[q] = 0.0;
                    (&F3(s->vz, 1, 1, NZ_LOCAL))[q] = 0.0;
                }

            /* ---- seismograms : :1123-1129 */
            if (r == src_rank)
                for (int irec = 1; irec <= NREC; irec++) {
                    sisvx[(size_t)(it - 1) + (size_t)NSTEP * (irec - 1)] = F3(s->vx, ix_rec[irec - 1], iy_rec[irec - 1], src_klocal);
                    sisvy[(size_t)(it - 1) + (size_t)NSTEP * (irec - 1)] = F3(s->vy, ix_rec[irec - 1], iy_rec[irec - 1], src_klocal);
                }

            /* ---- energy : :1131-1180 */
            {
                double total_energy_kinetic = 0.0, total_energy_potential = 0.0;
                int kmin = 1, kmax = NZ_LOCAL;
                if (r == 0) kmin = NPOINTS_PML + 1;
                if (r == NPROC - 1) kmax = NZ_LOCAL - NPOINTS_PML;
#pragma omp parallel for schedule(static) reduction(+ : total_energy_kinetic, total_energy_potential)
                for (int k = kmin; k <= kmax; k++) {
                    for (int j = NPOINTS_PML + 1; j <= NY - NPOINTS_PML; j++) {
                        for (int i = NPOINTS_PML + 1; i <= NX - NPOINTS_PML; i++) {
                            double vxv = F3(s->vx, i, j, k), vyv = F3(s->vy, i, j, k), vzv = F3(s->vz, i, j, k);
                            double sxx = F3(s->sigmaxx, i, j, k), syy = F3(s->sigmayy, i, j, k), szz = F3(s->sigmazz, i, j, k);
                            double sxy = F3(s->sigmaxy, i, j, k), sxz = F3(s->sigmaxz, i, j, k), syz = F3(s->sigmayz, i, j, k);
                            total_energy_kinetic = total_energy_kinetic + 0.5 * rho * (vxv * vxv + vyv * vyv + vzv * vzv);

                            double epsilon_xx = (2.0 * (lambda + mu) * sxx - lambda * syy - lambda * szz) / (2.0 * mu * (3.0 * lambda + 2.0 * mu));
                            double epsilon_yy = (2.0 * (lambda + mu) * syy - lambda * sxx - lambda * szz) / (2.0 * mu * (3.0 * lambda + 2.0 * mu));
                            double epsilon_zz = (2.0 * (lambda + mu) * szz - lambda * sxx - lambda * syy) / (2.0 * mu * (3.0 * lambda + 2.0 * mu));
                            double epsilon_xy = sxy / (2.0 * mu);
                            double epsilon_xz = sxz / (2.0 * mu);
                            double epsilon_yz = syz / (2.0 * mu);

                            if (cfg->energy_bug_compat)   /* :1169-1172, quirk B2 */
                                total_energy_potential = total_energy_potential
                                    + 0.5 * (epsilon_xx * sxx + epsilon_yy * syy + epsilon_yy * syy
                                             + 2.0 * epsilon_xy * sxy + 2.0 * epsilon_xz * sxz + 2.0 * epsilon_yz * syz);
                            else
                                total_energy_potential = total_energy_potential
                                    + 0.5 * (epsilon_xx * sxx + epsilon_yy * syy + epsilon_zz * szz
                                             + 2.0 * epsilon_xy * sxy + 2.0 * epsilon_xz * sxz + 2.0 * epsilon_yz * syz);
                        }
                    }
                }
                energy_sum += total_energy_kinetic + total_energy_potential;   /* MPI_REDUCE(SUM) :1179 */
            }
        }
        total_energy[it - 1] = energy_sum;
    }
    oracle_g_loop_seconds = oracle_now_seconds() - t_loop_start;

    /* ---- results */
    if (plane_vx) memcpy(plane_vx, &F3(S[src_rank].vx, 1, 1, src_klocal), PLANE * sizeof(double)); /* :1236 */
    if (plane_vy) memcpy(plane_vy, &F3(S[src_rank].vy, 1, 1, src_klocal), PLANE * sizeof(double));
    if (vnorm_final) {                             /* :1185 */
        double vmax = 0.0;
        for (int r = 0; r < NPROC; r++)
            for (int k = 1; k <= NZ_LOCAL; k++)
                for (size_t q = 0; q < PLANE; q++) {
                    double a = (&F3(S[r].vx, 1, 1, k))[q], b = (&F3(S[r].vy, 1, 1, k))[q], c = (&F3(S[r].vz, 1, 1, k))[q];
                    double v = sqrt(a * a + b * b + c * c);
                    if (v > vmax) vmax = v;
                }
        *vnorm_final = vmax;
    }
    if (fields_final) {
        const size_t G = PLANE * (size_t)NZ;
        for (int r = 0; r < NPROC; r++) {
            double **f = (double **)&S[r];
            for (int q = 0; q < 9; q++)
                memcpy(fields_final + (size_t)q * G + PLANE * (size_t)(r * NZ_LOCAL),
                       f[q] + PLANE, PLANE * (size_t)NZ_LOCAL * sizeof(double));
        }
    }
    for (int r = 0; r < NPROC; r++) {
        double **f = (double **)&S[r];
        for (int q = 0; q < 27; q++) free(f[q]);
    }
    free(S);
#undef F3
#undef M3
#undef P1
    return 0;
}
