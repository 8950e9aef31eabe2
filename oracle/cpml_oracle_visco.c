/*
 * cpml_oracle_visco.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 * See cpml_oracle.h for the contract and how parity is pinned (an execution of the reference source, oracle/f90_exec.py).
 *
 * Restates, loop nest by loop nest and operation by operation, the hot path of
 *   /root/reference/seismic_CPML_3D_viscoelastic_MPI.f90   (3D-visco)
 * fourth-order staggered grid, N_SLS = 2 standard linear solids (Carcione 1993), C-PML with
 * K_MAX_PML = 7, z-slab MPI decomposition emulated as nproc slabs in one address space.
 * The program's MPI exchange sends only half of the z halo its fourth-order stencils read
 * (SURVEY.md quirk B6): the slots that are never received stay zero, exactly as here.
 */
#include "cpml_oracle.h"
#include "oracle_internal.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* 3D-visco :839-853 */
void oracle_find_receivers_visco(int nx, int ny, double deltax, double deltay, int nrec,
                                 const double *xrec, const double *yrec,
                                 int *ix_rec, int *iy_rec, double *dist_rec)
{
    const double HUGEVAL = 1.e+30; /* :226 */
    for (int irec = 1; irec <= nrec; irec++) {
        double dist = HUGEVAL;
        for (int j = 1; j <= ny; j++) {
            for (int i = 1; i <= nx; i++) {
                double dx = deltax * (double)i - xrec[irec - 1];
                double dy = deltay * (double)j - yrec[irec - 1];
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

/* One emulated MPI rank: the arrays declared at 3D-visco :246-264 and :301-303. */
typedef struct {
    double *vx, *vy, *vz, *sigmaxx, *sigmayy, *sigmazz, *sigmaxy, *sigmaxz, *sigmayz;
    double *sigmaxx_R, *sigmayy_R, *sigmazz_R, *sigmaxy_R, *sigmaxz_R, *sigmayz_R;
    double *memory_dvx_dx, *memory_dvx_dy, *memory_dvx_dz;
    double *memory_dvy_dx, *memory_dvy_dy, *memory_dvy_dz;
    double *memory_dvz_dx, *memory_dvz_dy, *memory_dvz_dz;
    double *memory_dsigmaxx_dx, *memory_dsigmayy_dy, *memory_dsigmazz_dz;
    double *memory_dsigmaxy_dx, *memory_dsigmaxy_dy;
    double *memory_dsigmaxz_dx, *memory_dsigmaxz_dz;
    double *memory_dsigmayz_dy, *memory_dsigmayz_dz;
    double *e1, *e11, *e22, *e12, *e13, *e23;     /* (N_SLS, ...) : SLS index fastest */
} vslab_t;
#define NV_PLAIN 33
#define NV_ALL 39

static double *zalloc_par(size_t n)
{
    double *p = malloc(n * sizeof(double));
    if (!p) return NULL;
#pragma omp parallel for schedule(static)
    for (long long s = 0; s < (long long)n; s++) p[s] = 0.0;
    return p;
}

int oracle_run_3d_visco(const oraclev3d_config *cfg,
                        const double *a_x, const double *b_x, const double *K_x,
                        const double *a_x_half, const double *b_x_half, const double *K_x_half,
                        const double *a_y, const double *b_y, const double *K_y,
                        const double *a_y_half, const double *b_y_half, const double *K_y_half,
                        const double *a_z, const double *b_z, const double *K_z,
                        const double *a_z_half, const double *b_z_half, const double *K_z_half,
                        const double *force_x, const double *force_y,
                        const int *ix_rec, const int *iy_rec,
                        double *sisvx, double *sisvy,
                        double *energy_total, double *energy_kinetic, double *energy_potential,
                        double *fields_final, double *vnorm_final)
{
    const int NX = cfg->nx, NY = cfg->ny, NZ = cfg->nz, NPROC = cfg->nproc;
    const int NSTEP = cfg->nstep, NREC = cfg->nrec, NPOINTS_PML = cfg->npoints_pml;
    /* topology checks, 3D-visco :518-528 (evenness relaxed for nproc == 1 as in the 3D-iso oracle) */
    if (NPROC < 1) return 1;
    if (NPROC > 1 && NPROC % 2 != 0) return 2;
    if (NZ % NPROC != 0) return 3;
    const int NZ_LOCAL = NZ / NPROC;
    if (NZ_LOCAL < NPOINTS_PML) return 4;
    if (NZ % 2 != 0) return 5;
    if (NZ_LOCAL < 2) return 6;

    const double ONE = 1.0, TWO = 2.0, DIM = 3.0;             /* :165 */
    const double ONE_OVER_DELTAX = 1.0 / cfg->deltax;         /* :162-164 */
    const double ONE_OVER_DELTAY = 1.0 / cfg->deltay;
    const double ONE_OVER_DELTAZ = 1.0 / cfg->deltaz;
    const double lambda = cfg->lambda, mu = cfg->mu, rho = cfg->rho;
    const double DELTAT = cfg->deltat, deltat = cfg->deltat;
    const double DELTAT_over_rho = DELTAT / rho;              /* :337 */

    /* :458-477 */
    const double *tau_epsilon_nu1 = cfg->tau_epsilon_nu1, *tau_sigma_nu1 = cfg->tau_sigma_nu1;
    const double *tau_epsilon_nu2 = cfg->tau_epsilon_nu2, *tau_sigma_nu2 = cfg->tau_sigma_nu2;
    double inv_tau_sigma_nu1[2], inv_tau_sigma_nu2[2], phi_nu1[2], phi_nu2[2];
    inv_tau_sigma_nu1[0] = ONE / tau_sigma_nu1[0];
    inv_tau_sigma_nu2[0] = ONE / tau_sigma_nu2[0];
    inv_tau_sigma_nu1[1] = ONE / tau_sigma_nu1[1];
    inv_tau_sigma_nu2[1] = ONE / tau_sigma_nu2[1];
    phi_nu1[0] = (ONE - tau_epsilon_nu1[0] / tau_sigma_nu1[0]) / tau_sigma_nu1[0];
    phi_nu2[0] = (ONE - tau_epsilon_nu2[0] / tau_sigma_nu2[0]) / tau_sigma_nu2[0];
    phi_nu1[1] = (ONE - tau_epsilon_nu1[1] / tau_sigma_nu1[1]) / tau_sigma_nu1[1];
    phi_nu2[1] = (ONE - tau_epsilon_nu2[1] / tau_sigma_nu2[1]) / tau_sigma_nu2[1];
    const double Mu_nu1 = ONE - (ONE - tau_epsilon_nu1[0] / tau_sigma_nu1[0]) - (ONE - tau_epsilon_nu1[1] / tau_sigma_nu1[1]);
    const double Mu_nu2 = ONE - (ONE - tau_epsilon_nu2[0] / tau_sigma_nu2[0]) - (ONE - tau_epsilon_nu2[1] / tau_sigma_nu2[1]);

    const size_t LDX = (size_t)NX + 2, LDY = (size_t)NY + 2;
    const size_t PLANE = LDX * LDY;
    const size_t NF = PLANE * ((size_t)NZ_LOCAL + 4);          /* (0:NX+1,0:NY+1,-1:NZ_LOCAL+2) */
#define IDX(i, j, k) ((size_t)(i) + LDX * ((size_t)(j) + LDY * (size_t)((k) + 1)))
#define F(arr, i, j, k) arr[IDX(i, j, k)]
#define E(arr, l, i, j, k) arr[(size_t)((l) - 1) + 2 * IDX(i, j, k)]
#define P1(arr, i) arr[(i) - 1]

    vslab_t *S = calloc((size_t)NPROC, sizeof(vslab_t));
    if (!S) return 7;
    for (int r = 0; r < NPROC; r++) {
        /* :860-897 (quirk B1: sigmaxx and the sigma*_R are never zeroed; static storage is zero) */
        double **f = (double **)&S[r];
        for (int q = 0; q < NV_PLAIN; q++) f[q] = zalloc_par(NF);
        for (int q = NV_PLAIN; q < NV_ALL; q++) f[q] = zalloc_par(2 * NF);
        for (int q = 0; q < NV_ALL; q++) if (!f[q]) return 7;
    }
    memset(sisvx, 0, sizeof(double) * (size_t)NSTEP * NREC);   /* :899-906 */
    memset(sisvy, 0, sizeof(double) * (size_t)NSTEP * NREC);
    memset(energy_total, 0, sizeof(double) * (size_t)NSTEP);
    memset(energy_kinetic, 0, sizeof(double) * (size_t)NSTEP);
    memset(energy_potential, 0, sizeof(double) * (size_t)NSTEP);

    const int rank_cut_plane = NPROC / 2 - 1;                   /* :480 */
    const int src_rank = NPROC > 1 ? rank_cut_plane : 0;
    const int src_klocal = NPROC > 1 ? NZ_LOCAL : NZ / 2;
    const size_t TWOPLANES = 2 * PLANE * sizeof(double);        /* number_of_values, :374 */

    double t_loop_start = oracle_now_seconds();
    for (int it = 1; it <= NSTEP; it++) {                       /* :954 */
        if (it == oracle_g_warmup_steps + 1) t_loop_start = oracle_now_seconds();

        /* ---- halo exchange of v : :962-975 (two planes each) */
        for (int r = 0; r + 1 < NPROC; r++) {
            /* vx(:,:,1:2) of rank r+1 -> vx(:,:,NZ_LOCAL+1:NZ_LOCAL+2) of rank r (left shift) */
            memcpy(&F(S[r].vx, 0, 0, NZ_LOCAL + 1), &F(S[r + 1].vx, 0, 0, 1), TWOPLANES);
            memcpy(&F(S[r].vy, 0, 0, NZ_LOCAL + 1), &F(S[r + 1].vy, 0, 0, 1), TWOPLANES);
            /* vz(:,:,NZ_LOCAL-1:NZ_LOCAL) of rank r -> vz(:,:,-1:0) of rank r+1 (right shift) */
            memcpy(&F(S[r + 1].vz, 0, 0, -1), &F(S[r].vz, 0, 0, NZ_LOCAL - 1), TWOPLANES);
            if (cfg->complete_halos) {   /* NOT in the reference (quirk B6) */
                memcpy(&F(S[r + 1].vx, 0, 0, 0), &F(S[r].vx, 0, 0, NZ_LOCAL), PLANE * sizeof(double));
                memcpy(&F(S[r + 1].vy, 0, 0, 0), &F(S[r].vy, 0, 0, NZ_LOCAL), PLANE * sizeof(double));
                memcpy(&F(S[r].vz, 0, 0, NZ_LOCAL + 1), &F(S[r + 1].vz, 0, 0, 1), PLANE * sizeof(double));
            }
        }

        for (int r = 0; r < NPROC; r++) {
            vslab_t *s = &S[r];
            const int offset_k = r * NZ_LOCAL;                                   /* :531 */
            const int k2begin = (r == 0) ? 2 : 1;                                /* :942-943 */
            const int kminus1end = (r == NPROC - 1) ? NZ_LOCAL - 1 : NZ_LOCAL;   /* :945-946 */

            /* ---- sigmaxx, sigmayy, sigmazz, e1, e11, e22 : :977-1096 */
#pragma omp parallel for schedule(static)
            for (int k = k2begin; k <= NZ_LOCAL; k++) {
                int kglobal = k + offset_k;
                for (int j = 2; j <= NY; j++) {
                    for (int i = 1; i <= NX - 1; i++) {
                        double mul_relaxed = mu;                                  /* :982-987 */
                        double lambdal_relaxed = lambda;
                        double lambdalplus2mul_relaxed = lambdal_relaxed + TWO * mul_relaxed;
                        double lambdal_unrelaxed = (lambdal_relaxed + 2.0 / DIM * mul_relaxed) * Mu_nu1 - 2.0 / DIM * mul_relaxed * Mu_nu2;
                        double mul_unrelaxed = mul_relaxed * Mu_nu2;
                        double lambdalplus2mul_unrelaxed = lambdal_unrelaxed + TWO * mul_unrelaxed;

                        /* :989-991 */
                        double value_dvx_dx = (27.0 * F(s->vx, i + 1, j, k) - 27.0 * F(s->vx, i, j, k) - F(s->vx, i + 2, j, k) + F(s->vx, i - 1, j, k)) * ONE_OVER_DELTAX / 24.0;
                        double value_dvy_dy = (27.0 * F(s->vy, i, j, k) - 27.0 * F(s->vy, i, j - 1, k) - F(s->vy, i, j + 1, k) + F(s->vy, i, j - 2, k)) * ONE_OVER_DELTAY / 24.0;
                        double value_dvz_dz = (27.0 * F(s->vz, i, j, k) - 27.0 * F(s->vz, i, j, k - 1) - F(s->vz, i, j, k + 1) + F(s->vz, i, j, k - 2)) * ONE_OVER_DELTAZ / 24.0;

                        /* :993-995 */
                        F(s->memory_dvx_dx, i, j, k) = P1(b_x_half, i) * F(s->memory_dvx_dx, i, j, k) + P1(a_x_half, i) * value_dvx_dx;
                        F(s->memory_dvy_dy, i, j, k) = P1(b_y, j) * F(s->memory_dvy_dy, i, j, k) + P1(a_y, j) * value_dvy_dy;
                        F(s->memory_dvz_dz, i, j, k) = P1(b_z, kglobal) * F(s->memory_dvz_dz, i, j, k) + P1(a_z, kglobal) * value_dvz_dz;

                        /* :997-999 */
                        double duxdx = value_dvx_dx / P1(K_x_half, i) + F(s->memory_dvx_dx, i, j, k);
                        double duydy = value_dvy_dy / P1(K_y, j) + F(s->memory_dvy_dy, i, j, k);
                        double duzdz = value_dvz_dz / P1(K_z, kglobal) + F(s->memory_dvz_dz, i, j, k);

                        double div = duxdx + duydy + duzdz;                       /* :1001 */

                        double tauinv, Un, Sn, tauinvUn, Unp1;
                        /* evolution e1(1), e1(2) : :1003-1017 */
                        for (int l = 1; l <= 2; l++) {
                            tauinv = -inv_tau_sigma_nu1[l - 1];
                            Un = E(s->e1, l, i, j, k);
                            Sn = div * phi_nu1[l - 1];
                            tauinvUn = tauinv * Un;
                            Unp1 = (Un + deltat * (Sn + 0.5 * tauinvUn)) / (1.0 - deltat * 0.5 * tauinv);
                            E(s->e1, l, i, j, k) = Unp1;
                        }
                        /* evolution e11(1), e11(2) : :1019-1033 */
                        for (int l = 1; l <= 2; l++) {
                            tauinv = -inv_tau_sigma_nu2[l - 1];
                            Un = E(s->e11, l, i, j, k);
                            Sn = (duxdx - div / DIM) * phi_nu2[l - 1];
                            tauinvUn = tauinv * Un;
                            Unp1 = (Un + deltat * (Sn + 0.5 * tauinvUn)) / (1.0 - deltat * 0.5 * tauinv);
                            E(s->e11, l, i, j, k) = Unp1;
                        }
                        /* evolution e22(1), e22(2) : :1035-1049 */
                        for (int l = 1; l <= 2; l++) {
                            tauinv = -inv_tau_sigma_nu2[l - 1];
                            Un = E(s->e22, l, i, j, k);
                            Sn = (duydy - div / DIM) * phi_nu2[l - 1];
                            tauinvUn = tauinv * Un;
                            Unp1 = (Un + deltat * (Sn + 0.5 * tauinvUn)) / (1.0 - deltat * 0.5 * tauinv);
                            E(s->e22, l, i, j, k) = Unp1;
                        }

                        /* memory variables with the relaxed parameters : :1054-1060 */
                        F(s->sigmaxx, i, j, k) = F(s->sigmaxx, i, j, k) + deltat * ((lambdal_relaxed + 2.0 / DIM * mul_relaxed) *
                            (E(s->e1, 1, i, j, k) + E(s->e1, 2, i, j, k)) + TWO * mul_relaxed * (E(s->e11, 1, i, j, k) + E(s->e11, 2, i, j, k)));
                        F(s->sigmayy, i, j, k) = F(s->sigmayy, i, j, k) + deltat * ((lambdal_relaxed + 2.0 / DIM * mul_relaxed) *
                            (E(s->e1, 1, i, j, k) + E(s->e1, 2, i, j, k)) + TWO * mul_relaxed * (E(s->e22, 1, i, j, k) + E(s->e22, 2, i, j, k)));
                        if (cfg->sigmazz_isotropic)   /* NOT the reference (quirk B14): the isotropic form, for the analytical check */
                            F(s->sigmazz, i, j, k) = F(s->sigmazz, i, j, k) + deltat * ((lambdal_relaxed + 2.0 / DIM * mul_relaxed) *
                                (E(s->e1, 1, i, j, k) + E(s->e1, 2, i, j, k)) - TWO * mul_relaxed * (E(s->e11, 1, i, j, k) + E(s->e11, 2, i, j, k)
                                + E(s->e22, 1, i, j, k) + E(s->e22, 2, i, j, k)));
                        else
                        F(s->sigmazz, i, j, k) = F(s->sigmazz, i, j, k) + deltat * ((lambdal_relaxed + 2.0 * mul_relaxed) *
                            (E(s->e1, 1, i, j, k) + E(s->e1, 2, i, j, k)) - TWO / DIM * mul_relaxed * (E(s->e11, 1, i, j, k) + E(s->e11, 2, i, j, k)
                            + E(s->e22, 1, i, j, k) + E(s->e22, 2, i, j, k)));

                        /* unrelaxed Lame parameters : :1064-1077 */
                        F(s->sigmaxx, i, j, k) = F(s->sigmaxx, i, j, k) +
                            (lambdalplus2mul_unrelaxed * (duxdx) + lambdal_unrelaxed * (duydy) + lambdal_unrelaxed * (duzdz)) * DELTAT;
                        F(s->sigmayy, i, j, k) = F(s->sigmayy, i, j, k) +
                            (lambdal_unrelaxed * (duxdx) + lambdalplus2mul_unrelaxed * (duydy) + lambdal_unrelaxed * (duzdz)) * DELTAT;
                        F(s->sigmazz, i, j, k) = F(s->sigmazz, i, j, k) +
                            (lambdal_unrelaxed * (duxdx) + lambdal_unrelaxed * (duydy) + lambdalplus2mul_unrelaxed * (duzdz)) * DELTAT;

                        /* relaxed stresses, used by the energy only : :1079-1092 */
                        F(s->sigmaxx_R, i, j, k) = F(s->sigmaxx_R, i, j, k) +
                            (lambdalplus2mul_relaxed * (duxdx) + lambdal_relaxed * (duydy) + lambdal_relaxed * (duzdz)) * DELTAT;
                        F(s->sigmayy_R, i, j, k) = F(s->sigmayy_R, i, j, k) +
                            (lambdal_relaxed * (duxdx) + lambdalplus2mul_relaxed * (duydy) + lambdal_relaxed * (duzdz)) * DELTAT;
                        F(s->sigmazz_R, i, j, k) = F(s->sigmazz_R, i, j, k) +
                            (lambdal_relaxed * (duxdx) + lambdal_relaxed * (duydy) + lambdalplus2mul_relaxed * (duzdz)) * DELTAT;
                    }
                }
            }

            /* ---- sigmaxy, e12 : :1098-1139 */
#pragma omp parallel for schedule(static)
            for (int k = 1; k <= NZ_LOCAL; k++) {
                for (int j = 1; j <= NY - 1; j++) {
                    for (int i = 2; i <= NX; i++) {
                        double mul_relaxed = mu;
                        double mul_unrelaxed = mul_relaxed * Mu_nu2;

                        double value_dvy_dx = (27.0 * F(s->vy, i, j, k) - 27.0 * F(s->vy, i - 1, j, k) - F(s->vy, i + 1, j, k) + F(s->vy, i - 2, j, k)) * ONE_OVER_DELTAX / 24.0;
                        double value_dvx_dy = (27.0 * F(s->vx, i, j + 1, k) - 27.0 * F(s->vx, i, j, k) - F(s->vx, i, j + 2, k) + F(s->vx, i, j - 1, k)) * ONE_OVER_DELTAY / 24.0;

                        F(s->memory_dvy_dx, i, j, k) = P1(b_x, i) * F(s->memory_dvy_dx, i, j, k) + P1(a_x, i) * value_dvy_dx;
                        F(s->memory_dvx_dy, i, j, k) = P1(b_y_half, j) * F(s->memory_dvx_dy, i, j, k) + P1(a_y_half, j) * value_dvx_dy;

                        double duydx = value_dvy_dx / P1(K_x, i) + F(s->memory_dvy_dx, i, j, k);
                        double duxdy = value_dvx_dy / P1(K_y_half, j) + F(s->memory_dvx_dy, i, j, k);

                        for (int l = 1; l <= 2; l++) {                            /* :1113-1127 */
                            double tauinv = -inv_tau_sigma_nu2[l - 1];
                            double Un = E(s->e12, l, i, j, k);
                            double Sn = (duxdy + duydx) * phi_nu2[l - 1];
                            double tauinvUn = tauinv * Un;
                            double Unp1 = (Un + deltat * (Sn + 0.5 * tauinvUn)) / (1.0 - deltat * 0.5 * tauinv);
                            E(s->e12, l, i, j, k) = Unp1;
                        }

                        F(s->sigmaxy, i, j, k) = F(s->sigmaxy, i, j, k) + deltat * mul_relaxed * (E(s->e12, 1, i, j, k) + E(s->e12, 2, i, j, k));
                        F(s->sigmaxy, i, j, k) = F(s->sigmaxy, i, j, k) + mul_unrelaxed * (duxdy + duydx) * DELTAT;
                        F(s->sigmaxy_R, i, j, k) = F(s->sigmaxy_R, i, j, k) + mul_relaxed * (duxdy + duydx) * DELTAT;
                    }
                }
            }

            /* ---- sigmaxz, e13 and sigmayz, e23 : :1141-1223 */
#pragma omp parallel for schedule(static)
            for (int k = 1; k <= kminus1end; k++) {
                int kglobal = k + offset_k;
                for (int j = 1; j <= NY; j++) {
                    for (int i = 2; i <= NX; i++) {
                        double mul_relaxed = mu;
                        double mul_unrelaxed = mul_relaxed * Mu_nu2;

                        double value_dvz_dx = (27.0 * F(s->vz, i, j, k) - 27.0 * F(s->vz, i - 1, j, k) - F(s->vz, i + 1, j, k) + F(s->vz, i - 2, j, k)) * ONE_OVER_DELTAX / 24.0;
                        double value_dvx_dz = (27.0 * F(s->vx, i, j, k + 1) - 27.0 * F(s->vx, i, j, k) - F(s->vx, i, j, k + 2) + F(s->vx, i, j, k - 1)) * ONE_OVER_DELTAZ / 24.0;

                        F(s->memory_dvz_dx, i, j, k) = P1(b_x, i) * F(s->memory_dvz_dx, i, j, k) + P1(a_x, i) * value_dvz_dx;
                        F(s->memory_dvx_dz, i, j, k) = P1(b_z_half, kglobal) * F(s->memory_dvx_dz, i, j, k) + P1(a_z_half, kglobal) * value_dvx_dz;

                        double duzdx = value_dvz_dx / P1(K_x, i) + F(s->memory_dvz_dx, i, j, k);
                        double duxdz = value_dvx_dz / P1(K_z_half, kglobal) + F(s->memory_dvx_dz, i, j, k);

                        for (int l = 1; l <= 2; l++) {                            /* :1157-1171 */
                            double tauinv = -inv_tau_sigma_nu2[l - 1];
                            double Un = E(s->e13, l, i, j, k);
                            double Sn = (duxdz + duzdx) * phi_nu2[l - 1];
                            double tauinvUn = tauinv * Un;
                            double Unp1 = (Un + deltat * (Sn + 0.5 * tauinvUn)) / (1.0 - deltat * 0.5 * tauinv);
                            E(s->e13, l, i, j, k) = Unp1;
                        }

                        F(s->sigmaxz, i, j, k) = F(s->sigmaxz, i, j, k) + deltat * mul_relaxed * (E(s->e13, 1, i, j, k) + E(s->e13, 2, i, j, k));
                        F(s->sigmaxz, i, j, k) = F(s->sigmaxz, i, j, k) + mul_unrelaxed * (duxdz + duzdx) * DELTAT;
                        F(s->sigmaxz_R, i, j, k) = F(s->sigmaxz_R, i, j, k) + mul_relaxed * (duxdz + duzdx) * DELTAT;
                    }
                }
                for (int j = 1; j <= NY - 1; j++) {
                    for (int i = 1; i <= NX; i++) {
                        double mul_relaxed = mu;
                        double mul_unrelaxed = mul_relaxed * Mu_nu2;

                        double value_dvz_dy = (27.0 * F(s->vz, i, j + 1, k) - 27.0 * F(s->vz, i, j, k) - F(s->vz, i, j + 2, k) + F(s->vz, i, j - 1, k)) * ONE_OVER_DELTAY / 24.0;
                        double value_dvy_dz = (27.0 * F(s->vy, i, j, k + 1) - 27.0 * F(s->vy, i, j, k) - F(s->vy, i, j, k + 2) + F(s->vy, i, j, k - 1)) * ONE_OVER_DELTAZ / 24.0;

                        F(s->memory_dvz_dy, i, j, k) = P1(b_y_half, j) * F(s->memory_dvz_dy, i, j, k) + P1(a_y_half, j) * value_dvz_dy;
                        F(s->memory_dvy_dz, i, j, k) = P1(b_z_half, kglobal) * F(s->memory_dvy_dz, i, j, k) + P1(a_z_half, kglobal) * value_dvy_dz;

                        double duzdy = value_dvz_dy / P1(K_y_half, j) + F(s->memory_dvz_dy, i, j, k);
                        double duydz = value_dvy_dz / P1(K_z_half, kglobal) + F(s->memory_dvy_dz, i, j, k);

                        for (int l = 1; l <= 2; l++) {                            /* :1197-1211 */
                            double tauinv = -inv_tau_sigma_nu2[l - 1];
                            double Un = E(s->e23, l, i, j, k);
                            double Sn = (duydz + duzdy) * phi_nu2[l - 1];
                            double tauinvUn = tauinv * Un;
                            double Unp1 = (Un + deltat * (Sn + 0.5 * tauinvUn)) / (1.0 - deltat * 0.5 * tauinv);
                            E(s->e23, l, i, j, k) = Unp1;
                        }

                        F(s->sigmayz, i, j, k) = F(s->sigmayz, i, j, k) + deltat * mul_relaxed * (E(s->e23, 1, i, j, k) + E(s->e23, 2, i, j, k));
                        F(s->sigmayz, i, j, k) = F(s->sigmayz, i, j, k) + mul_unrelaxed * (duydz + duzdy) * DELTAT;
                        F(s->sigmayz_R, i, j, k) = F(s->sigmayz_R, i, j, k) + mul_relaxed * (duydz + duzdy) * DELTAT;
                    }
                }
            }
        }

        /* ---- halo exchange of sigma : :1229-1242 */
        for (int r = 0; r + 1 < NPROC; r++) {
            memcpy(&F(S[r].sigmazz, 0, 0, NZ_LOCAL + 1), &F(S[r + 1].sigmazz, 0, 0, 1), TWOPLANES);
            memcpy(&F(S[r + 1].sigmayz, 0, 0, -1), &F(S[r].sigmayz, 0, 0, NZ_LOCAL - 1), TWOPLANES);
            memcpy(&F(S[r + 1].sigmaxz, 0, 0, -1), &F(S[r].sigmaxz, 0, 0, NZ_LOCAL - 1), TWOPLANES);
            if (cfg->complete_halos) {   /* NOT in the reference (quirk B6) */
                memcpy(&F(S[r + 1].sigmazz, 0, 0, 0), &F(S[r].sigmazz, 0, 0, NZ_LOCAL), PLANE * sizeof(double));
                memcpy(&F(S[r].sigmayz, 0, 0, NZ_LOCAL + 1), &F(S[r + 1].sigmayz, 0, 0, 1), PLANE * sizeof(double));
                memcpy(&F(S[r].sigmaxz, 0, 0, NZ_LOCAL + 1), &F(S[r + 1].sigmaxz, 0, 0, 1), PLANE * sizeof(double));
            }
        }

        double sum_total = 0.0, sum_kinetic = 0.0, sum_potential = 0.0;

        for (int r = 0; r < NPROC; r++) {
            vslab_t *s = &S[r];
            const int offset_k = r * NZ_LOCAL;
            const int k2begin = (r == 0) ? 2 : 1;
            const int kminus1end = (r == NPROC - 1) ? NZ_LOCAL - 1 : NZ_LOCAL;

            /* ---- vx, vy : :1244-1285 */
#pragma omp parallel for schedule(static)
            for (int k = k2begin; k <= NZ_LOCAL; k++) {
                int kglobal = k + offset_k;
                for (int j = 2; j <= NY; j++) {
                    for (int i = 2; i <= NX; i++) {
                        double value_dsigmaxx_dx = (27.0 * F(s->sigmaxx, i, j, k) - 27.0 * F(s->sigmaxx, i - 1, j, k) - F(s->sigmaxx, i + 1, j, k) + F(s->sigmaxx, i - 2, j, k)) * ONE_OVER_DELTAX / 24.0;
                        double value_dsigmaxy_dy = (27.0 * F(s->sigmaxy, i, j, k) - 27.0 * F(s->sigmaxy, i, j - 1, k) - F(s->sigmaxy, i, j + 1, k) + F(s->sigmaxy, i, j - 2, k)) * ONE_OVER_DELTAY / 24.0;
                        double value_dsigmaxz_dz = (27.0 * F(s->sigmaxz, i, j, k) - 27.0 * F(s->sigmaxz, i, j, k - 1) - F(s->sigmaxz, i, j, k + 1) + F(s->sigmaxz, i, j, k - 2)) * ONE_OVER_DELTAZ / 24.0;

                        F(s->memory_dsigmaxx_dx, i, j, k) = P1(b_x, i) * F(s->memory_dsigmaxx_dx, i, j, k) + P1(a_x, i) * value_dsigmaxx_dx;
                        F(s->memory_dsigmaxy_dy, i, j, k) = P1(b_y, j) * F(s->memory_dsigmaxy_dy, i, j, k) + P1(a_y, j) * value_dsigmaxy_dy;
                        F(s->memory_dsigmaxz_dz, i, j, k) = P1(b_z, kglobal) * F(s->memory_dsigmaxz_dz, i, j, k) + P1(a_z, kglobal) * value_dsigmaxz_dz;

                        value_dsigmaxx_dx = value_dsigmaxx_dx / P1(K_x, i) + F(s->memory_dsigmaxx_dx, i, j, k);
                        value_dsigmaxy_dy = value_dsigmaxy_dy / P1(K_y, j) + F(s->memory_dsigmaxy_dy, i, j, k);
                        value_dsigmaxz_dz = value_dsigmaxz_dz / P1(K_z, kglobal) + F(s->memory_dsigmaxz_dz, i, j, k);

                        F(s->vx, i, j, k) = DELTAT_over_rho * (value_dsigmaxx_dx + value_dsigmaxy_dy + value_dsigmaxz_dz) + F(s->vx, i, j, k);
                    }
                }
                for (int j = 1; j <= NY - 1; j++) {
                    for (int i = 1; i <= NX - 1; i++) {
                        double value_dsigmaxy_dx = (27.0 * F(s->sigmaxy, i + 1, j, k) - 27.0 * F(s->sigmaxy, i, j, k) - F(s->sigmaxy, i + 2, j, k) + F(s->sigmaxy, i - 1, j, k)) * ONE_OVER_DELTAX / 24.0;
                        double value_dsigmayy_dy = (27.0 * F(s->sigmayy, i, j + 1, k) - 27.0 * F(s->sigmayy, i, j, k) - F(s->sigmayy, i, j + 2, k) + F(s->sigmayy, i, j - 1, k)) * ONE_OVER_DELTAY / 24.0;
                        double value_dsigmayz_dz = (27.0 * F(s->sigmayz, i, j, k) - 27.0 * F(s->sigmayz, i, j, k - 1) - F(s->sigmayz, i, j, k + 1) + F(s->sigmayz, i, j, k - 2)) * ONE_OVER_DELTAZ / 24.0;

                        F(s->memory_dsigmaxy_dx, i, j, k) = P1(b_x_half, i) * F(s->memory_dsigmaxy_dx, i, j, k) + P1(a_x_half, i) * value_dsigmaxy_dx;
                        F(s->memory_dsigmayy_dy, i, j, k) = P1(b_y_half, j) * F(s->memory_dsigmayy_dy, i, j, k) + P1(a_y_half, j) * value_dsigmayy_dy;
                        F(s->memory_dsigmayz_dz, i, j, k) = P1(b_z, kglobal) * F(s->memory_dsigmayz_dz, i, j, k) + P1(a_z, kglobal) * value_dsigmayz_dz;

                        value_dsigmaxy_dx = value_dsigmaxy_dx / P1(K_x_half, i) + F(s->memory_dsigmaxy_dx, i, j, k);
                        value_dsigmayy_dy = value_dsigmayy_dy / P1(K_y_half, j) + F(s->memory_dsigmayy_dy, i, j, k);
                        value_dsigmayz_dz = value_dsigmayz_dz / P1(K_z, kglobal) + F(s->memory_dsigmayz_dz, i, j, k);

                        F(s->vy, i, j, k) = DELTAT_over_rho * (value_dsigmaxy_dx + value_dsigmayy_dy + value_dsigmayz_dz) + F(s->vy, i, j, k);
                    }
                }
            }

            /* ---- vz : :1287-1308 */
#pragma omp parallel for schedule(static)
            for (int k = 1; k <= kminus1end; k++) {
                int kglobal = k + offset_k;
                for (int j = 2; j <= NY; j++) {
                    for (int i = 1; i <= NX - 1; i++) {
                        double value_dsigmaxz_dx = (27.0 * F(s->sigmaxz, i + 1, j, k) - 27.0 * F(s->sigmaxz, i, j, k) - F(s->sigmaxz, i + 2, j, k) + F(s->sigmaxz, i - 1, j, k)) * ONE_OVER_DELTAX / 24.0;
                        double value_dsigmayz_dy = (27.0 * F(s->sigmayz, i, j, k) - 27.0 * F(s->sigmayz, i, j - 1, k) - F(s->sigmayz, i, j + 1, k) + F(s->sigmayz, i, j - 2, k)) * ONE_OVER_DELTAY / 24.0;
                        double value_dsigmazz_dz = (27.0 * F(s->sigmazz, i, j, k + 1) - 27.0 * F(s->sigmazz, i, j, k) - F(s->sigmazz, i, j, k + 2) + F(s->sigmazz, i, j, k - 1)) * ONE_OVER_DELTAZ / 24.0;

                        F(s->memory_dsigmaxz_dx, i, j, k) = P1(b_x_half, i) * F(s->memory_dsigmaxz_dx, i, j, k) + P1(a_x_half, i) * value_dsigmaxz_dx;
                        F(s->memory_dsigmayz_dy, i, j, k) = P1(b_y, j) * F(s->memory_dsigmayz_dy, i, j, k) + P1(a_y, j) * value_dsigmayz_dy;
                        F(s->memory_dsigmazz_dz, i, j, k) = P1(b_z_half, kglobal) * F(s->memory_dsigmazz_dz, i, j, k) + P1(a_z_half, kglobal) * value_dsigmazz_dz;

                        value_dsigmaxz_dx = value_dsigmaxz_dx / P1(K_x_half, i) + F(s->memory_dsigmaxz_dx, i, j, k);
                        value_dsigmayz_dy = value_dsigmayz_dy / P1(K_y, j) + F(s->memory_dsigmayz_dy, i, j, k);
                        value_dsigmazz_dz = value_dsigmazz_dz / P1(K_z_half, kglobal) + F(s->memory_dsigmazz_dz, i, j, k);

                        F(s->vz, i, j, k) = DELTAT_over_rho * (value_dsigmaxz_dx + value_dsigmayz_dy + value_dsigmazz_dz) + F(s->vz, i, j, k);
                    }
                }
            }

            /* ---- source : :1310-1335 */
            if (r == src_rank) {
                int i = cfg->isource, j = cfg->jsource;
                F(s->vx, i, j, src_klocal) = F(s->vx, i, j, src_klocal) + force_x[it - 1] * DELTAT / rho;
                F(s->vy, i, j, src_klocal) = F(s->vy, i, j, src_klocal) + force_y[it - 1] * DELTAT / rho;
            }

            /* ---- Dirichlet, two planes per face : :1337-1371 ; the (:,:) sections span k = -1..NZ_LOCAL+2 */
#pragma omp parallel for schedule(static)
            for (int k = -1; k <= NZ_LOCAL + 2; k++) {
                for (int j = 0; j <= NY + 1; j++) {
                    for (int i = 0; i <= 1; i++) { F(s->vx, i, j, k) = 0.0; F(s->vy, i, j, k) = 0.0; F(s->vz, i, j, k) = 0.0; }
                    for (int i = NX; i <= NX + 1; i++) { F(s->vx, i, j, k) = 0.0; F(s->vy, i, j, k) = 0.0; F(s->vz, i, j, k) = 0.0; }
                }
                for (int i = 0; i <= NX + 1; i++) {
                    for (int j = 0; j <= 1; j++) { F(s->vx, i, j, k) = 0.0; F(s->vy, i, j, k) = 0.0; F(s->vz, i, j, k) = 0.0; }
                    for (int j = NY; j <= NY + 1; j++) { F(s->vx, i, j, k) = 0.0; F(s->vy, i, j, k) = 0.0; F(s->vz, i, j, k) = 0.0; }
                }
            }
            if (r == 0)
                for (int k = 0; k <= 1; k++)
                    for (size_t q = 0; q < PLANE; q++) {
                        (&F(s->vx, 0, 0, k))[q] = 0.0; (&F(s->vy, 0, 0, k))[q] = 0.0; (&F(s->vz, 0, 0, k))[q] = 0.0;
                    }
            if (r == NPROC - 1)
                for (int k = NZ_LOCAL; k <= NZ_LOCAL + 1; k++)
                    for (size_t q = 0; q < PLANE; q++) {
                        (&F(s->vx, 0, 0, k))[q] = 0.0; (&F(s->vy, 0, 0, k))[q] = 0.0; (&F(s->vz, 0, 0, k))[q] = 0.0;
                    }

            /* ---- seismograms : :1373-1379 */
            if (r == src_rank)
                for (int irec = 1; irec <= NREC; irec++) {
                    sisvx[(size_t)(it - 1) + (size_t)NSTEP * (irec - 1)] = F(s->vx, ix_rec[irec - 1], iy_rec[irec - 1], src_klocal);
                    sisvy[(size_t)(it - 1) + (size_t)NSTEP * (irec - 1)] = F(s->vy, ix_rec[irec - 1], iy_rec[irec - 1], src_klocal);
                }

            /* ---- energy : :1381-1430 */
            {
                double local_energy_kinetic = 0.0, local_energy_potential = 0.0;
                int kmin = 1, kmax = NZ_LOCAL;
                if (r == 0) kmin = NPOINTS_PML;
                if (r == NPROC - 1) kmax = NZ_LOCAL - NPOINTS_PML + 1;
#pragma omp parallel for schedule(static) reduction(+ : local_energy_kinetic, local_energy_potential)
                for (int k = kmin; k <= kmax; k++) {
                    for (int j = NPOINTS_PML; j <= NY - NPOINTS_PML + 1; j++) {
                        for (int i = NPOINTS_PML; i <= NX - NPOINTS_PML + 1; i++) {
                            double vxv = F(s->vx, i, j, k), vyv = F(s->vy, i, j, k), vzv = F(s->vz, i, j, k);
                            local_energy_kinetic = local_energy_kinetic + 0.5 * rho * (vxv * vxv + vyv * vyv + vzv * vzv);

                            double sxx = F(s->sigmaxx, i, j, k), syy = F(s->sigmayy, i, j, k), szz = F(s->sigmazz, i, j, k);
                            double epsilon_xx = (2.0 * (lambda + mu) * sxx - lambda * syy - lambda * szz) / (2.0 * mu * (3.0 * lambda + 2.0 * mu));
                            double epsilon_yy = (2.0 * (lambda + mu) * syy - lambda * sxx - lambda * szz) / (2.0 * mu * (3.0 * lambda + 2.0 * mu));
                            /* epsilon_zz is computed (:1408-1409) but never used (quirk B2) */
                            double epsilon_xy = F(s->sigmaxy_R, i, j, k) / (2.0 * mu);
                            double epsilon_xz = F(s->sigmaxz_R, i, j, k) / (2.0 * mu);
                            double epsilon_yz = F(s->sigmayz_R, i, j, k) / (2.0 * mu);

                            local_energy_potential = local_energy_potential +
                                0.5 * (epsilon_xx * F(s->sigmaxx_R, i, j, k) + epsilon_yy * F(s->sigmayy_R, i, j, k) +
                                       epsilon_yy * F(s->sigmayy_R, i, j, k) + 2.0 * epsilon_xy * F(s->sigmaxy_R, i, j, k) +
                                       2.0 * epsilon_xz * F(s->sigmaxz_R, i, j, k) + 2.0 * epsilon_yz * F(s->sigmayz_R, i, j, k));
                        }
                    }
                }
                sum_total += local_energy_kinetic + local_energy_potential;       /* MPI_REDUCE(SUM) :1425-1430 */
                sum_kinetic += local_energy_kinetic;
                sum_potential += local_energy_potential;
            }
        }
        energy_total[it - 1] = sum_total;
        energy_kinetic[it - 1] = sum_kinetic;
        energy_potential[it - 1] = sum_potential;
    }
    oracle_g_loop_seconds = oracle_now_seconds() - t_loop_start;

    /* ---- results */
    if (vnorm_final) {                             /* :1435-1436 */
        double vmax = 0.0;
        for (int r = 0; r < NPROC; r++)
            for (int k = 1; k <= NZ_LOCAL; k++)
                for (size_t q = 0; q < PLANE; q++) {
                    double a = (&F(S[r].vx, 0, 0, k))[q], b = (&F(S[r].vy, 0, 0, k))[q], c = (&F(S[r].vz, 0, 0, k))[q];
                    double v = sqrt(a * a + b * b + c * c);
                    if (v > vmax) vmax = v;
                }
        *vnorm_final = vmax;
    }
    if (fields_final) {
        const size_t G = (size_t)NX * NY * NZ;
        for (int r = 0; r < NPROC; r++) {
            double **f = (double **)&S[r];
            for (int q = 0; q < 15; q++)
                for (int k = 1; k <= NZ_LOCAL; k++)
                    for (int j = 1; j <= NY; j++)
                        memcpy(fields_final + (size_t)q * G + (size_t)NX * ((size_t)(j - 1) + (size_t)NY * (size_t)(r * NZ_LOCAL + k - 1)),
                               &F(f[q], 1, j, k), (size_t)NX * sizeof(double));
        }
    }
    for (int r = 0; r < NPROC; r++) {
        double **f = (double **)&S[r];
        for (int q = 0; q < NV_ALL; q++) free(f[q]);
    }
    free(S);
#undef IDX
#undef F
#undef E
#undef P1
    return 0;
}
