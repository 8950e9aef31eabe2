/*
 * cpml_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Plain-C restatement of the SEISMIC_CPML reference time loops, used only as
 * the checker in tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.  Nothing under seismic_cpml_b200/ may include, link
 * or call this.
 *
 * PARITY PINNED BY AN EXECUTION OF THE REFERENCE SOURCE: the reference is
 * Fortran90, no Fortran compiler (nor MPI) exists in this image or on the GPU
 * box, and the reference ships no golden vectors.  oracle/f90_exec.py therefore
 * executes the main program of each reference .f90 file from its own source
 * text (statement-by-statement transliteration to Python, IEEE double, source
 * order, MPI ranks as threads); tests/golden/ref_*.npz hold what the six
 * programs computed that way on reduced grids, and this restatement reproduces
 * every one of them bit for bit (tests/test_reference_vectors.py: profiles,
 * receivers, seismograms, energies, final fields).  It is not the compiled
 * reference: a gfortran build could differ where a compiler contracts or
 * reassociates.  Further pins: (i) closed-form setup constants (SURVEY.md App.
 * C.1), (ii) an independent numpy restatement (oracle/np_restatement.py) that
 * agrees bit-for-bit, (iii) slab-count invariance of the 3-D code, (iv)
 * closed-form physical solutions.  Every function cites the reference
 * file:line it follows.
 *
 * All arrays are Fortran-ordered (i fastest) so that they can be compared with
 * the reference's arrays index by index.  All indices crossing this API are
 * 1-based like the reference's.
 */
#ifndef CPML_ORACLE_H
#define CPML_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ setup */

/* One axis of C-PML damping profiles, integer and half grid points.
 * Follows seismic_CPML_3D_isotropic_MPI_OpenMP.f90:399-667 (same text in
 * seismic_CPML_2D_isotropic_second_order.f90:283-465).
 *   origin_top_uses_n: 0 -> right origin = (n-1)*delta - L   (all programs)
 *                      1 -> right origin =  n   *delta - L   (quirk B4,
 *                           seismic_CPML_2D_isotropic_fourth_order.f90:401, y only)
 *   clamp_alpha:       1 for the x axis (3D-iso :514-515), 0 for y, z.
 * Output arrays have length n (index 0 <-> Fortran index 1). */
void oracle_pml_profile(int n, double delta, double deltat, int npoints_pml,
                        int use_pml_min, int use_pml_max,
                        double cp, double rcoef, double npower,
                        double k_max_pml, double alpha_max_pml,
                        int origin_top_uses_n, int clamp_alpha,
                        double *a, double *b, double *K,
                        double *a_half, double *b_half, double *K_half);

/* Same profiles with the viscoelastic program's d0 = -(NPOWER+1)*cp*dsqrt(taumax)*log(Rcoef)/(2L)
 * (seismic_CPML_3D_viscoelastic_MPI.f90:547-549; profiles :553-801, identical text otherwise;
 * Rcoef = 1e-4 :541, K_MAX_PML = 7 :243).  sqrt_taumax = 1 is oracle_pml_profile. */
void oracle_pml_profile_visco(int n, double delta, double deltat, int npoints_pml,
                              int use_pml_min, int use_pml_max,
                              double cp, double sqrt_taumax, double rcoef, double npower,
                              double k_max_pml, double alpha_max_pml,
                              int origin_top_uses_n, int clamp_alpha,
                              double *a, double *b, double *K,
                              double *a_half, double *b_half, double *K_half);

/* Source time function, first derivative of a Gaussian.
 * Follows 3D-iso :1058-1071 / 2D-2nd :644-657.  force_x/force_y have length nstep
 * (entry it-1 holds the value used at time step it). */
void oracle_source_series(int nstep, double deltat, double f0, double t0,
                          double factor, double angle_force_deg,
                          double *force_x, double *force_y);

/* Nearest-grid-point receiver search, strict '<', j outer / i inner.
 * Follows 3D-iso :683-706 / 2D-2nd :486-509.  ix_rec/iy_rec are 1-based. */
void oracle_find_receivers(int nx, int ny, double deltax, double deltay, int nrec,
                           double xdeb, double ydeb, double xfin, double yfin,
                           int *ix_rec, int *iy_rec, double *dist_rec);

/* ---------------------------------------------------------------- 2-D iso */

typedef struct {
    int order;            /* 2: seismic_CPML_2D_isotropic_second_order.f90
                             4: seismic_CPML_2D_isotropic_fourth_order.f90 */
    int nx, ny;
    double deltax, deltay, deltat;
    int nstep;
    int npoints_pml;
    int isource, jsource; /* 1-based */
    int nrec;
} oracle2d_config;

/* Runs time steps 1..nstep of the 2-D isotropic program (2D-2nd :550-713,
 * 2D-4th :551-714).  lambda/mu/rho: nx*ny, i fastest.  Profiles: length nx / ny.
 * force_x/force_y: nstep.  Outputs: sisvx/sisvy (nstep*nrec, column-major
 * (it,irec) like the reference's sisvx(NSTEP,NREC)), energy_kinetic /
 * energy_potential (nstep), and optionally the final fields (nx*ny each, may be
 * NULL).  Returns 0. */
int oracle_run_2d(const oracle2d_config *cfg,
                  const double *lambda, const double *mu, const double *rho,
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
                  double *velocnorm_final);

/* ---------------------------------------------------------------- 3-D iso */

typedef struct {
    int nx, ny, nz;       /* global grid */
    int nproc;            /* number of emulated MPI z-slabs (reference: even) */
    double deltax, deltay, deltaz, deltat;
    double lambda, mu, lambdaplustwomu, rho; /* homogeneous medium, 3D-iso :139-144
                             (lambdaplustwomu = rho*cp*cp is its own constant there) */
    int nstep;
    int npoints_pml;
    int isource, jsource; /* 1-based; k of source = nz/2 (3D-iso :346,1080) */
    int nrec;
    int energy_bug_compat; /* 1 = reference formula (yy twice, no zz; quirk B2) */
} oracle3d_config;

/* Runs time steps 1..nstep of seismic_CPML_3D_isotropic_MPI_OpenMP.f90:802-1180
 * with the MPI ranks emulated as nproc slabs in one address space (plane
 * exchange = memcpy, MPI_REDUCE = sum in rank order).  Full-grid memory
 * variables and separate Dirichlet / energy passes exactly as the reference.
 * Profiles have length nx / ny / nz (global).  Outputs: sisvx/sisvy
 * (nstep*nrec), total_energy (nstep); optional (may be NULL): plane_vx /
 * plane_vy = vx,vy(:,:,nz/2) at the end (nx*ny), fields_final = the 9 global
 * fields vx,vy,vz,sxx,syy,szz,sxy,sxz,syz each nx*ny*nz (k = 1..nz) back to
 * back, vnorm_final = max |v|.  Returns 0, or nonzero for a bad topology
 * (3D-iso :381-394). */
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
                      double *fields_final, double *vnorm_final);

/* ----------------------------------------------------------- 3-D viscoelastic */

/* Nearest-grid-point search of 3D-visco :839-853: like oracle_find_receivers but the
 * abscissa of grid point i is DELTAX*i (not i-1) and the targets are given explicitly
 * (:832-837: (xs+500, ys+500), (xs, ys+2260), (xs+500, ys+2260)). */
void oracle_find_receivers_visco(int nx, int ny, double deltax, double deltay, int nrec,
                                 const double *xrec, const double *yrec,
                                 int *ix_rec, int *iy_rec, double *dist_rec);

typedef struct {
    int nx, ny, nz;       /* global grid */
    int nproc;            /* emulated MPI z-slabs (reference default 4, :158) */
    double deltax, deltay, deltaz, deltat;
    double lambda, mu, rho;   /* relaxed Lame parameters and density, :167-172 */
    int nstep;
    int npoints_pml;
    int isource, jsource; /* 1-based; k of source = nz/2 (:479, :1332) */
    int nrec;
    /* N_SLS = 2 relaxation mechanisms (:189); nu1 = dilatation (QKappa), nu2 = shear (QMu) */
    double tau_epsilon_nu1[2], tau_sigma_nu1[2], tau_epsilon_nu2[2], tau_sigma_nu2[2];
    int complete_halos;   /* 0 = the reference's exchange (quirk B6: half of the 4th-order z halo is
                             never received, so the result depends on nproc); 1 = every plane the
                             stencils read is exchanged (nproc-independent; NOT the reference) */
    int sigmazz_isotropic; /* 0 = the reference's memory-variable term of sigmazz (:1058-1060: (lambda+2mu) sum e1
                              - 2/3 mu sum(e11+e22), quirk B14); 1 = the isotropic form sigmaxx / sigmayy use,
                              (lambda+2/3 mu) sum e1 - 2 mu sum(e11+e22) (NOT the reference; analytical check only) */
} oraclev3d_config;

/* Runs time steps 1..nstep of seismic_CPML_3D_viscoelastic_MPI.f90:954-1430 with the MPI ranks
 * emulated as nproc slabs (arrays (0:NX+1,0:NY+1,-1:NZ_LOCAL+2) per slab, :301-303).  Outputs:
 * sisvx/sisvy (nstep*nrec), energy_total/kinetic/potential (nstep, :1425-1430); optional
 * fields_final = the 9 global fields vx..sigmayz then the 6 sigma*_R (xx,yy,zz,xy,xz,yz), each
 * nx*ny*nz (i = 1..nx, j = 1..ny, k = 1..nz) back to back; vnorm_final = max |v| (:1435). */
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
                        double *fields_final, double *vnorm_final);

/* ----------------------------------------------------------- 2-D viscoelastic */

typedef struct {
    int order;            /* 2: seismic_CPML_2D_velocity_and_stress_second_order_viscoelastic.f90
                             4: seismic_CPML_2D_velocity_and_stress_fourth_order_viscoelastic.f90 */
    int nx, ny;
    double deltax, deltay, deltat;
    int nstep;
    int npoints_pml;
    int isource, jsource; /* 1-based */
    int nrec;
    int viscoelastic_attenuation;   /* VISCOELASTIC_ATTENUATION (:140); 0 runs the elastic branch :713-760 */
    int compute_energy;             /* COMPUTE_ENERGY (:201), .false. in the reference */
    /* N_SLS = 3 Zener solids (:329); nu1 from Qp, nu2 from Qs */
    double tau_epsilon_nu1[3], tau_sigma_nu1[3], tau_epsilon_nu2[3], tau_sigma_nu2[3];
} oraclev2d_config;

/* Ricker source of the 2-D viscoelastic programs divided by the cell area (2D-visco-4th :931-947):
 * force_source_term = factor*(1 - 2a(t-t0)^2) exp(-a(t-t0)^2) / (DELTAX*DELTAY). */
void oracle_source_series_ricker(int nstep, double deltat, double f0, double t0, double factor,
                                 double angle_force_deg, double deltax, double deltay,
                                 double *force_x, double *force_y);

/* Runs time steps 1..nstep of the 2-D viscoelastic programs (2D-visco-4th :705-1060; the
 * second-order file differs only in the difference operator).  lambda/mu are the UNRELAXED
 * parameters (:596-601), arrays of nx*ny values, i fastest.  Outputs: sisvx/sisvy/sispressure
 * (nstep*nrec), energies (nstep; zero unless compute_energy), optional final fields vx, vy,
 * sigma_xx, sigma_yy, sigma_xy (nx*ny each) and the nine memory variables e1(1..3), e11(1..3),
 * e13(1..3) back to back (9*nx*ny). */
int oracle_run_2d_visco(const oraclev2d_config *cfg,
                        const double *lambda_unrelaxed, const double *mu_unrelaxed, const double *rho,
                        const double *a_x, const double *b_x, const double *K_x,
                        const double *a_x_half, const double *b_x_half, const double *K_x_half,
                        const double *a_y, const double *b_y, const double *K_y,
                        const double *a_y_half, const double *b_y_half, const double *K_y_half,
                        const double *force_x, const double *force_y,
                        const int *ix_rec, const int *iy_rec,
                        double *sisvx, double *sisvy, double *sispressure,
                        double *energy_kinetic, double *energy_potential,
                        double *fields_final, double *memvar_final, double *velocnorm_final);

/* Timing of the last oracle_run_* call: wall seconds of its time loop only (set-up and
 * allocation excluded), not counting the first `w` warm-up steps. */
void oracle_set_warmup_steps(int w);
double oracle_last_loop_seconds(void);

/* Number of OpenMP threads the timed build will use (1 if built without). */
int oracle_num_threads(void);
/* Overrides OMP_NUM_THREADS for the following oracle_run_* calls (timed build). */
void oracle_set_num_threads(int n);
/* Flush-to-zero / denormals-are-zero for the timed CPU baseline (cf. the
 * reference Makefile:18 remark on -ftz).  No-op in the golden build. */
void oracle_set_ftz(int on);

#ifdef __cplusplus
}
#endif
#endif
