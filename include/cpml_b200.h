/*
 * cpml_b200.h -- C ABI of libcpml_b200.so, the B200 (sm_100a) implementation of the
 * SEISMIC_CPML staggered-grid velocity-stress C-PML time loop.
 *
 * The reference (geodynamics/seismic_cpml) has no library/FFI boundary: every solver
 * is one Fortran `program` whose hot loop `do it = 1,NSTEP` is inline.  This header
 * IS the boundary a maintainer cuts at that line:
 *     seismic_CPML_3D_isotropic_MPI_OpenMP.f90:802      (3-D isotropic)
 *     seismic_CPML_2D_isotropic_second_order.f90:550    (2-D, 2nd order)
 *     seismic_CPML_2D_isotropic_fourth_order.f90:551    (2-D, 4th order)
 *     seismic_CPML_3D_viscoelastic_MPI.f90:954          (3-D viscoelastic, 4th order)
 * Everything above that line (parameters, C-PML profiles, source law, receiver
 * search, CFL check) stays in the driver and is handed over through the setters;
 * everything the loop body touches lives on the GPU behind an opaque handle; the
 * output phase pulls seismograms / energy / snapshot planes back through getters.
 * The Fortran binding (module cpml_b200, ISO_C_BINDING) is in drivers/fortran/ and
 * INTEGRATION.md.
 *
 * Conventions: extern "C", plain pointers and sizes, int32/double only, every
 * function returns 0 on success or a CPML_E* code (message via cpml_last_error);
 * no exceptions, no global state; the caller owns every host buffer; the library
 * owns all device memory.  All array arguments are Fortran-ordered (i fastest) and
 * all grid indices are 1-based, exactly like the reference's arrays, so a Fortran
 * driver passes its arrays as they are.  One host thread per handle.
 * There is NO CPU fallback: without a CUDA device cpml_create fails.
 */
#ifndef CPML_B200_H
#define CPML_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CPML_B200_ABI_VERSION 1

/* error codes */
#define CPML_OK            0
#define CPML_EINVAL        1   /* bad argument / inconsistent configuration            */
#define CPML_ETOPOLOGY     2   /* slab checks of 3D-iso :381-394 failed                */
#define CPML_ECFL          3   /* Courant number > 1 (3D-iso :717, 2D-2nd :516)        */
#define CPML_ECUDA         4   /* CUDA runtime error / no usable device                */
#define CPML_ESTATE        5   /* call order (e.g. run before profiles were set)       */
#define CPML_ENOMEM        6

/* axes for cpml_set_profiles */
#define CPML_AXIS_X 0
#define CPML_AXIS_Y 1
#define CPML_AXIS_Z 2

/* field ids for cpml_get_plane / cpml_get_field / cpml_halo_plane */
#define CPML_F_VX       0
#define CPML_F_VY       1
#define CPML_F_VZ       2
#define CPML_F_SIGMAXX  3
#define CPML_F_SIGMAYY  4
#define CPML_F_SIGMAZZ  5
#define CPML_F_SIGMAXY  6
#define CPML_F_SIGMAXZ  7
#define CPML_F_SIGMAYZ  8
/* viscoelastic only: the relaxed stresses sigma*_R of 3D-visco :302 (energy only) */
#define CPML_F_SIGMAXX_R 9
#define CPML_F_SIGMAYY_R 10
#define CPML_F_SIGMAZZ_R 11
#define CPML_F_SIGMAXY_R 12
#define CPML_F_SIGMAXZ_R 13
#define CPML_F_SIGMAYZ_R 14

typedef struct cpml_handle cpml_handle;

/* Mirrors the compile-time `parameter` block of the reference programs
 * (3D-iso :124-218, 2D-2nd :138-218); names follow the Fortran names.
 * bind(C)-friendly: int32 and double only. */
typedef struct cpml_config {
    int32_t ndim;            /* 2 or 3                                                   */
    int32_t order;           /* spatial order: 2 or 4 (ndim==2); ndim==3: 2 (isotropic) or
                                4 (viscoelastic)                                            */
    int32_t nx, ny, nz;      /* NX, NY, NZ: GLOBAL grid; nz ignored when ndim==2         */
    int32_t nstep;           /* NSTEP: capacity of the seismogram / energy traces        */
    int32_t npoints_pml;     /* NPOINTS_PML: only used for the energy box (:1135-1145)   */
    int32_t nrec;            /* NREC                                                     */
    int32_t isource, jsource;/* ISOURCE, JSOURCE (1-based)                               */
    int32_t ksource;         /* 3-D: global k of source AND receiver plane; 0 => NZ/2,
                                the reference's cut plane (:346,1080,1126)               */
    int32_t nslabs;          /* number of z-slabs = reference NPROC (1 = whole grid)     */
    int32_t slab_rank;       /* which slab this handle owns, 0-based = MPI rank          */
    int32_t device;          /* CUDA device ordinal, -1 = current device                 */
    int32_t energy_bug_compat; /* 1 = reference 3-D potential energy (yy counted twice,
                                zz omitted, :1169-1172); 0 = physical formula            */
    int32_t rheology;        /* 0 = isotropic elastic; 1 = viscoelastic (needs cpml_set_attenuation):
                                ndim == 3, order == 4: seismic_CPML_3D_viscoelastic_MPI.f90, N_SLS = 2;
                                ndim == 2, order 2 | 4: seismic_CPML_2D_velocity_and_stress_
                                {second,fourth}_order_viscoelastic.f90, N_SLS = 3              */
    int32_t emulate_nproc;   /* viscoelastic only: the reference's MPI exchange delivers only
                                half of the z halo its 4th-order stencils read (3D-visco
                                :962-975,:1229-1242 vs :991,:1149,:1189,:1251,:1271,:1294), so
                                its result depends on NPROC.  n > 1 reproduces the reference run
                                with NPROC = n (default 4, :158) on ANY number of GPU slabs;
                                0 or 1 = every tap delivered (single-rank semantics)         */
    int32_t compute_energy;  /* 2-D viscoelastic only: COMPUTE_ENERGY of 2D-visco-4th :201 (.false. in the
                                reference, the energy traces then stay zero); the other solvers always
                                compute the energy, as their programs do                      */
    int32_t sigmazz_isotropic; /* 3-D viscoelastic only.  0 = the reference: the memory-variable term of sigmazz is
                                (lambda+2mu) sum e1 - 2/3 mu sum(e11+e22) (3D-visco :1058-1060), which is not the
                                isotropic form its sigmaxx / sigmayy use (6 % misfit to the
                                analytical viscoelastic solution, tests/test_analytical_visco3d.py).  1 = the
                                isotropic form (lambda+2/3 mu) sum e1 - 2 mu sum(e11+e22): NOT the reference,
                                for users who want the physics (1.6 % misfit)                              */
    int32_t precision;       /* 0 = double precision, the reference as shipped.  1 = single precision, the build the
                                reference endorses as "significantly faster" (3D-iso :114-116: declare everything
                                `real`): wavefields, memory variables, profiles and update constants in FP32 (the
                                energy is still summed in double); 3-D isotropic solver on one GPU only.  Getters keep
                                returning double arrays.  Not bit-comparable with the double-precision run: see
                                DESIGN.md for the measured differences                                          */
    double deltax, deltay, deltaz;   /* DELTAX, DELTAY, DELTAZ                           */
    double deltat;                   /* DELTAT                                           */
    /* homogeneous medium of the 3-D program (:139-144); the 2-D programs take arrays
       through cpml_set_material_2d and ignore these four */
    double lambda, mu, lambdaplustwomu, rho;
    double cp;                       /* P velocity, only for the Courant check           */
    double reserved_d[4];
} cpml_config;

/* ---- life cycle ---------------------------------------------------------- */

int32_t cpml_abi_version(void);

/* Validates (topology checks of 3D-iso :381-394, CFL check :712-717), allocates and
 * zeroes the device state (:720-756).  *out is NULL on failure; the reason is then
 * available from cpml_last_error(NULL). */
int32_t cpml_create(const cpml_config *cfg, cpml_handle **out);
int32_t cpml_destroy(cpml_handle *h);
const char *cpml_last_error(const cpml_handle *h);

/* Re-zero fields, memory variables, seismograms and energy (== :720-756). */
int32_t cpml_reset(cpml_handle *h);

/* Run the kernels on this CUDA stream (cudaStream_t passed as void*); default 0. */
int32_t cpml_set_stream(cpml_handle *h, void *cuda_stream);

/* ---- inputs (driver -> device), replacing the arrays built before the loop -- */

/* 1-D damping profiles of one axis, built by the driver at :425-667.  n must be
 * NX, NY or the GLOBAL NZ.  The library stores C-PML memory variables only where a
 * profile is non-trivial (a != 0 or K != 1); elsewhere they are identically zero
 * in the reference too, so this is exact. */
int32_t cpml_set_profiles(cpml_handle *h, int32_t axis,
                          const double *a, const double *b, const double *K,
                          const double *a_half, const double *b_half, const double *K_half,
                          int32_t n);

/* 2-D only: lambda, mu, rho as filled at 2D-2nd :468-474, NX*NY values each. */
int32_t cpml_set_material_2d(cpml_handle *h, const double *lambda, const double *mu,
                             const double *rho);

/* Viscoelastic only: the relaxation times of the standard linear solids, as returned by
 * compute_attenuation_coeffs.
 * 3-D (n_sls = 2, 3D-visco :439-443; nu1 = dilatation / QKappa, nu2 = shear / QMu): the library
 * derives inv_tau_sigma, phi_nu, Mu_nu and the unrelaxed Lame parameters with the reference's own
 * expressions (:458-477, :982-987).  lambda, mu of cpml_config are the RELAXED parameters
 * (:171-172); lambdaplustwomu is ignored (the loop uses lambda + 2 mu, :984).
 * 2-D (n_sls = 3, 2D-visco-4th :366-370; nu1 from Qp, nu2 from Qs): the library derives
 * HALF_DELTAT_over_tau_sigma, multiplication_factor_tau_sigma and DELTAT_phi (:386-399); the arrays
 * of cpml_set_material_2d are then the UNRELAXED lambda, mu (:596-601).  tau_epsilon == tau_sigma
 * gives the elastic branch (VISCOELASTIC_ATTENUATION = .false., :713-760) bit for bit. */
int32_t cpml_set_attenuation(cpml_handle *h, int32_t n_sls,
                             const double *tau_epsilon_nu1, const double *tau_sigma_nu1,
                             const double *tau_epsilon_nu2, const double *tau_sigma_nu2);

/* force_x(it), force_y(it) of :1058-1071 for it = 1..n (n <= NSTEP).  The kernel
 * adds force*DELTAT/rho like :1080-1081. */
int32_t cpml_set_source_series(cpml_handle *h, const double *force_x, const double *force_y,
                               int32_t n);

/* Per-step forms for drivers that keep the reference's loop structure: the source term of
 * step `it` (evaluated inside `do it`, :1058-1071) goes to the device through pinned staging,
 * and the per-step outputs (kinetic and potential energy of this slab, :1179, and the sample
 * of receiver 1, :1126-1127) come back the same way.  All copies are ordered on the handle's
 * stream and do not synchronise; cpml_get_fetched_step is valid after cpml_synchronize. */
int32_t cpml_set_source_step(cpml_handle *h, int32_t it, double force_x, double force_y);
int32_t cpml_fetch_step(cpml_handle *h, int32_t it);
int32_t cpml_get_fetched_step(cpml_handle *h, int32_t it, double *out4);

/* ix_rec, iy_rec of :691-706 (1-based), n == NREC. */
int32_t cpml_set_receivers(cpml_handle *h, const int32_t *ix_rec, const int32_t *iy_rec,
                           int32_t n);

/* ---- the loop body -------------------------------------------------------- */

/* Executes time steps it_begin..it_end (inclusive, 1-based) == the body of
 * `do it` (:804-1180): stress update, velocity update, source, Dirichlet, seismogram
 * sample, energy.  Single-slab handles only (nslabs == 1, or after cpml_attach_*).
 * Returns after the stream has drained. */
int32_t cpml_run(cpml_handle *h, int32_t it_begin, int32_t it_end);

/* The same body in the pieces a slab-decomposed driver interleaves with its plane
 * exchange (the MPI_SENDRECV calls at :811-823 and :951-963).  Asynchronous: they
 * only enqueue work on the handle's stream.
 *   cpml_step_stress   == :825-944
 *   cpml_step_velocity == :965-1129 (+ per-block energy partials of :1131-1177)
 *   cpml_step_finish   == finishes step `it`: energy sum of this slab, seismogram sample */
int32_t cpml_step_stress(cpml_handle *h, int32_t it);
int32_t cpml_step_velocity(cpml_handle *h, int32_t it);
int32_t cpml_step_finish(cpml_handle *h, int32_t it);
int32_t cpml_synchronize(cpml_handle *h);

/* Device address and size of one z-plane of a field in the library's internal
 * (padded) layout, klocal = 0..NZ_LOCAL+1 (0 and NZ_LOCAL+1 are the halo planes,
 * 3D-iso :273); viscoelastic: klocal = -1..NZ_LOCAL+2 (two halo planes per side, 3D-visco :301).  Identical layout on every slab of the same grid, so a plane can be
 * moved slab-to-slab with ncclSend/ncclRecv, cudaMemcpyPeer or CUDA-aware MPI. */
int32_t cpml_halo_plane(cpml_handle *h, int32_t field, int32_t klocal,
                        void **device_ptr, int64_t *nbytes);

/* Copies plane klocal_src of `field` in slab `src` into plane klocal_dst of the same
 * field in slab `dst` (device to device, cudaMemcpyPeerAsync on dst's stream after
 * src's stream has drained).  This is one MPI_SENDRECV of :811-823 / :951-963 for a
 * single-process driver that owns several slab handles (on one or several GPUs). */
int32_t cpml_copy_plane(cpml_handle *dst, int32_t klocal_dst, cpml_handle *src, int32_t klocal_src,
                        int32_t field);

/* Direct slab-to-slab stores over NVLink: once a neighbour is attached, cpml_step_stress /
 * cpml_step_velocity write the boundary planes that neighbour needs (the six planes of the
 * MPI_SENDRECV calls at 3D-iso :811-823 and :951-963) straight into its halo planes from
 * inside the update kernels, and order the steps of the two slabs with device-side flags
 * (no host synchronisation, no separate copy).  side 0 = slab rank-1, side 1 = slab rank+1.
 * The end slabs have no outer neighbour (MPI_PROC_NULL, :775-790).
 *   one process per GPU : cpml_p2p_export on every rank, exchange the 64-byte blobs through
 *                         the launcher (MPI_Sendrecv / torch.distributed / a file), then
 *                         cpml_p2p_attach_ipc (CUDA IPC);
 *   one process, N GPUs : cpml_p2p_attach_local with the neighbour's handle.
 * Every slab must be attached on both sides it has a neighbour on before the first step, and
 * all slabs must call cpml_reset / start at the same `it` together. */
int32_t cpml_p2p_export(cpml_handle *h, void *blob, int64_t blob_capacity, int64_t *nbytes);
int32_t cpml_p2p_attach_ipc(cpml_handle *h, int32_t side, const void *blob, int64_t nbytes);
int32_t cpml_p2p_attach_local(cpml_handle *h, int32_t side, cpml_handle *neighbour);
int32_t cpml_p2p_detach(cpml_handle *h);

/* Launch geometry chosen for the 3-D kernels (diagnostics for bench.py / profiles):
 * info[0..9] = uses_tma, tile_x, tile_y, stages, planes per work item, z chunks, work items (velocity
 * kernel), CTAs of the stress kernel, CTAs of the velocity kernel, attached sides bit mask;
 * info[10..13] = tile_y, planes per work item, z chunks, work items of the stress kernel. */
int32_t cpml_get_launch_info(cpml_handle *h, int32_t *info, int32_t n);

/* ---- outputs (device -> driver) ------------------------------------------- */

/* sisvx, sisvy as declared at :286: (NSTEP,NREC) column-major, zero beyond the
 * last executed step.  On a slab that does not own the receiver plane they are zero. */
int32_t cpml_get_seismograms(cpml_handle *h, double *sisvx, double *sisvy);

/* EXTENSION (not in the reference): the 3-D programs record Vx and Vy only although
 * their plot script reads Vz_file_NNN.dat; sisvz(it, irec) = vz(ix_rec, iy_rec, NZ/2), same layout as
 * cpml_get_seismograms.  3-D only. */
int32_t cpml_get_seismograms_vz(cpml_handle *h, double *sisvz);

/* 2-D viscoelastic only: sispressure(NSTEP,NREC) of 2D-visco-4th :304, filled at :1004-1035
 * (pressure = -(lambda + 2/3 mu)(epsilon_xx + epsilon_yy) at the receiver). */
int32_t cpml_get_pressure_seismograms(cpml_handle *h, double *sispressure);

/* Energy traces, length NSTEP.  3-D: total = this slab's share of total_energy(it)
 * (:1179 sums the shares); kinetic/potential may be NULL.  2-D: kinetic and
 * potential are total_energy_kinetic/potential of 2D-2nd :211. */
int32_t cpml_get_energy(cpml_handle *h, double *total, double *kinetic, double *potential);

/* One (NX,NY) plane of a field at GLOBAL k (3-D) -- e.g. vx(:,:,NZ_LOCAL) handed to
 * create_color_image at :1236 -- or the whole field (ndim==2, kglobal ignored).
 * Returns CPML_EINVAL if this slab does not hold that plane. */
int32_t cpml_get_plane(cpml_handle *h, int32_t field, int32_t kglobal, double *out);

/* The same plane without stalling the time loop: _begin copies it device-side (so later steps cannot change it) and
 * starts the transfer into pinned host memory on a side stream, then returns; _end waits for that transfer and
 * copies the dense (NX,NY) plane to `out` (may be NULL) and/or hands out the pinned buffer itself (`pinned`, may be
 * NULL; valid until the next _begin on that slot).  Two slots (0, 1): the reference displays Vx and Vy. */
int32_t cpml_snapshot_begin(cpml_handle *h, int32_t slot, int32_t field, int32_t kglobal);
int32_t cpml_snapshot_end(cpml_handle *h, int32_t slot, double *out, const double **pinned);

/* Whole field of this slab, (NX,NY,NZ_LOCAL) (3-D) or (NX,NY) (2-D), dense. */
int32_t cpml_get_field(cpml_handle *h, int32_t field, double *out);

/* maxval(sqrt(vx**2+vy**2+vz**2)) over this slab (:1185); the driver keeps the
 * STABILITY_THRESHOLD test (:1195). */
int32_t cpml_get_maxnorm(cpml_handle *h, double *out);

/* Device time (ms, CUDA events on the handle's stream) spent in the stress and
 * velocity kernels since the last call with reset != 0; n_launches counts kernels. */
int32_t cpml_get_kernel_times(cpml_handle *h, double *ms_stress, double *ms_velocity,
                              int64_t *n_launches, int32_t reset);
int32_t cpml_enable_kernel_timing(cpml_handle *h, int32_t on);

/* Algorithmic HBM bytes one full time step of this slab must move (each field read
 * once / written once where updated, memory variables only inside the PML shells);
 * DESIGN.md states the formula.  Also split per kernel. */
int32_t cpml_algorithmic_bytes(cpml_handle *h, double *bytes_stress, double *bytes_velocity);

/* ---- the whole z-slab decomposition behind one handle (single host thread, no MPI) --- */
/* Replaces what the reference's Fortran program does around its MPI calls: rank set-up (3D-iso :337-346),
 * neighbours (:770-796), the plane exchange (:811-823, :951-963; 3D-visco :962-975, :1229-1242), the energy
 * reduction (:1179) and the seismograms of rank_cut_plane (:1124-1129).  cpml_multi_create makes `ngpus` slab
 * handles (cfg->nslabs / slab_rank / device are ignored) on devices[0..ngpus-1] (NULL: devices 0..ngpus-1; a
 * device may appear more than once), attaches neighbouring slabs to each other (cpml_p2p_attach_local) and gives
 * every device its own stream.  The setters forward to every slab; cpml_multi_step launches one pass of the loop
 * body on every slab and returns without waiting; cpml_multi_run loops and synchronises.  Getters return what the
 * reference's rank_cut_plane holds: the summed energy, the cut plane's seismograms, any plane by GLOBAL k. */
typedef struct cpml_multi cpml_multi;
int32_t cpml_multi_create(const cpml_config *cfg, int32_t ngpus, const int32_t *devices, cpml_multi **out);
int32_t cpml_multi_destroy(cpml_multi *m);
const char *cpml_multi_last_error(const cpml_multi *m);
int32_t cpml_multi_ngpus(const cpml_multi *m);
int32_t cpml_multi_slab(cpml_multi *m, int32_t rank, cpml_handle **out);     /* borrowed: timing, launch info, ... */
int32_t cpml_multi_reset(cpml_multi *m);
int32_t cpml_multi_set_profiles(cpml_multi *m, int32_t axis, const double *a, const double *b, const double *K,
                                const double *a_half, const double *b_half, const double *K_half, int32_t n);
int32_t cpml_multi_set_attenuation(cpml_multi *m, int32_t n_sls, const double *tau_epsilon_nu1, const double *tau_sigma_nu1,
                                   const double *tau_epsilon_nu2, const double *tau_sigma_nu2);
int32_t cpml_multi_set_source_series(cpml_multi *m, const double *force_x, const double *force_y, int32_t n);
int32_t cpml_multi_set_receivers(cpml_multi *m, const int32_t *ix_rec, const int32_t *iy_rec, int32_t n);
int32_t cpml_multi_step(cpml_multi *m, int32_t it);
int32_t cpml_multi_run(cpml_multi *m, int32_t it_begin, int32_t it_end);
int32_t cpml_multi_synchronize(cpml_multi *m);
int32_t cpml_multi_get_seismograms(cpml_multi *m, double *sisvx, double *sisvy);
int32_t cpml_multi_get_seismograms_vz(cpml_multi *m, double *sisvz);
int32_t cpml_multi_get_energy(cpml_multi *m, double *total, double *kinetic, double *potential);
int32_t cpml_multi_get_plane(cpml_multi *m, int32_t field, int32_t kglobal, double *out);
int32_t cpml_multi_get_field(cpml_multi *m, int32_t field, double *out);      /* (NX,NY,NZ), slabs concatenated */
int32_t cpml_multi_get_maxnorm(cpml_multi *m, double *out);

/* ---- host-side helpers that mirror the reference's set-up phase ------------- */
/* (pure host code, no device needed; the drivers in drivers/ use them)          */

/* One axis of damping profiles, 3D-iso :399-667.  origin_top_uses_n=1 reproduces
 * `yorigintop = NY*DELTAY - L` of 2D-4th :401; clamp_alpha=1 the x-axis clamp :514. */
int32_t cpml_host_pml_profile(int32_t n, double delta, double deltat, int32_t npoints_pml,
                              int32_t use_pml_min, int32_t use_pml_max,
                              double cp, double rcoef, double npower,
                              double k_max_pml, double alpha_max_pml,
                              int32_t origin_top_uses_n, int32_t clamp_alpha,
                              double *a, double *b, double *K,
                              double *a_half, double *b_half, double *K_half);

/* The same with the viscoelastic program's d0 = -(NPOWER+1)*cp*dsqrt(taumax)*log(Rcoef)/(2L)
 * (3D-visco :547; Rcoef = 1e-4 :541, K_MAX_PML = 7 :243 are the driver's choices). */
int32_t cpml_host_pml_profile_visco(int32_t n, double delta, double deltat, int32_t npoints_pml,
                                    int32_t use_pml_min, int32_t use_pml_max,
                                    double cp, double sqrt_taumax, double rcoef, double npower,
                                    double k_max_pml, double alpha_max_pml,
                                    int32_t origin_top_uses_n, int32_t clamp_alpha,
                                    double *a, double *b, double *K,
                                    double *a_half, double *b_half, double *K_half);

/* First-derivative-of-Gaussian source of :1058-1071 for it = 1..nstep. */
int32_t cpml_host_source_series(int32_t nstep, double deltat, double f0, double t0,
                                double factor, double angle_force_deg,
                                double *force_x, double *force_y);

/* Receiver line + nearest-grid-point search of :683-706. */
int32_t cpml_host_find_receivers(int32_t nx, int32_t ny, double deltax, double deltay,
                                 int32_t nrec, double xdeb, double ydeb, double xfin,
                                 double yfin, int32_t *ix_rec, int32_t *iy_rec, double *dist);

/* Nearest-grid-point search for explicit receiver targets (3D-visco :832-853); the abscissa of
 * grid point i is DELTAX*i when index_origin = 1 (that program), DELTAX*(i-1) when 0. */
int32_t cpml_host_find_receivers_at(int32_t nx, int32_t ny, double deltax, double deltay,
                                    int32_t nrec, const double *xrec, const double *yrec,
                                    int32_t index_origin, int32_t *ix_rec, int32_t *iy_rec,
                                    double *dist);

/* Relaxation times of N_SLS standard linear solids for a constant quality factor Qref around f0:
 * compute_attenuation_coeffs of attenuation_model_with_SolvOpt.f90 :122-169 (classical linear
 * least squares :446-485 as first guess, then SolvOpt :489-1751 on the misfit :1798-1824 under the
 * constraint :1883-1904).  The viscoelastic programs call it once for Qp and once for Qs with
 * f_min = exp(log(f0) - log(12)/2), f_max = 12 f_min (2D-visco-4th :366-376).
 * info (optional, 4 doubles): SolvOpt's options(9) (iterations, or a negative stop code), final
 * misfit, function and gradient evaluations. */
int32_t cpml_host_attenuation_fit(int32_t n_sls, double qref, double f0, double f_min,
                                  double f_max, double *tau_epsilon, double *tau_sigma,
                                  double *info);
/* The linear first guess alone (classical_linear_least_squares :446-485). */
int32_t cpml_host_attenuation_fit_linear(int32_t n_sls, double qref, double f_min, double f_max,
                                         double *tau_epsilon, double *tau_sigma);

/* Courant number of :712 (deltaz <= 0 => 2-D form of 2D-2nd :513). */
double cpml_host_courant(double cp, double deltat, double deltax, double deltay, double deltaz);

/* Output writers with the reference's file names and column layout
 * (write_seismograms :1330-1364; energy.dat :1253-1257 / 2D-2nd :741-746;
 * create_color_image :1371-1509).  List-directed Fortran formatting is compiler
 * specific; these write gnuplot-parsable text with the same columns. */
int32_t cpml_host_write_seismograms(const char *dir, const double *sisvx, const double *sisvy,
                                    int32_t nt, int32_t nrec, double deltat);
/* The viscoelastic programs' write_seismograms: time axis minus t0 (3D-visco :1596-1616); with
 * sispressure != NULL the 2-D form (2D-visco-4th :1145-1193): Vx_file_NNN.dat,
 * Vy_file_half_a_grid_cell_away_from_Vx_NNN.dat and pressure_file_NNN.dat (time + DELTAT/2), the
 * files plotall_fit_is_perfect_for_viscoelastic_fourth_order.gnu reads. */
int32_t cpml_host_write_seismograms_visco(const char *dir, const double *sisvx, const double *sisvy,
                                          const double *sispressure, int32_t nt, int32_t nrec,
                                          double deltat, double t0);
/* timestampNNNNNN, the progress file of the 3-D programs (3D-iso :1219-1229, 3D-visco :1469-1479). */
int32_t cpml_host_write_timestamp(const char *dir, int32_t it, double deltat, double vsolidnorm,
                                  double total_energy, double tcpu);
/* Vz_file_NNN.dat for cpml_get_seismograms_vz (time axis minus t0; t0 = 0 for the isotropic program). */
int32_t cpml_host_write_seismograms_vz(const char *dir, const double *sisvz, int32_t nt, int32_t nrec,
                                       double deltat, double t0);
int32_t cpml_host_write_energy_3d(const char *path, const double *total, int32_t nt,
                                  double deltat);
int32_t cpml_host_write_energy_2d(const char *path, const double *kinetic,
                                  const double *potential, int32_t nt, double deltat);
int32_t cpml_host_create_color_image(const char *dir, const double *image_data_2d,
                                     int32_t nx, int32_t ny, int32_t it,
                                     int32_t isource, int32_t jsource,
                                     const int32_t *ix_rec, const int32_t *iy_rec, int32_t nrec,
                                     int32_t npoints_pml, int32_t use_pml_xmin,
                                     int32_t use_pml_xmax, int32_t use_pml_ymin,
                                     int32_t use_pml_ymax, int32_t field_number);

/* Output fidelity.  The writers above produce, byte for byte, what the gfortran build of the reference writes with
 * its list-directed `write(unit,*)` statements (REAL(4) as 1PG16.9E2, REAL(8) as 1PG25.17E3, 9 / 17 significant digits
 * in F and E editing alike, one leading blank per record, one blank between items; libgfortran io/write.c).
 * cpml_host_format_real returns that text for one item (kind 4: the value is demoted like sngl(); kind 8).
 * cpml_host_write_gnuplot_scripts writes plot_energy, plotgnu (and plot_comparison for the 2-D programs) exactly as
 * the programs do (3D-iso :1260-1313, 2D-2nd :748-806): program 0 = 3-D, 1 = 2-D. */
int32_t cpml_host_format_real(double value, int32_t kind, char *out, int32_t capacity);
int32_t cpml_host_write_gnuplot_scripts(const char *dir, int32_t program);

#ifdef __cplusplus
}
#endif
#endif /* CPML_B200_H */
