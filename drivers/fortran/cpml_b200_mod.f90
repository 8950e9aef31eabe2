!
! cpml_b200_mod.f90 -- ISO_C_BINDING interface to libcpml_b200.so (include/cpml_b200.h).
!
! This is the whole "reference-side binding": a SEISMIC_CPML program keeps its parameter
! block, its set-up phase and its output phase, and replaces the body of
! "do it = 1,NSTEP" by calls to the routines below.
!
! NOTE: no Fortran compiler exists in the image this repository is built in, so this file
! is shipped UNCOMPILED; it is a literal transcription of the C header (same order, same
! types).  tests/test_drivers.py checks it against the header symbol by symbol.
!
! Build (on a machine with gfortran and the CUDA runtime):
!   gfortran -O3 -c cpml_b200_mod.f90
!   gfortran -O3 seismic_CPML_3D_isotropic_b200.f90 cpml_b200_mod.o -L<repo>/seismic_cpml_b200 -lcpml_b200
!
module cpml_b200

  use, intrinsic :: iso_c_binding
  implicit none

  integer(c_int32_t), parameter :: CPML_OK = 0, CPML_EINVAL = 1, CPML_ETOPOLOGY = 2, CPML_ECFL = 3, &
                                   CPML_ECUDA = 4, CPML_ESTATE = 5, CPML_ENOMEM = 6
  integer(c_int32_t), parameter :: CPML_AXIS_X = 0, CPML_AXIS_Y = 1, CPML_AXIS_Z = 2
  integer(c_int32_t), parameter :: CPML_F_VX = 0, CPML_F_VY = 1, CPML_F_VZ = 2, CPML_F_SIGMAXX = 3, &
                                   CPML_F_SIGMAYY = 4, CPML_F_SIGMAZZ = 5, CPML_F_SIGMAXY = 6, &
                                   CPML_F_SIGMAXZ = 7, CPML_F_SIGMAYZ = 8, &
                                   CPML_F_SIGMAXX_R = 9, CPML_F_SIGMAYY_R = 10, CPML_F_SIGMAZZ_R = 11, &
                                   CPML_F_SIGMAXY_R = 12, CPML_F_SIGMAXZ_R = 13, CPML_F_SIGMAYZ_R = 14

! struct cpml_config, field for field
  type, bind(C) :: cpml_config
    integer(c_int32_t) :: ndim, order
    integer(c_int32_t) :: nx, ny, nz
    integer(c_int32_t) :: nstep
    integer(c_int32_t) :: npoints_pml
    integer(c_int32_t) :: nrec
    integer(c_int32_t) :: isource, jsource
    integer(c_int32_t) :: ksource
    integer(c_int32_t) :: nslabs
    integer(c_int32_t) :: slab_rank
    integer(c_int32_t) :: device
    integer(c_int32_t) :: energy_bug_compat
    integer(c_int32_t) :: rheology
    integer(c_int32_t) :: emulate_nproc
    integer(c_int32_t) :: compute_energy
    integer(c_int32_t) :: sigmazz_isotropic   ! 3-D viscoelastic: 0 = the reference's sigmazz memory term (quirk B14) ; 1 = isotropic
    integer(c_int32_t) :: precision           ! 0 = double precision (the reference as shipped) ; 1 = single precision (3D-iso :114-116)
    real(c_double) :: deltax, deltay, deltaz
    real(c_double) :: deltat
    real(c_double) :: lambda, mu, lambdaplustwomu, rho
    real(c_double) :: cp
    real(c_double) :: reserved_d(4)
  end type cpml_config

  interface

    function cpml_abi_version() bind(C, name='cpml_abi_version') result(v)
      import :: c_int32_t
      integer(c_int32_t) :: v
    end function

    function cpml_create(cfg, handle) bind(C, name='cpml_create') result(ierr)
      import :: c_int32_t, c_ptr, cpml_config
      type(cpml_config), intent(in) :: cfg
      type(c_ptr), intent(out) :: handle
      integer(c_int32_t) :: ierr
    end function

    function cpml_destroy(handle) bind(C, name='cpml_destroy') result(ierr)
      import :: c_int32_t, c_ptr
      type(c_ptr), value :: handle
      integer(c_int32_t) :: ierr
    end function

    function cpml_last_error(handle) bind(C, name='cpml_last_error') result(msg)
      import :: c_ptr
      type(c_ptr), value :: handle
      type(c_ptr) :: msg
    end function

    function cpml_reset(handle) bind(C, name='cpml_reset') result(ierr)
      import :: c_int32_t, c_ptr
      type(c_ptr), value :: handle
      integer(c_int32_t) :: ierr
    end function

    function cpml_set_stream(handle, cuda_stream) bind(C, name='cpml_set_stream') result(ierr)
      import :: c_int32_t, c_ptr
      type(c_ptr), value :: handle, cuda_stream
      integer(c_int32_t) :: ierr
    end function

    function cpml_set_profiles(handle, axis, a, b, K, a_half, b_half, K_half, n) &
        bind(C, name='cpml_set_profiles') result(ierr)
      import :: c_int32_t, c_ptr, c_double
      type(c_ptr), value :: handle
      integer(c_int32_t), value :: axis, n
      real(c_double), intent(in) :: a(*), b(*), K(*), a_half(*), b_half(*), K_half(*)
      integer(c_int32_t) :: ierr
    end function

    function cpml_set_material_2d(handle, lambda, mu, rho) bind(C, name='cpml_set_material_2d') result(ierr)
      import :: c_int32_t, c_ptr, c_double
      type(c_ptr), value :: handle
      real(c_double), intent(in) :: lambda(*), mu(*), rho(*)
      integer(c_int32_t) :: ierr
    end function

    function cpml_set_attenuation(handle, n_sls, tau_epsilon_nu1, tau_sigma_nu1, tau_epsilon_nu2, tau_sigma_nu2) &
        bind(C, name='cpml_set_attenuation') result(ierr)
      import :: c_int32_t, c_ptr, c_double
      type(c_ptr), value :: handle
      integer(c_int32_t), value :: n_sls
      real(c_double), intent(in) :: tau_epsilon_nu1(*), tau_sigma_nu1(*), tau_epsilon_nu2(*), tau_sigma_nu2(*)
      integer(c_int32_t) :: ierr
    end function

    function cpml_set_source_series(handle, force_x, force_y, n) bind(C, name='cpml_set_source_series') result(ierr)
      import :: c_int32_t, c_ptr, c_double
      type(c_ptr), value :: handle
      real(c_double), intent(in) :: force_x(*), force_y(*)
      integer(c_int32_t), value :: n
      integer(c_int32_t) :: ierr
    end function

    function cpml_set_receivers(handle, ix_rec, iy_rec, n) bind(C, name='cpml_set_receivers') result(ierr)
      import :: c_int32_t, c_ptr
      type(c_ptr), value :: handle
      integer(c_int32_t), intent(in) :: ix_rec(*), iy_rec(*)
      integer(c_int32_t), value :: n
      integer(c_int32_t) :: ierr
    end function

    function cpml_run(handle, it_begin, it_end) bind(C, name='cpml_run') result(ierr)
      import :: c_int32_t, c_ptr
      type(c_ptr), value :: handle
      integer(c_int32_t), value :: it_begin, it_end
      integer(c_int32_t) :: ierr
    end function

    function cpml_step_stress(handle, it) bind(C, name='cpml_step_stress') result(ierr)
      import :: c_int32_t, c_ptr
      type(c_ptr), value :: handle
      integer(c_int32_t), value :: it
      integer(c_int32_t) :: ierr
    end function

    function cpml_step_velocity(handle, it) bind(C, name='cpml_step_velocity') result(ierr)
      import :: c_int32_t, c_ptr
      type(c_ptr), value :: handle
      integer(c_int32_t), value :: it
      integer(c_int32_t) :: ierr
    end function

    function cpml_step_finish(handle, it) bind(C, name='cpml_step_finish') result(ierr)
      import :: c_int32_t, c_ptr
      type(c_ptr), value :: handle
      integer(c_int32_t), value :: it
      integer(c_int32_t) :: ierr
    end function

    function cpml_synchronize(handle) bind(C, name='cpml_synchronize') result(ierr)
      import :: c_int32_t, c_ptr
      type(c_ptr), value :: handle
      integer(c_int32_t) :: ierr
    end function

    function cpml_halo_plane(handle, field, klocal, device_ptr, nbytes) bind(C, name='cpml_halo_plane') result(ierr)
      import :: c_int32_t, c_int64_t, c_ptr
      type(c_ptr), value :: handle
      integer(c_int32_t), value :: field, klocal
      type(c_ptr), intent(out) :: device_ptr
      integer(c_int64_t), intent(out) :: nbytes
      integer(c_int32_t) :: ierr
    end function

    function cpml_copy_plane(dst, klocal_dst, src, klocal_src, field) bind(C, name='cpml_copy_plane') result(ierr)
      import :: c_int32_t, c_ptr
      type(c_ptr), value :: dst, src
      integer(c_int32_t), value :: klocal_dst, klocal_src, field
      integer(c_int32_t) :: ierr
    end function

    function cpml_p2p_export(handle, blob, blob_capacity, nbytes) bind(C, name='cpml_p2p_export') result(ierr)
      import :: c_int32_t, c_int64_t, c_ptr, c_char
      type(c_ptr), value :: handle
      character(kind=c_char), intent(out) :: blob(*)
      integer(c_int64_t), value :: blob_capacity
      integer(c_int64_t), intent(out) :: nbytes
      integer(c_int32_t) :: ierr
    end function

    function cpml_p2p_attach_ipc(handle, side, blob, nbytes) bind(C, name='cpml_p2p_attach_ipc') result(ierr)
      import :: c_int32_t, c_int64_t, c_ptr, c_char
      type(c_ptr), value :: handle
      integer(c_int32_t), value :: side
      character(kind=c_char), intent(in) :: blob(*)
      integer(c_int64_t), value :: nbytes
      integer(c_int32_t) :: ierr
    end function

    function cpml_p2p_attach_local(handle, side, neighbour) bind(C, name='cpml_p2p_attach_local') result(ierr)
      import :: c_int32_t, c_ptr
      type(c_ptr), value :: handle, neighbour
      integer(c_int32_t), value :: side
      integer(c_int32_t) :: ierr
    end function

    function cpml_p2p_detach(handle) bind(C, name='cpml_p2p_detach') result(ierr)
      import :: c_int32_t, c_ptr
      type(c_ptr), value :: handle
      integer(c_int32_t) :: ierr
    end function

    function cpml_get_launch_info(handle, info, n) bind(C, name='cpml_get_launch_info') result(ierr)
      import :: c_int32_t, c_ptr
      type(c_ptr), value :: handle
      integer(c_int32_t), intent(out) :: info(*)
      integer(c_int32_t), value :: n
      integer(c_int32_t) :: ierr
    end function

    function cpml_set_source_step(handle, it, force_x, force_y) bind(C, name='cpml_set_source_step') result(ierr)
      import :: c_int32_t, c_ptr, c_double
      type(c_ptr), value :: handle
      integer(c_int32_t), value :: it
      real(c_double), value :: force_x, force_y
      integer(c_int32_t) :: ierr
    end function

    function cpml_fetch_step(handle, it) bind(C, name='cpml_fetch_step') result(ierr)
      import :: c_int32_t, c_ptr
      type(c_ptr), value :: handle
      integer(c_int32_t), value :: it
      integer(c_int32_t) :: ierr
    end function

    function cpml_get_fetched_step(handle, it, out4) bind(C, name='cpml_get_fetched_step') result(ierr)
      import :: c_int32_t, c_ptr, c_double
      type(c_ptr), value :: handle
      integer(c_int32_t), value :: it
      real(c_double), intent(out) :: out4(4)
      integer(c_int32_t) :: ierr
    end function

    function cpml_get_seismograms(handle, sisvx, sisvy) bind(C, name='cpml_get_seismograms') result(ierr)
      import :: c_int32_t, c_ptr, c_double
      type(c_ptr), value :: handle
      real(c_double), intent(out) :: sisvx(*), sisvy(*)
      integer(c_int32_t) :: ierr
    end function

    function cpml_get_seismograms_vz(handle, sisvz) bind(C, name='cpml_get_seismograms_vz') result(ierr)
      import :: c_int32_t, c_ptr, c_double
      type(c_ptr), value :: handle
      real(c_double), intent(out) :: sisvz(*)
      integer(c_int32_t) :: ierr
    end function

    function cpml_get_pressure_seismograms(handle, sispressure) bind(C, name='cpml_get_pressure_seismograms') result(ierr)
      import :: c_int32_t, c_ptr, c_double
      type(c_ptr), value :: handle
      real(c_double), intent(out) :: sispressure(*)
      integer(c_int32_t) :: ierr
    end function

    function cpml_get_energy(handle, total, kinetic, potential) bind(C, name='cpml_get_energy') result(ierr)
      import :: c_int32_t, c_ptr, c_double
      type(c_ptr), value :: handle
      real(c_double), intent(out) :: total(*), kinetic(*), potential(*)
      integer(c_int32_t) :: ierr
    end function

    function cpml_get_plane(handle, field, kglobal, plane) bind(C, name='cpml_get_plane') result(ierr)
      import :: c_int32_t, c_ptr, c_double
      type(c_ptr), value :: handle
      integer(c_int32_t), value :: field, kglobal
      real(c_double), intent(out) :: plane(*)
      integer(c_int32_t) :: ierr
    end function

    function cpml_get_field(handle, field, values) bind(C, name='cpml_get_field') result(ierr)
      import :: c_int32_t, c_ptr, c_double
      type(c_ptr), value :: handle
      integer(c_int32_t), value :: field
      real(c_double), intent(out) :: values(*)
      integer(c_int32_t) :: ierr
    end function

    function cpml_get_maxnorm(handle, vnorm) bind(C, name='cpml_get_maxnorm') result(ierr)
      import :: c_int32_t, c_ptr, c_double
      type(c_ptr), value :: handle
      real(c_double), intent(out) :: vnorm
      integer(c_int32_t) :: ierr
    end function

    function cpml_get_kernel_times(handle, ms_stress, ms_velocity, n_launches, reset) &
        bind(C, name='cpml_get_kernel_times') result(ierr)
      import :: c_int32_t, c_int64_t, c_ptr, c_double
      type(c_ptr), value :: handle
      real(c_double), intent(out) :: ms_stress, ms_velocity
      integer(c_int64_t), intent(out) :: n_launches
      integer(c_int32_t), value :: reset
      integer(c_int32_t) :: ierr
    end function

    function cpml_enable_kernel_timing(handle, on) bind(C, name='cpml_enable_kernel_timing') result(ierr)
      import :: c_int32_t, c_ptr
      type(c_ptr), value :: handle
      integer(c_int32_t), value :: on
      integer(c_int32_t) :: ierr
    end function

    function cpml_algorithmic_bytes(handle, bytes_stress, bytes_velocity) &
        bind(C, name='cpml_algorithmic_bytes') result(ierr)
      import :: c_int32_t, c_ptr, c_double
      type(c_ptr), value :: handle
      real(c_double), intent(out) :: bytes_stress, bytes_velocity
      integer(c_int32_t) :: ierr
    end function

    function cpml_host_pml_profile(n, delta, deltat, npoints_pml, use_pml_min, use_pml_max, cp, rcoef, npower, &
        k_max_pml, alpha_max_pml, origin_top_uses_n, clamp_alpha, a, b, K, a_half, b_half, K_half) &
        bind(C, name='cpml_host_pml_profile') result(ierr)
      import :: c_int32_t, c_double
      integer(c_int32_t), value :: n, npoints_pml, use_pml_min, use_pml_max, origin_top_uses_n, clamp_alpha
      real(c_double), value :: delta, deltat, cp, rcoef, npower, k_max_pml, alpha_max_pml
      real(c_double), intent(out) :: a(*), b(*), K(*), a_half(*), b_half(*), K_half(*)
      integer(c_int32_t) :: ierr
    end function

    function cpml_host_pml_profile_visco(n, delta, deltat, npoints_pml, use_pml_min, use_pml_max, cp, sqrt_taumax, &
        rcoef, npower, k_max_pml, alpha_max_pml, origin_top_uses_n, clamp_alpha, a, b, K, a_half, b_half, K_half) &
        bind(C, name='cpml_host_pml_profile_visco') result(ierr)
      import :: c_int32_t, c_double
      integer(c_int32_t), value :: n, npoints_pml, use_pml_min, use_pml_max, origin_top_uses_n, clamp_alpha
      real(c_double), value :: delta, deltat, cp, sqrt_taumax, rcoef, npower, k_max_pml, alpha_max_pml
      real(c_double), intent(out) :: a(*), b(*), K(*), a_half(*), b_half(*), K_half(*)
      integer(c_int32_t) :: ierr
    end function

    function cpml_host_find_receivers_at(nx, ny, deltax, deltay, nrec, xrec, yrec, index_origin, ix_rec, iy_rec, dist) &
        bind(C, name='cpml_host_find_receivers_at') result(ierr)
      import :: c_int32_t, c_double
      integer(c_int32_t), value :: nx, ny, nrec, index_origin
      real(c_double), value :: deltax, deltay
      real(c_double), intent(in) :: xrec(*), yrec(*)
      integer(c_int32_t), intent(out) :: ix_rec(*), iy_rec(*)
      real(c_double), intent(out) :: dist(*)
      integer(c_int32_t) :: ierr
    end function

    function cpml_host_source_series(nstep, deltat, f0, t0, factor, angle_force_deg, force_x, force_y) &
        bind(C, name='cpml_host_source_series') result(ierr)
      import :: c_int32_t, c_double
      integer(c_int32_t), value :: nstep
      real(c_double), value :: deltat, f0, t0, factor, angle_force_deg
      real(c_double), intent(out) :: force_x(*), force_y(*)
      integer(c_int32_t) :: ierr
    end function

    function cpml_host_find_receivers(nx, ny, deltax, deltay, nrec, xdeb, ydeb, xfin, yfin, ix_rec, iy_rec, dist) &
        bind(C, name='cpml_host_find_receivers') result(ierr)
      import :: c_int32_t, c_double
      integer(c_int32_t), value :: nx, ny, nrec
      real(c_double), value :: deltax, deltay, xdeb, ydeb, xfin, yfin
      integer(c_int32_t), intent(out) :: ix_rec(*), iy_rec(*)
      real(c_double), intent(out) :: dist(*)
      integer(c_int32_t) :: ierr
    end function

    function cpml_host_courant(cp, deltat, deltax, deltay, deltaz) bind(C, name='cpml_host_courant') result(c)
      import :: c_double
      real(c_double), value :: cp, deltat, deltax, deltay, deltaz
      real(c_double) :: c
    end function

    ! compute_attenuation_coeffs of attenuation_model_with_SolvOpt.f90 :122-169
    function cpml_host_attenuation_fit(n_sls, qref, f0, f_min, f_max, tau_epsilon, tau_sigma, info) &
        bind(C, name='cpml_host_attenuation_fit') result(ierr)
      import :: c_int32_t, c_double
      integer(c_int32_t), value :: n_sls
      real(c_double), value :: qref, f0, f_min, f_max
      real(c_double), intent(out) :: tau_epsilon(*), tau_sigma(*)
      real(c_double), intent(out) :: info(4)
      integer(c_int32_t) :: ierr
    end function

    function cpml_host_attenuation_fit_linear(n_sls, qref, f_min, f_max, tau_epsilon, tau_sigma) &
        bind(C, name='cpml_host_attenuation_fit_linear') result(ierr)
      import :: c_int32_t, c_double
      integer(c_int32_t), value :: n_sls
      real(c_double), value :: qref, f_min, f_max
      real(c_double), intent(out) :: tau_epsilon(*), tau_sigma(*)
      integer(c_int32_t) :: ierr
    end function

    function cpml_host_write_seismograms(dir, sisvx, sisvy, nt, nrec, deltat) &
        bind(C, name='cpml_host_write_seismograms') result(ierr)
      import :: c_int32_t, c_double, c_char
      character(kind=c_char), intent(in) :: dir(*)
      real(c_double), intent(in) :: sisvx(*), sisvy(*)
      integer(c_int32_t), value :: nt, nrec
      real(c_double), value :: deltat
      integer(c_int32_t) :: ierr
    end function

    function cpml_host_write_seismograms_visco(dir, sisvx, sisvy, sispressure, nt, nrec, deltat, t0) &
        bind(C, name='cpml_host_write_seismograms_visco') result(ierr)
      import :: c_int32_t, c_double, c_char, c_ptr
      character(kind=c_char), intent(in) :: dir(*)
      real(c_double), intent(in) :: sisvx(*), sisvy(*)
      type(c_ptr), value :: sispressure        ! c_loc(sispressure) in the 2-D programs, c_null_ptr in 3-D
      integer(c_int32_t), value :: nt, nrec
      real(c_double), value :: deltat, t0
      integer(c_int32_t) :: ierr
    end function

    function cpml_host_write_seismograms_vz(dir, sisvz, nt, nrec, deltat, t0) &
        bind(C, name='cpml_host_write_seismograms_vz') result(ierr)
      import :: c_int32_t, c_double, c_char
      character(kind=c_char), intent(in) :: dir(*)
      real(c_double), intent(in) :: sisvz(*)
      integer(c_int32_t), value :: nt, nrec
      real(c_double), value :: deltat, t0
      integer(c_int32_t) :: ierr
    end function

    function cpml_host_write_timestamp(dir, it, deltat, vsolidnorm, total_energy, tcpu) &
        bind(C, name='cpml_host_write_timestamp') result(ierr)
      import :: c_int32_t, c_double, c_char
      character(kind=c_char), intent(in) :: dir(*)
      integer(c_int32_t), value :: it
      real(c_double), value :: deltat, vsolidnorm, total_energy, tcpu
      integer(c_int32_t) :: ierr
    end function

    function cpml_host_write_energy_3d(path, total, nt, deltat) bind(C, name='cpml_host_write_energy_3d') result(ierr)
      import :: c_int32_t, c_double, c_char
      character(kind=c_char), intent(in) :: path(*)
      real(c_double), intent(in) :: total(*)
      integer(c_int32_t), value :: nt
      real(c_double), value :: deltat
      integer(c_int32_t) :: ierr
    end function

    function cpml_host_write_energy_2d(path, kinetic, potential, nt, deltat) &
        bind(C, name='cpml_host_write_energy_2d') result(ierr)
      import :: c_int32_t, c_double, c_char
      character(kind=c_char), intent(in) :: path(*)
      real(c_double), intent(in) :: kinetic(*), potential(*)
      integer(c_int32_t), value :: nt
      real(c_double), value :: deltat
      integer(c_int32_t) :: ierr
    end function

    function cpml_host_create_color_image(dir, image_data_2d, nx, ny, it, isource, jsource, ix_rec, iy_rec, nrec, &
        npoints_pml, use_pml_xmin, use_pml_xmax, use_pml_ymin, use_pml_ymax, field_number) &
        bind(C, name='cpml_host_create_color_image') result(ierr)
      import :: c_int32_t, c_double, c_char
      character(kind=c_char), intent(in) :: dir(*)
      real(c_double), intent(in) :: image_data_2d(*)
      integer(c_int32_t), value :: nx, ny, it, isource, jsource, nrec, npoints_pml
      integer(c_int32_t), intent(in) :: ix_rec(*), iy_rec(*)
      integer(c_int32_t), value :: use_pml_xmin, use_pml_xmax, use_pml_ymin, use_pml_ymax, field_number
      integer(c_int32_t) :: ierr
    end function

    function cpml_snapshot_begin(handle, slot, field, kglobal) bind(C, name='cpml_snapshot_begin') result(ierr)
      import :: c_int32_t, c_ptr
      type(c_ptr), value :: handle
      integer(c_int32_t), value :: slot, field, kglobal
      integer(c_int32_t) :: ierr
    end function

    function cpml_snapshot_end(handle, slot, plane, pinned) bind(C, name='cpml_snapshot_end') result(ierr)
      import :: c_int32_t, c_ptr
      type(c_ptr), value :: handle
      integer(c_int32_t), value :: slot
      type(c_ptr), value :: plane          ! c_loc of a real(c_double) (NX,NY) array, or c_null_ptr
      type(c_ptr), value :: pinned         ! c_loc of a type(c_ptr) that receives the pinned buffer, or c_null_ptr
      integer(c_int32_t) :: ierr
    end function

    ! ---- the whole z-slab decomposition behind one handle (one host thread, no MPI): cpml_multi_* ----
    function cpml_multi_create(cfg, ngpus, devices, multi) bind(C, name='cpml_multi_create') result(ierr)
      import :: c_int32_t, c_ptr, cpml_config
      type(cpml_config), intent(in) :: cfg
      integer(c_int32_t), value :: ngpus
      type(c_ptr), value :: devices          ! c_null_ptr: devices 0..ngpus-1, else c_loc of an integer(c_int32_t) array
      type(c_ptr), intent(out) :: multi
      integer(c_int32_t) :: ierr
    end function

    function cpml_multi_destroy(multi) bind(C, name='cpml_multi_destroy') result(ierr)
      import :: c_int32_t, c_ptr
      type(c_ptr), value :: multi
      integer(c_int32_t) :: ierr
    end function

    function cpml_multi_last_error(multi) bind(C, name='cpml_multi_last_error') result(msg)
      import :: c_ptr
      type(c_ptr), value :: multi
      type(c_ptr) :: msg
    end function

    function cpml_multi_ngpus(multi) bind(C, name='cpml_multi_ngpus') result(n)
      import :: c_int32_t, c_ptr
      type(c_ptr), value :: multi
      integer(c_int32_t) :: n
    end function

    function cpml_multi_slab(multi, rank, handle) bind(C, name='cpml_multi_slab') result(ierr)
      import :: c_int32_t, c_ptr
      type(c_ptr), value :: multi
      integer(c_int32_t), value :: rank
      type(c_ptr), intent(out) :: handle
      integer(c_int32_t) :: ierr
    end function

    function cpml_multi_reset(multi) bind(C, name='cpml_multi_reset') result(ierr)
      import :: c_int32_t, c_ptr
      type(c_ptr), value :: multi
      integer(c_int32_t) :: ierr
    end function

    function cpml_multi_set_profiles(multi, axis, a, b, K, a_half, b_half, K_half, n) &
        bind(C, name='cpml_multi_set_profiles') result(ierr)
      import :: c_int32_t, c_ptr, c_double
      type(c_ptr), value :: multi
      integer(c_int32_t), value :: axis, n
      real(c_double), intent(in) :: a(*), b(*), K(*), a_half(*), b_half(*), K_half(*)
      integer(c_int32_t) :: ierr
    end function

    function cpml_multi_set_attenuation(multi, n_sls, tau_epsilon_nu1, tau_sigma_nu1, tau_epsilon_nu2, tau_sigma_nu2) &
        bind(C, name='cpml_multi_set_attenuation') result(ierr)
      import :: c_int32_t, c_ptr, c_double
      type(c_ptr), value :: multi
      integer(c_int32_t), value :: n_sls
      real(c_double), intent(in) :: tau_epsilon_nu1(*), tau_sigma_nu1(*), tau_epsilon_nu2(*), tau_sigma_nu2(*)
      integer(c_int32_t) :: ierr
    end function

    function cpml_multi_set_source_series(multi, force_x, force_y, n) bind(C, name='cpml_multi_set_source_series') result(ierr)
      import :: c_int32_t, c_ptr, c_double
      type(c_ptr), value :: multi
      real(c_double), intent(in) :: force_x(*), force_y(*)
      integer(c_int32_t), value :: n
      integer(c_int32_t) :: ierr
    end function

    function cpml_multi_set_receivers(multi, ix_rec, iy_rec, n) bind(C, name='cpml_multi_set_receivers') result(ierr)
      import :: c_int32_t, c_ptr
      type(c_ptr), value :: multi
      integer(c_int32_t), intent(in) :: ix_rec(*), iy_rec(*)
      integer(c_int32_t), value :: n
      integer(c_int32_t) :: ierr
    end function

    function cpml_multi_step(multi, it) bind(C, name='cpml_multi_step') result(ierr)
      import :: c_int32_t, c_ptr
      type(c_ptr), value :: multi
      integer(c_int32_t), value :: it
      integer(c_int32_t) :: ierr
    end function

    function cpml_multi_run(multi, it_begin, it_end) bind(C, name='cpml_multi_run') result(ierr)
      import :: c_int32_t, c_ptr
      type(c_ptr), value :: multi
      integer(c_int32_t), value :: it_begin, it_end
      integer(c_int32_t) :: ierr
    end function

    function cpml_multi_synchronize(multi) bind(C, name='cpml_multi_synchronize') result(ierr)
      import :: c_int32_t, c_ptr
      type(c_ptr), value :: multi
      integer(c_int32_t) :: ierr
    end function

    function cpml_multi_get_seismograms(multi, sisvx, sisvy) bind(C, name='cpml_multi_get_seismograms') result(ierr)
      import :: c_int32_t, c_ptr, c_double
      type(c_ptr), value :: multi
      real(c_double), intent(out) :: sisvx(*), sisvy(*)
      integer(c_int32_t) :: ierr
    end function

    function cpml_multi_get_seismograms_vz(multi, sisvz) bind(C, name='cpml_multi_get_seismograms_vz') result(ierr)
      import :: c_int32_t, c_ptr, c_double
      type(c_ptr), value :: multi
      real(c_double), intent(out) :: sisvz(*)
      integer(c_int32_t) :: ierr
    end function

    function cpml_multi_get_energy(multi, total, kinetic, potential) bind(C, name='cpml_multi_get_energy') result(ierr)
      import :: c_int32_t, c_ptr, c_double
      type(c_ptr), value :: multi
      real(c_double), intent(out) :: total(*), kinetic(*), potential(*)
      integer(c_int32_t) :: ierr
    end function

    function cpml_multi_get_plane(multi, field, kglobal, plane) bind(C, name='cpml_multi_get_plane') result(ierr)
      import :: c_int32_t, c_ptr, c_double
      type(c_ptr), value :: multi
      integer(c_int32_t), value :: field, kglobal
      real(c_double), intent(out) :: plane(*)
      integer(c_int32_t) :: ierr
    end function

    function cpml_multi_get_field(multi, field, values) bind(C, name='cpml_multi_get_field') result(ierr)
      import :: c_int32_t, c_ptr, c_double
      type(c_ptr), value :: multi
      integer(c_int32_t), value :: field
      real(c_double), intent(out) :: values(*)
      integer(c_int32_t) :: ierr
    end function

    function cpml_multi_get_maxnorm(multi, vmax) bind(C, name='cpml_multi_get_maxnorm') result(ierr)
      import :: c_int32_t, c_ptr, c_double
      type(c_ptr), value :: multi
      real(c_double), intent(out) :: vmax
      integer(c_int32_t) :: ierr
    end function

    function cpml_host_format_real(value, kind, text, capacity) bind(C, name='cpml_host_format_real') result(ierr)
      import :: c_int32_t, c_double, c_char
      real(c_double), value :: value
      integer(c_int32_t), value :: kind, capacity
      character(kind=c_char), intent(out) :: text(*)
      integer(c_int32_t) :: ierr
    end function

    function cpml_host_write_gnuplot_scripts(dir, program) bind(C, name='cpml_host_write_gnuplot_scripts') result(ierr)
      import :: c_int32_t, c_char
      character(kind=c_char), intent(in) :: dir(*)
      integer(c_int32_t), value :: program
      integer(c_int32_t) :: ierr
    end function

  end interface

contains

! stop with the library's message, like the reference's "stop '...'" statements
  subroutine cpml_check(ierr, handle, where)
    integer(c_int32_t), intent(in) :: ierr
    type(c_ptr), intent(in) :: handle
    character(len=*), intent(in) :: where
    type(c_ptr) :: msg
    character(kind=c_char), pointer :: chars(:)
    integer :: n
    if (ierr == CPML_OK) return
    msg = cpml_last_error(handle)
    if (c_associated(msg)) then
      call c_f_pointer(msg, chars, [512])
      n = 0
      do while (n < 512)
        if (chars(n+1) == c_null_char) exit
        n = n + 1
      enddo
      print *,'libcpml_b200 error in ',where,': ',chars(1:n)
    endif
    stop 'libcpml_b200 call failed'
  end subroutine cpml_check

end module cpml_b200
