!
! seismic_CPML_3D_viscoelastic_b200.f90 -- Fortran host driver for the 3-D viscoelastic C-PML solver
! (seismic_CPML_3D_viscoelastic_MPI.f90) with the time loop on a B200 through libcpml_b200.so (module cpml_b200).
!
! Same parameter names and defaults as the reference program (:152-244), same set-up order (relaxation times
! from QKappa_att, QMu_att, f0_attenuation by compute_attenuation_coeffs :433-443, taumax :450-455, C-PML
! profiles with d0 ~ cp sqrt(taumax) :533-801, explicit receiver coordinates :825-853, Courant check :856),
! same display schedule and outputs (Vx_file_NNN.dat, Vy_file_NNN.dat with the time axis minus t0, energy.dat
! with kinetic / potential / total columns, timestampNNNNNN, imageNNNNNN_V{x,y}.pnm).  Single process, whole
! grid on one GPU; NPROC is still a parameter because the reference's incomplete halo exchange makes its
! result depend on it (cfg%emulate_nproc, SURVEY.md quirk B6).
!
! SHIPPED UNCOMPILED: no Fortran compiler exists in the image this repository is built in; the tested
! equivalents are drivers/cpp (--program 3d_visco) and seismic_cpml_b200/programs.py.
!
program seismic_visco_CPML_3D_b200

  use, intrinsic :: iso_c_binding
  use cpml_b200
  implicit none

  integer(c_int32_t), parameter :: NX = 210, NY = 800, NZ = 220
  integer(c_int32_t), parameter :: NPROC = 4
  real(c_double), parameter :: DELTAX = 4.d0, DELTAY = DELTAX, DELTAZ = DELTAX
  real(c_double), parameter :: cp = 3000.d0, cs = 2000.d0, rho = 2000.d0
  real(c_double), parameter :: mu = rho*cs*cs, lambda = rho*(cp*cp - 2.d0*cs*cs)
  integer(c_int32_t), parameter :: NSTEP = 100000
  real(c_double), parameter :: DELTAT = 4.d-4
  real(c_double), parameter :: f0 = 18.d0, t0 = 1.20d0 / f0, factor = 1.d7
  integer(c_int32_t), parameter :: N_SLS = 2
  real(c_double), parameter :: QKappa_att = 20.d0, QMu_att = 10.d0, f0_attenuation = 16.d0
  logical, parameter :: USE_PML_XMIN = .true., USE_PML_XMAX = .true., USE_PML_YMIN = .true., &
                        USE_PML_YMAX = .true., USE_PML_ZMIN = .true., USE_PML_ZMAX = .true.
  integer(c_int32_t), parameter :: NPOINTS_PML = 10
  integer(c_int32_t), parameter :: ISOURCE = NPOINTS_PML + 20, JSOURCE = NY / 5 + 1
  real(c_double), parameter :: xsource = ISOURCE * DELTAX, ysource = JSOURCE * DELTAY
  real(c_double), parameter :: ANGLE_FORCE = 0.d0
  integer(c_int32_t), parameter :: NREC = 3
  integer(c_int32_t), parameter :: IT_DISPLAY = 10000
  real(c_double), parameter :: PI = 3.141592653589793238462643d0
  real(c_double), parameter :: STABILITY_THRESHOLD = 1.d+25
  real(c_double), parameter :: NPOWER = 2.d0, K_MAX_PML = 7.d0, ALPHA_MAX_PML = 2.d0*PI*(f0/2.d0), Rcoef = 0.0001d0

  real(c_double) :: a_x(NX), b_x(NX), K_x(NX), a_x_half(NX), b_x_half(NX), K_x_half(NX)
  real(c_double) :: a_y(NY), b_y(NY), K_y(NY), a_y_half(NY), b_y_half(NY), K_y_half(NY)
  real(c_double) :: a_z(NZ), b_z(NZ), K_z(NZ), a_z_half(NZ), b_z_half(NZ), K_z_half(NZ)
  real(c_double) :: tau_epsilon_nu1(N_SLS), tau_sigma_nu1(N_SLS), tau_epsilon_nu2(N_SLS), tau_sigma_nu2(N_SLS)
  real(c_double) :: f_min_attenuation, f_max_attenuation, fit_info(4), taumax, sqrt_taumax
  real(c_double) :: force_x(NSTEP), force_y(NSTEP)
  real(c_double) :: xrec(NREC), yrec(NREC), dist_rec(NREC)
  integer(c_int32_t) :: ix_rec(NREC), iy_rec(NREC)
  real(c_double), allocatable :: sisvx(:,:), sisvy(:,:), total_energy(:), energy_kinetic(:), energy_potential(:), plane(:,:)
  real(c_double) :: Vsolidnorm, Courant_number, time_start, tCPU
  integer :: time_values(8), l

  type(cpml_config) :: cfg
  type(c_ptr) :: h
  integer(c_int32_t) :: ierr, it, it_begin, it_end
  character(kind=c_char, len=2) :: here = '.' // c_null_char

  h = c_null_ptr
  allocate(sisvx(NSTEP,NREC), sisvy(NSTEP,NREC), total_energy(NSTEP), energy_kinetic(NSTEP), energy_potential(NSTEP))
  allocate(plane(NX,NY))

! --- relaxation times (:433-443) and the constants derived from them (:450-455)
  f_min_attenuation = exp(log(f0_attenuation)-log(12.d0)/2.d0)
  f_max_attenuation = 12.d0 * f_min_attenuation
  ierr = cpml_host_attenuation_fit(N_SLS, QKappa_att, f0_attenuation, f_min_attenuation, f_max_attenuation, &
           tau_epsilon_nu1, tau_sigma_nu1, fit_info)
  if (ierr /= CPML_OK) stop 'compute_attenuation_coeffs failed for QKappa_att'
  ierr = cpml_host_attenuation_fit(N_SLS, QMu_att, f0_attenuation, f_min_attenuation, f_max_attenuation, &
           tau_epsilon_nu2, tau_sigma_nu2, fit_info)
  if (ierr /= CPML_OK) stop 'compute_attenuation_coeffs failed for QMu_att'
  print *,'tau_epsilon_nu1 = ',tau_epsilon_nu1
  print *,'tau_sigma_nu1 = ',tau_sigma_nu1
  print *,'tau_epsilon_nu2 = ',tau_epsilon_nu2
  print *,'tau_sigma_nu2 = ',tau_sigma_nu2
  taumax = 0.d0
  do l = 1,N_SLS
    taumax = max(taumax, 1.d0/(tau_sigma_nu1(l)/tau_epsilon_nu1(l)), 1.d0/(tau_sigma_nu2(l)/tau_epsilon_nu2(l)))
  enddo
  sqrt_taumax = dsqrt(taumax)

! --- C-PML profiles (:533-801), source (:1310-1322), receivers (:825-853), Courant number (:856)
  ierr = cpml_host_pml_profile_visco(NX, DELTAX, DELTAT, NPOINTS_PML, b2i(USE_PML_XMIN), b2i(USE_PML_XMAX), cp, sqrt_taumax, &
           Rcoef, NPOWER, K_MAX_PML, ALPHA_MAX_PML, 0, 1, a_x, b_x, K_x, a_x_half, b_x_half, K_x_half)
  ierr = cpml_host_pml_profile_visco(NY, DELTAY, DELTAT, NPOINTS_PML, b2i(USE_PML_YMIN), b2i(USE_PML_YMAX), cp, sqrt_taumax, &
           Rcoef, NPOWER, K_MAX_PML, ALPHA_MAX_PML, 0, 0, a_y, b_y, K_y, a_y_half, b_y_half, K_y_half)
  ierr = cpml_host_pml_profile_visco(NZ, DELTAZ, DELTAT, NPOINTS_PML, b2i(USE_PML_ZMIN), b2i(USE_PML_ZMAX), cp, sqrt_taumax, &
           Rcoef, NPOWER, K_MAX_PML, ALPHA_MAX_PML, 0, 0, a_z, b_z, K_z, a_z_half, b_z_half, K_z_half)
  ierr = cpml_host_source_series(NSTEP, DELTAT, f0, t0, factor, ANGLE_FORCE, force_x, force_y)
  xrec(1) = xsource + 500.d0;  yrec(1) = ysource + 500.d0
  xrec(2) = xsource;           yrec(2) = ysource + 2260.d0
  xrec(3) = xsource + 500.d0;  yrec(3) = ysource + 2260.d0
  ierr = cpml_host_find_receivers_at(NX, NY, DELTAX, DELTAY, NREC, xrec, yrec, 1, ix_rec, iy_rec, dist_rec)

  Courant_number = cpml_host_courant(cp*sqrt_taumax, DELTAT, DELTAX, DELTAY, DELTAZ)
  print *,'Courant number is ',Courant_number
  if (Courant_number > 1.d0) stop 'time step is too large, simulation will be unstable'

! --- hand everything to the GPU
  cfg%ndim = 3;  cfg%order = 4
  cfg%nx = NX;  cfg%ny = NY;  cfg%nz = NZ
  cfg%nstep = NSTEP;  cfg%npoints_pml = NPOINTS_PML;  cfg%nrec = NREC
  cfg%isource = ISOURCE;  cfg%jsource = JSOURCE;  cfg%ksource = 0
  cfg%nslabs = 1;  cfg%slab_rank = 0;  cfg%device = -1;  cfg%energy_bug_compat = 1
  cfg%rheology = 1
  cfg%emulate_nproc = NPROC
  cfg%compute_energy = 0
  cfg%sigmazz_isotropic = 0
  cfg%deltax = DELTAX;  cfg%deltay = DELTAY;  cfg%deltaz = DELTAZ;  cfg%deltat = DELTAT
  cfg%lambda = lambda;  cfg%mu = mu;  cfg%lambdaplustwomu = 0.d0;  cfg%rho = rho;  cfg%cp = cp*sqrt_taumax
  cfg%reserved_d = 0.d0

  call cpml_check(cpml_create(cfg, h), h, 'cpml_create')
  call cpml_check(cpml_set_profiles(h, CPML_AXIS_X, a_x, b_x, K_x, a_x_half, b_x_half, K_x_half, NX), h, 'profiles x')
  call cpml_check(cpml_set_profiles(h, CPML_AXIS_Y, a_y, b_y, K_y, a_y_half, b_y_half, K_y_half, NY), h, 'profiles y')
  call cpml_check(cpml_set_profiles(h, CPML_AXIS_Z, a_z, b_z, K_z, a_z_half, b_z_half, K_z_half, NZ), h, 'profiles z')
  call cpml_check(cpml_set_attenuation(h, N_SLS, tau_epsilon_nu1, tau_sigma_nu1, tau_epsilon_nu2, tau_sigma_nu2), h, 'tau')
  call cpml_check(cpml_set_source_series(h, force_x, force_y, NSTEP), h, 'source')
  call cpml_check(cpml_set_receivers(h, ix_rec, iy_rec, NREC), h, 'receivers')

  call date_and_time(values=time_values)
  time_start = 86400.d0*time_values(3) + 3600.d0*time_values(5) + 60.d0*time_values(6) + time_values(7) + time_values(8)/1000.d0

! --- time loop: the GPU runs up to the next display step, then the driver does its output (:1433-1503)
  it_begin = 1
  do while (it_begin <= NSTEP)
    it_end = min(NSTEP, (it_begin / IT_DISPLAY + 1) * IT_DISPLAY)
    if (it_begin <= 5 .and. it_end > 5) it_end = 5
    call cpml_check(cpml_run(h, it_begin, it_end), h, 'cpml_run')
    it = it_end

    if (mod(it,IT_DISPLAY) == 0 .or. it == 5) then
      call cpml_check(cpml_get_maxnorm(h, Vsolidnorm), h, 'maxnorm')
      call cpml_check(cpml_get_energy(h, total_energy, energy_kinetic, energy_potential), h, 'energy')
      print *,'Time step # ',it,' out of ',NSTEP
      print *,'Time: ',sngl((it-1)*DELTAT),' seconds'
      print *,'Max norm velocity vector V (m/s) = ',Vsolidnorm
      print *,'Total energy = ',total_energy(it)
      if (Vsolidnorm > STABILITY_THRESHOLD) stop 'code became unstable and blew up in solid'
      call date_and_time(values=time_values)
      tCPU = 86400.d0*time_values(3) + 3600.d0*time_values(5) + 60.d0*time_values(6) + time_values(7) + &
             time_values(8)/1000.d0 - time_start
      ierr = cpml_host_write_timestamp(here, it, DELTAT, Vsolidnorm, total_energy(it), tCPU)
      call cpml_check(cpml_get_seismograms(h, sisvx, sisvy), h, 'seismograms')
      ierr = cpml_host_write_seismograms_visco(here, sisvx, sisvy, c_null_ptr, NSTEP, NREC, DELTAT, t0)
      call cpml_check(cpml_get_plane(h, CPML_F_VX, NZ/2, plane), h, 'plane vx')
      ierr = cpml_host_create_color_image(here, plane, NX, NY, it, ISOURCE, JSOURCE, ix_rec, iy_rec, NREC, NPOINTS_PML, &
               b2i(USE_PML_XMIN), b2i(USE_PML_XMAX), b2i(USE_PML_YMIN), b2i(USE_PML_YMAX), 1)
      call cpml_check(cpml_get_plane(h, CPML_F_VY, NZ/2, plane), h, 'plane vy')
      ierr = cpml_host_create_color_image(here, plane, NX, NY, it, ISOURCE, JSOURCE, ix_rec, iy_rec, NREC, NPOINTS_PML, &
               b2i(USE_PML_XMIN), b2i(USE_PML_XMAX), b2i(USE_PML_YMIN), b2i(USE_PML_YMAX), 2)
    endif
    it_begin = it_end + 1
  enddo

! --- final output (:1505-1560)
  call cpml_check(cpml_get_seismograms(h, sisvx, sisvy), h, 'seismograms')
  ierr = cpml_host_write_seismograms_visco(here, sisvx, sisvy, c_null_ptr, NSTEP, NREC, DELTAT, t0)
  call cpml_check(cpml_get_energy(h, total_energy, energy_kinetic, energy_potential), h, 'energy')
  ierr = cpml_host_write_energy_2d('energy.dat' // c_null_char, energy_kinetic, energy_potential, NSTEP, DELTAT)
  ierr = cpml_host_write_gnuplot_scripts('.' // c_null_char, 0)   ! plot_energy, plotgnu
  ierr = cpml_destroy(h)

  print *
  print *,'End of the simulation'
  print *

contains

  integer(c_int32_t) function b2i(flag)
    logical, intent(in) :: flag
    b2i = 0
    if (flag) b2i = 1
  end function b2i

end program seismic_visco_CPML_3D_b200
