!
! seismic_CPML_2D_viscoelastic_b200.f90 -- Fortran host driver for the 2-D viscoelastic C-PML solvers
! (seismic_CPML_2D_velocity_and_stress_{second,fourth}_order_viscoelastic.f90) with the time loop on a
! B200 through libcpml_b200.so (module cpml_b200).
!
! Same parameter names and defaults as the reference programs (:140-230, :310-317), same set-up order
! (relaxation times from Qp, Qs by compute_attenuation_coeffs :366-376, C-PML profiles, Ricker source over
! the cell area :918-935, receiver line :187-199, Courant check :654-657), same display schedule and
! the same output files: Vx_file_NNN.dat, Vy_file_half_a_grid_cell_away_from_Vx_NNN.dat,
! pressure_file_NNN.dat -- the ones plotall_fit_is_perfect_for_viscoelastic_fourth_order.gnu overlays on the
! analytical solution -- and imageNNNNNN_V{x,y}.pnm.  ORDER selects the second- or fourth-order program.
!
! SHIPPED UNCOMPILED: no Fortran compiler exists in the image this repository is built in; the tested
! equivalents are drivers/cpp (--program 2d_visco_fourth) and seismic_cpml_b200/programs.py.
!
program seismic_CPML_2D_visco_b200

  use, intrinsic :: iso_c_binding
  use cpml_b200
  implicit none

  integer(c_int32_t), parameter :: ORDER = 4
  logical, parameter :: VISCOELASTIC_ATTENUATION = .true.
  integer(c_int32_t), parameter :: NX = 2001, NY = 2001
  real(c_double), parameter :: DELTAX = 1.5d0, DELTAY = DELTAX
  logical, parameter :: USE_PML_XMIN = .true., USE_PML_XMAX = .true., USE_PML_YMIN = .true., USE_PML_YMAX = .true.
  integer(c_int32_t), parameter :: NPOINTS_PML = 10
  real(c_double), parameter :: cp_unrelaxed = 2000.d0, cs_unrelaxed = cp_unrelaxed / 1.732d0, density = 2000.d0
  real(c_double), parameter :: DELTAT = 2.2d-4
  integer(c_int32_t), parameter :: NSTEP = 5200
  real(c_double), parameter :: f0 = 35.d0, t0 = 1.20d0 / f0, factor = 1.d0
  real(c_double), parameter :: xsource = 1500.d0, ysource = 1500.d0
  integer(c_int32_t), parameter :: ISOURCE = int(xsource / DELTAX) + 1, JSOURCE = int(ysource / DELTAY) + 1
  real(c_double), parameter :: ANGLE_FORCE = 0.d0
  integer(c_int32_t), parameter :: NREC = 1
  real(c_double), parameter :: xdeb = 2301.d0, ydeb = 2301.d0, xfin = 2301.d0, yfin = 2301.d0
  logical, parameter :: COMPUTE_ENERGY = .false.
  integer(c_int32_t), parameter :: IT_DISPLAY = 200
  real(c_double), parameter :: PI = 3.141592653589793238462643d0, DEGREES_TO_RADIANS = PI / 180.d0
  real(c_double), parameter :: STABILITY_THRESHOLD = 1.d+25
  real(c_double), parameter :: NPOWER = 2.d0, K_MAX_PML = 1.d0, ALPHA_MAX_PML = 2.d0*PI*(f0/2.d0), Rcoef = 0.001d0
  integer(c_int32_t), parameter :: N_SLS = 3
  real(c_double), parameter :: Qp = 65.d0, Qs = 55.d0

  real(c_double) :: a_x(NX), b_x(NX), K_x(NX), a_x_half(NX), b_x_half(NX), K_x_half(NX)
  real(c_double) :: a_y(NY), b_y(NY), K_y(NY), a_y_half(NY), b_y_half(NY), K_y_half(NY)
  real(c_double) :: tau_epsilon_nu1(N_SLS), tau_sigma_nu1(N_SLS), tau_epsilon_nu2(N_SLS), tau_sigma_nu2(N_SLS)
  real(c_double) :: f_min_attenuation, f_max_attenuation, fit_info(4)
  real(c_double) :: force_x(NSTEP), force_y(NSTEP)
  integer(c_int32_t) :: ix_rec(NREC), iy_rec(NREC)
  real(c_double) :: dist_rec(NREC)
  real(c_double), target :: sispressure(NSTEP,NREC)
  real(c_double) :: sisvx(NSTEP,NREC), sisvy(NSTEP,NREC)
  real(c_double) :: total_energy(NSTEP), energy_kinetic(NSTEP), energy_potential(NSTEP)
  real(c_double), allocatable :: lambda(:,:), mu(:,:), rho(:,:), plane(:,:)
  real(c_double) :: Vsolidnorm, Courant_number, a, t, force_source_term

  type(cpml_config) :: cfg
  type(c_ptr) :: h
  integer(c_int32_t) :: ierr, it, it_begin, it_end
  character(kind=c_char, len=2) :: here = '.' // c_null_char

  h = c_null_ptr

! --- relaxation times (:362-380)
  if (VISCOELASTIC_ATTENUATION) then
    f_min_attenuation = exp(log(f0)-log(12.d0)/2.d0)
    f_max_attenuation = 12.d0 * f_min_attenuation
    ierr = cpml_host_attenuation_fit(N_SLS, Qp, f0, f_min_attenuation, f_max_attenuation, tau_epsilon_nu1, tau_sigma_nu1, fit_info)
    if (ierr /= CPML_OK) stop 'compute_attenuation_coeffs failed for Qp'
    ierr = cpml_host_attenuation_fit(N_SLS, Qs, f0, f_min_attenuation, f_max_attenuation, tau_epsilon_nu2, tau_sigma_nu2, fit_info)
    if (ierr /= CPML_OK) stop 'compute_attenuation_coeffs failed for Qs'
  else
    tau_epsilon_nu1(:) = 1.d0;  tau_sigma_nu1(:) = 1.d0;  tau_epsilon_nu2(:) = 1.d0;  tau_sigma_nu2(:) = 1.d0
  endif
  print *,'tau_epsilon_nu1 = ',tau_epsilon_nu1
  print *,'tau_sigma_nu1 = ',tau_sigma_nu1
  print *,'tau_epsilon_nu2 = ',tau_epsilon_nu2
  print *,'tau_sigma_nu2 = ',tau_sigma_nu2

! --- C-PML profiles (:401-593), source (:918-935), receivers (:604-650)
  ierr = cpml_host_pml_profile(NX, DELTAX, DELTAT, NPOINTS_PML, b2i(USE_PML_XMIN), b2i(USE_PML_XMAX), cp_unrelaxed, Rcoef, &
           NPOWER, K_MAX_PML, ALPHA_MAX_PML, 0, 1, a_x, b_x, K_x, a_x_half, b_x_half, K_x_half)
  ierr = cpml_host_pml_profile(NY, DELTAY, DELTAT, NPOINTS_PML, b2i(USE_PML_YMIN), b2i(USE_PML_YMAX), cp_unrelaxed, Rcoef, &
           NPOWER, K_MAX_PML, ALPHA_MAX_PML, 0, 0, a_y, b_y, K_y, a_y_half, b_y_half, K_y_half)

  a = PI*PI*f0*f0
  do it = 1,NSTEP
    t = dble(it-1)*DELTAT
    force_source_term = factor * (1.d0 - 2.d0*a*(t-t0)**2)*exp(-a*(t-t0)**2)
    force_source_term = force_source_term / (DELTAX * DELTAY)
    force_x(it) = sin(ANGLE_FORCE * DEGREES_TO_RADIANS) * force_source_term
    force_y(it) = cos(ANGLE_FORCE * DEGREES_TO_RADIANS) * force_source_term
  enddo

  ierr = cpml_host_find_receivers(NX, NY, DELTAX, DELTAY, NREC, xdeb, ydeb, xfin, yfin, ix_rec, iy_rec, dist_rec)

  Courant_number = cp_unrelaxed * DELTAT / DELTAX
  print *,'Courant number is ',Courant_number
  if (ORDER == 4 .and. Courant_number > 0.606d0) stop 'time step is too large, simulation will be unstable'
  if (ORDER == 2 .and. Courant_number > 1.d0/sqrt(2.d0)) stop 'time step is too large, simulation will be unstable'

! --- unrelaxed material arrays (:596-602), i fastest like the reference's (NX,NY) arrays
  allocate(lambda(NX,NY), mu(NX,NY), rho(NX,NY), plane(NX,NY))
  rho(:,:) = density
  mu(:,:) = density*cs_unrelaxed*cs_unrelaxed
  lambda(:,:) = density*cp_unrelaxed*cp_unrelaxed - 2.d0*mu(:,:)

! --- hand everything to the GPU
  cfg%ndim = 2;  cfg%order = ORDER
  cfg%nx = NX;  cfg%ny = NY;  cfg%nz = 1
  cfg%nstep = NSTEP;  cfg%npoints_pml = NPOINTS_PML;  cfg%nrec = NREC
  cfg%isource = ISOURCE;  cfg%jsource = JSOURCE;  cfg%ksource = 0
  cfg%nslabs = 1;  cfg%slab_rank = 0;  cfg%device = -1;  cfg%energy_bug_compat = 1
  cfg%rheology = 1
  cfg%emulate_nproc = 0
  cfg%compute_energy = b2i(COMPUTE_ENERGY)
  cfg%sigmazz_isotropic = 0
  cfg%deltax = DELTAX;  cfg%deltay = DELTAY;  cfg%deltaz = 0.d0;  cfg%deltat = DELTAT
  cfg%lambda = 0.d0;  cfg%mu = 0.d0;  cfg%lambdaplustwomu = 0.d0;  cfg%rho = 0.d0;  cfg%cp = 0.d0
  cfg%reserved_d = 0.d0

  call cpml_check(cpml_create(cfg, h), h, 'cpml_create')
  call cpml_check(cpml_set_profiles(h, CPML_AXIS_X, a_x, b_x, K_x, a_x_half, b_x_half, K_x_half, NX), h, 'profiles x')
  call cpml_check(cpml_set_profiles(h, CPML_AXIS_Y, a_y, b_y, K_y, a_y_half, b_y_half, K_y_half, NY), h, 'profiles y')
  call cpml_check(cpml_set_material_2d(h, lambda, mu, rho), h, 'material')
  call cpml_check(cpml_set_attenuation(h, N_SLS, tau_epsilon_nu1, tau_sigma_nu1, tau_epsilon_nu2, tau_sigma_nu2), h, 'tau')
  call cpml_check(cpml_set_source_series(h, force_x, force_y, NSTEP), h, 'source')
  call cpml_check(cpml_set_receivers(h, ix_rec, iy_rec, NREC), h, 'receivers')

! --- time loop: the GPU runs up to the next display step, then the driver does its output (:987-1093)
  it_begin = 1
  do while (it_begin <= NSTEP)
    it_end = min(NSTEP, (it_begin / IT_DISPLAY + 1) * IT_DISPLAY)
    if (it_begin <= 5 .and. it_end > 5) it_end = 5
    call cpml_check(cpml_run(h, it_begin, it_end), h, 'cpml_run')
    it = it_end

    if (mod(it,IT_DISPLAY) == 0 .or. it == 5) then
      call cpml_check(cpml_get_maxnorm(h, Vsolidnorm), h, 'maxnorm')
      print *,'Time step # ',it,' out of ',NSTEP
      print *,'Time: ',sngl((it-1)*DELTAT),' seconds'
      print *,'Max norm velocity vector V (m/s) = ',Vsolidnorm
      if (COMPUTE_ENERGY) then
        call cpml_check(cpml_get_energy(h, total_energy, energy_kinetic, energy_potential), h, 'energy')
        print *,'total energy = ',total_energy(it)
      endif
      if (Vsolidnorm > STABILITY_THRESHOLD) stop 'code became unstable and blew up'
      call write_all_seismograms()
      call cpml_check(cpml_get_plane(h, CPML_F_VX, 0, plane), h, 'plane vx')
      ierr = cpml_host_create_color_image(here, plane, NX, NY, it, ISOURCE, JSOURCE, ix_rec, iy_rec, NREC, NPOINTS_PML, &
               b2i(USE_PML_XMIN), b2i(USE_PML_XMAX), b2i(USE_PML_YMIN), b2i(USE_PML_YMAX), 1)
      call cpml_check(cpml_get_plane(h, CPML_F_VY, 0, plane), h, 'plane vy')
      ierr = cpml_host_create_color_image(here, plane, NX, NY, it, ISOURCE, JSOURCE, ix_rec, iy_rec, NREC, NPOINTS_PML, &
               b2i(USE_PML_XMIN), b2i(USE_PML_XMAX), b2i(USE_PML_YMIN), b2i(USE_PML_YMAX), 2)
    endif
    it_begin = it_end + 1
  enddo

! --- final output (:1074-1093)
  call write_all_seismograms()
  if (COMPUTE_ENERGY) then
    call cpml_check(cpml_get_energy(h, total_energy, energy_kinetic, energy_potential), h, 'energy')
    ierr = cpml_host_write_energy_2d('energy.dat' // c_null_char, energy_kinetic, energy_potential, NSTEP, DELTAT)
    ierr = cpml_host_write_gnuplot_scripts('.' // c_null_char, 1)   ! plot_energy, plotgnu
  endif
  ierr = cpml_destroy(h)

  print *
  print *,'End of the simulation'
  print *

contains

  subroutine write_all_seismograms()      ! write_seismograms(sisvx,sisvy,sispressure,NSTEP,NREC,DELTAT,t0), :1145-1193
    call cpml_check(cpml_get_seismograms(h, sisvx, sisvy), h, 'seismograms')
    call cpml_check(cpml_get_pressure_seismograms(h, sispressure), h, 'pressure seismograms')
    ierr = cpml_host_write_seismograms_visco(here, sisvx, sisvy, c_loc(sispressure), NSTEP, NREC, DELTAT, t0)
  end subroutine write_all_seismograms

  integer(c_int32_t) function b2i(flag)
    logical, intent(in) :: flag
    b2i = 0
    if (flag) b2i = 1
  end function b2i

end program seismic_CPML_2D_visco_b200
