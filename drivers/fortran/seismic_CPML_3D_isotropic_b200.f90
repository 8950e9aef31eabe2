!
! seismic_CPML_3D_isotropic_b200.f90 -- Fortran host driver for the 3-D isotropic C-PML solver
! with the time loop on a B200 through libcpml_b200.so (module cpml_b200).
!
! Same parameter names and defaults as seismic_CPML_3D_isotropic_MPI_OpenMP.f90:124-218, same
! outputs (Vx_file_NNN.dat, Vy_file_NNN.dat, energy.dat, imageNNNNNN_V{x,y}.pnm); the set-up
! formulas and file writers live in the library (cpml_host_*), so this program is only the
! parameter block, the display schedule and the calls.  Single process, whole grid on one GPU.
! An MPI build would create one handle per rank with nslabs = nb_procs, slab_rank = rank and
! interleave cpml_step_stress / cpml_step_velocity with its MPI_SENDRECV on the device
! pointers returned by cpml_halo_plane (CUDA-aware MPI) -- see INTEGRATION.md.
!
! SHIPPED UNCOMPILED: no Fortran compiler exists in the image this repository is built in.
!
program seismic_CPML_3D_iso_b200

  use, intrinsic :: iso_c_binding
  use cpml_b200
  implicit none

  integer(c_int32_t), parameter :: NX = 101, NY = 641, NZ = 640
  real(c_double), parameter :: DELTAX = 10.d0, DELTAY = DELTAX, DELTAZ = DELTAX
  real(c_double), parameter :: cp = 3300.d0, cs = cp / 1.732d0, rho = 2800.d0
  real(c_double), parameter :: mu = rho*cs*cs, lambda = rho*(cp*cp - 2.d0*cs*cs), lambdaplustwomu = rho*cp*cp
  integer(c_int32_t), parameter :: NSTEP = 2500
  real(c_double), parameter :: DELTAT = 1.6d-3
  real(c_double), parameter :: f0 = 7.d0, t0 = 1.20d0 / f0, factor = 1.d7
  logical, parameter :: USE_PML_XMIN = .true., USE_PML_XMAX = .true., USE_PML_YMIN = .true., &
                        USE_PML_YMAX = .true., USE_PML_ZMIN = .true., USE_PML_ZMAX = .true.
  integer(c_int32_t), parameter :: NPOINTS_PML = 10
  integer(c_int32_t), parameter :: ISOURCE = NX - 2*NPOINTS_PML - 1, JSOURCE = 2 * NY / 3 + 1
  real(c_double), parameter :: xsource = (ISOURCE - 1) * DELTAX
  real(c_double), parameter :: ANGLE_FORCE = 135.d0
  integer(c_int32_t), parameter :: NREC = 2
  real(c_double), parameter :: xdeb = xsource - 100.d0, ydeb = 2300.d0, xfin = xsource, yfin = 300.d0
  integer(c_int32_t), parameter :: IT_DISPLAY = 100
  real(c_double), parameter :: PI = 3.141592653589793238462643d0
  real(c_double), parameter :: STABILITY_THRESHOLD = 1.d+25
  real(c_double), parameter :: NPOWER = 2.d0, K_MAX_PML = 1.d0, ALPHA_MAX_PML = 2.d0*PI*(f0/2.d0), Rcoef = 0.001d0

  real(c_double) :: a_x(NX), b_x(NX), K_x(NX), a_x_half(NX), b_x_half(NX), K_x_half(NX)
  real(c_double) :: a_y(NY), b_y(NY), K_y(NY), a_y_half(NY), b_y_half(NY), K_y_half(NY)
  real(c_double) :: a_z(NZ), b_z(NZ), K_z(NZ), a_z_half(NZ), b_z_half(NZ), K_z_half(NZ)
  real(c_double) :: force_x(NSTEP), force_y(NSTEP)
  integer(c_int32_t) :: ix_rec(NREC), iy_rec(NREC)
  real(c_double) :: dist_rec(NREC)
  real(c_double) :: sisvx(NSTEP,NREC), sisvy(NSTEP,NREC)
  real(c_double) :: total_energy(NSTEP), energy_kinetic(NSTEP), energy_potential(NSTEP)
  real(c_double), allocatable :: plane(:,:)
  real(c_double) :: Vsolidnorm, Courant_number

  type(cpml_config) :: cfg
  type(c_ptr) :: h
  integer(c_int32_t) :: ierr, it, it_begin, it_end
  character(kind=c_char, len=2) :: here = '.' // c_null_char

  h = c_null_ptr

! --- set-up phase (what the reference computes before "do it = 1,NSTEP")
  ierr = cpml_host_pml_profile(NX, DELTAX, DELTAT, NPOINTS_PML, b2i(USE_PML_XMIN), b2i(USE_PML_XMAX), cp, Rcoef, &
           NPOWER, K_MAX_PML, ALPHA_MAX_PML, 0, 1, a_x, b_x, K_x, a_x_half, b_x_half, K_x_half)
  ierr = cpml_host_pml_profile(NY, DELTAY, DELTAT, NPOINTS_PML, b2i(USE_PML_YMIN), b2i(USE_PML_YMAX), cp, Rcoef, &
           NPOWER, K_MAX_PML, ALPHA_MAX_PML, 0, 0, a_y, b_y, K_y, a_y_half, b_y_half, K_y_half)
  ierr = cpml_host_pml_profile(NZ, DELTAZ, DELTAT, NPOINTS_PML, b2i(USE_PML_ZMIN), b2i(USE_PML_ZMAX), cp, Rcoef, &
           NPOWER, K_MAX_PML, ALPHA_MAX_PML, 0, 0, a_z, b_z, K_z, a_z_half, b_z_half, K_z_half)
  ierr = cpml_host_source_series(NSTEP, DELTAT, f0, t0, factor, ANGLE_FORCE, force_x, force_y)
  ierr = cpml_host_find_receivers(NX, NY, DELTAX, DELTAY, NREC, xdeb, ydeb, xfin, yfin, ix_rec, iy_rec, dist_rec)

  Courant_number = cpml_host_courant(cp, DELTAT, DELTAX, DELTAY, DELTAZ)
  print *,'Courant number is ',Courant_number
  if (Courant_number > 1.d0) stop 'time step is too large, simulation will be unstable'

! --- hand everything to the GPU
  cfg%ndim = 3;  cfg%order = 2
  cfg%nx = NX;  cfg%ny = NY;  cfg%nz = NZ
  cfg%nstep = NSTEP;  cfg%npoints_pml = NPOINTS_PML;  cfg%nrec = NREC
  cfg%isource = ISOURCE;  cfg%jsource = JSOURCE;  cfg%ksource = 0
  cfg%nslabs = 1;  cfg%slab_rank = 0;  cfg%device = -1;  cfg%energy_bug_compat = 1
  cfg%rheology = 0
  cfg%emulate_nproc = 0
  cfg%compute_energy = 0
  cfg%sigmazz_isotropic = 0
  cfg%deltax = DELTAX;  cfg%deltay = DELTAY;  cfg%deltaz = DELTAZ;  cfg%deltat = DELTAT
  cfg%lambda = lambda;  cfg%mu = mu;  cfg%lambdaplustwomu = lambdaplustwomu;  cfg%rho = rho;  cfg%cp = cp
  cfg%reserved_d = 0.d0

  call cpml_check(cpml_create(cfg, h), h, 'cpml_create')
  call cpml_check(cpml_set_profiles(h, CPML_AXIS_X, a_x, b_x, K_x, a_x_half, b_x_half, K_x_half, NX), h, 'profiles x')
  call cpml_check(cpml_set_profiles(h, CPML_AXIS_Y, a_y, b_y, K_y, a_y_half, b_y_half, K_y_half, NY), h, 'profiles y')
  call cpml_check(cpml_set_profiles(h, CPML_AXIS_Z, a_z, b_z, K_z, a_z_half, b_z_half, K_z_half, NZ), h, 'profiles z')
  call cpml_check(cpml_set_source_series(h, force_x, force_y, NSTEP), h, 'source')
  call cpml_check(cpml_set_receivers(h, ix_rec, iy_rec, NREC), h, 'receivers')

  allocate(plane(NX,NY))

! --- time loop: the GPU runs up to the next display step, then the driver does its output
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
      call cpml_check(cpml_get_seismograms(h, sisvx, sisvy), h, 'seismograms')
      ierr = cpml_host_write_seismograms(here, sisvx, sisvy, NSTEP, NREC, DELTAT)
      call cpml_check(cpml_get_plane(h, CPML_F_VX, NZ/2, plane), h, 'plane vx')
      ierr = cpml_host_create_color_image(here, plane, NX, NY, it, ISOURCE, JSOURCE, ix_rec, iy_rec, NREC, NPOINTS_PML, &
               b2i(USE_PML_XMIN), b2i(USE_PML_XMAX), b2i(USE_PML_YMIN), b2i(USE_PML_YMAX), 1)
      call cpml_check(cpml_get_plane(h, CPML_F_VY, NZ/2, plane), h, 'plane vy')
      ierr = cpml_host_create_color_image(here, plane, NX, NY, it, ISOURCE, JSOURCE, ix_rec, iy_rec, NREC, NPOINTS_PML, &
               b2i(USE_PML_XMIN), b2i(USE_PML_XMAX), b2i(USE_PML_YMIN), b2i(USE_PML_YMAX), 2)
    endif
    it_begin = it_end + 1
  enddo

! --- final output
  call cpml_check(cpml_get_seismograms(h, sisvx, sisvy), h, 'seismograms')
  ierr = cpml_host_write_seismograms(here, sisvx, sisvy, NSTEP, NREC, DELTAT)
! Vz_file_NNN.dat: an extension -- the reference records Vx and Vy only although its plotgnu reads Vz files
  call cpml_check(cpml_get_seismograms_vz(h, sisvx), h, 'seismograms vz')
  ierr = cpml_host_write_seismograms_vz(here, sisvx, NSTEP, NREC, DELTAT, 0.d0)
  call cpml_check(cpml_get_energy(h, total_energy, energy_kinetic, energy_potential), h, 'energy')
  ierr = cpml_host_write_energy_3d('energy.dat' // c_null_char, total_energy, NSTEP, DELTAT)
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

end program seismic_CPML_3D_iso_b200
