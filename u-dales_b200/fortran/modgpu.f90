!> \file modgpu.f90
!! ISO_C_BINDING shim between the uDALES Fortran host (unchanged namelists, unchanged call
!! surface of src/program.f90:132-207) and libudales_gpu.so (include/udales_gpu.h).
!!
!! This is the reference-side binding a maintainer adds as src/modgpu.f90.  It has not been
!! compiled in the build container (no Fortran compiler there); every interface below is a 1:1
!! image of a prototype in include/udales_gpu.h and the Python ctypes binding
!! (u-dales_b200/__init__.py) exercises exactly the same symbols with the same argument shapes.
!!
!! Hook points (see INTEGRATION.md): first executable line of
!!   tstep_update (modtstep.f90:49), advection (modadvection.f90:36), subgrid (modsubgrid.f90:128),
!!   poisson (modpois.f90:419), tstep_integrate (modtstep.f90:171), halos (modboundary.f90:67),
!!   boundary (modboundary.f90:115):
!!     if (lgpu) then; call gpu_<name>(); return; end if
!! and `call gpu_init` right after `call initpois` (program.f90:89), `call gpu_exit` in exitmodules.
!! With -fdefault-real-8 (CMakeLists.txt:46) default real == real(c_double) and default integer ==
!! integer(c_int); logicals are converted explicitly.
module modgpu
  use, intrinsic :: iso_c_binding
  implicit none
  private
  public :: lgpu, gpu_init, gpu_exit, gpu_push_state, gpu_pull_state, gpu_push, gpu_pull, &
            gpu_tstep_update, gpu_advection, gpu_subgrid, gpu_poisson, gpu_tstep_integrate, &
            gpu_halos, gpu_boundary, gpu_chkdiv, gpu_rk3_step_host, gpu_ibm_init, gpu_ibmnorm, gpu_ibm_diffcorr, gpu_forces, &
            gpu_bottom, gpu_masscorr, gpu_thermo_init, gpu_thermodynamics, gpu_pull_points, gpu_add_points

  logical :: lgpu = .false.            !< namelist RUN switch (the only new option)
  type(c_ptr) :: handle = c_null_ptr

  ! field ids = enum udgpu_field
  integer(c_int), parameter :: F_U0 = 0, F_V0 = 1, F_W0 = 2, F_UM = 3, F_VM = 4, F_WM = 5, &
                               F_UP = 6, F_VP = 7, F_WP = 8, F_PRES0 = 9, F_P = 10, F_EKM = 11, &
                               F_EKH = 12, F_RHS = 13, F_SV0 = 14, F_SVM = 15, F_SVP = 16, F_MOMFLUXB = 17, &
                               F_THL0 = 18, F_THLM = 19, F_THLP = 20

  type, bind(C) :: udgpu_cfg
    integer(c_int) :: abi_version
    integer(c_int) :: itot, jtot, ktot
    integer(c_int) :: imax, jmax, kmax
    integer(c_int) :: ih, jh, kh
    integer(c_int) :: ihc, jhc, khc
    integer(c_int) :: nsv
    integer(c_int) :: zstart(3)
    integer(c_int) :: nprocx, nprocy
    integer(c_int) :: myidx, myidy
    integer(c_int) :: BCxm, BCym, BCtopm, BCzp
    integer(c_int) :: ipoiss, iadv_mom, iadv_sv
    integer(c_int) :: lles, lvreman, lsmagorinsky, loneeqn
    integer(c_int) :: ltempeq, lmoist
    real(c_double) :: dx, dy
    type(c_ptr)    :: dzf, dzh, delta
    real(c_double) :: numol, prandtlmoli, prandtli
    real(c_double) :: c_vreman, cs
    real(c_double) :: Uinf, Vinf
    real(c_double) :: e12min
    integer(c_int) :: device
    integer(c_int) :: flags
    integer(c_int) :: iadv_thl
  end type udgpu_cfg

  interface
    integer(c_int) function udgpu_nccl_unique_id(uid) bind(C, name="udgpu_nccl_unique_id")
      import :: c_int, c_char
      character(kind=c_char) :: uid(128)
    end function
    integer(c_int) function udgpu_init(cfg, uid, h) bind(C, name="udgpu_init")
      import :: c_int, c_ptr, udgpu_cfg
      type(udgpu_cfg), intent(in) :: cfg
      type(c_ptr), value :: uid
      type(c_ptr), intent(out) :: h
    end function
    integer(c_int) function udgpu_set_wfuno(h, z0h, prandtlturb, grav, thls, tcell) bind(C, name="udgpu_set_wfuno")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h
      real(c_double), value :: z0h, prandtlturb, grav, thls, tcell
    end function
    integer(c_int) function udgpu_set_thermo(h, lbuoyancy, grav, thls, BCtopT, wttop, thl_top, BCbotT, wtsurf, thlpcar) &
        bind(C, name="udgpu_set_thermo")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h
      integer(c_int), value :: lbuoyancy, BCtopT, BCbotT
      real(c_double), value :: grav, thls, wttop, thl_top, wtsurf
      real(c_double), intent(in) :: thlpcar(*)
    end function
    integer(c_int) function udgpu_set_buoycorr(h, lbuoycorr, Rigc) bind(C, name="udgpu_set_buoycorr")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h
      integer(c_int), value :: lbuoycorr
      real(c_double), value :: Rigc
    end function
    integer(c_int) function udgpu_thermodynamics(h) bind(C, name="udgpu_thermodynamics")
      import :: c_int, c_ptr
      type(c_ptr), value :: h
    end function
    integer(c_int) function udgpu_thermo_profile(h, which, host) bind(C, name="udgpu_thermo_profile")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h
      integer(c_int), value :: which
      real(c_double), intent(inout) :: host(*)
    end function
    integer(c_int) function udgpu_pull_points(h, field, n4, n, offsets, out) bind(C, name="udgpu_pull_points")
      import :: c_int, c_ptr, c_double, c_long_long
      type(c_ptr), value :: h
      integer(c_int), value :: field, n4
      integer(c_long_long), value :: n
      integer(c_long_long), intent(in) :: offsets(*)
      real(c_double), intent(inout) :: out(*)
    end function
    integer(c_int) function udgpu_add_points(h, field, n4, n, offsets, vals) bind(C, name="udgpu_add_points")
      import :: c_int, c_ptr, c_double, c_long_long
      type(c_ptr), value :: h
      integer(c_int), value :: field, n4
      integer(c_long_long), value :: n
      integer(c_long_long), intent(in) :: offsets(*)
      real(c_double), intent(in) :: vals(*)
    end function
    integer(c_int) function udgpu_finalize(h) bind(C, name="udgpu_finalize")
      import :: c_int, c_ptr
      type(c_ptr), value :: h
    end function
    function udgpu_last_error() bind(C, name="udgpu_last_error") result(msg)
      import :: c_ptr
      type(c_ptr) :: msg
    end function
    ! arrays are passed as assumed-size dummies (sequence association): um, up ... do not have the
    ! TARGET attribute in modfields (src/modfields.f90:30-32,66-68), so c_loc is not an option.
    integer(c_int) function udgpu_push(h, field, n4, host) bind(C, name="udgpu_push")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h
      integer(c_int), value :: field, n4
      real(c_double), intent(in) :: host(*)
    end function
    integer(c_int) function udgpu_pull(h, field, n4, host) bind(C, name="udgpu_pull")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h
      integer(c_int), value :: field, n4
      real(c_double), intent(inout) :: host(*)
    end function
    integer(c_int) function udgpu_sync(h) bind(C, name="udgpu_sync")
      import :: c_int, c_ptr
      type(c_ptr), value :: h
    end function
    integer(c_int) function udgpu_tstep_update(h, dt, courant, diffnr, dtmax, ladaptive, rk3step, courtot, diffnrtot) &
        bind(C, name="udgpu_tstep_update")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h
      real(c_double), intent(inout) :: dt
      real(c_double), value :: courant, diffnr, dtmax
      integer(c_int), value :: ladaptive
      integer(c_int), intent(inout) :: rk3step
      real(c_double), intent(out) :: courtot, diffnrtot
    end function
    integer(c_int) function udgpu_advection(h) bind(C, name="udgpu_advection")
      import :: c_int, c_ptr
      type(c_ptr), value :: h
    end function
    integer(c_int) function udgpu_subgrid(h) bind(C, name="udgpu_subgrid")
      import :: c_int, c_ptr
      type(c_ptr), value :: h
    end function
    integer(c_int) function udgpu_poisson(h, dt, rk3step) bind(C, name="udgpu_poisson")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h
      real(c_double), value :: dt
      integer(c_int), value :: rk3step
    end function
    integer(c_int) function udgpu_tstep_integrate(h, dt, rk3step) bind(C, name="udgpu_tstep_integrate")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h
      real(c_double), value :: dt
      integer(c_int), value :: rk3step
    end function
    integer(c_int) function udgpu_halos(h) bind(C, name="udgpu_halos")
      import :: c_int, c_ptr
      type(c_ptr), value :: h
    end function
    integer(c_int) function udgpu_boundary(h) bind(C, name="udgpu_boundary")
      import :: c_int, c_ptr
      type(c_ptr), value :: h
    end function
    integer(c_int) function udgpu_divergence(h, divmax, divtot, divrms) bind(C, name="udgpu_divergence")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h
      real(c_double), intent(out) :: divmax, divtot, divrms
    end function
    integer(c_int) function udgpu_set_forcing(h, dpdxl, dpdyl) bind(C, name="udgpu_set_forcing")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h
      real(c_double), intent(in) :: dpdxl(*), dpdyl(*)
    end function
    integer(c_int) function udgpu_forces(h) bind(C, name="udgpu_forces")
      import :: c_int, c_ptr
      type(c_ptr), value :: h
    end function
    integer(c_int) function udgpu_set_bottom(h, lbottom, BCbotm, BCbots, z0, fkar) bind(C, name="udgpu_set_bottom")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h
      integer(c_int), value :: lbottom, BCbotm, BCbots
      real(c_double), value :: z0, fkar
    end function
    integer(c_int) function udgpu_bottom(h) bind(C, name="udgpu_bottom")
      import :: c_int, c_ptr
      type(c_ptr), value :: h
    end function
    integer(c_int) function udgpu_set_masscorr(h, luvolflowr, lvvolflowr, uflowrate, vflowrate) bind(C, name="udgpu_set_masscorr")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h
      integer(c_int), value :: luvolflowr, lvvolflowr
      real(c_double), value :: uflowrate, vflowrate
    end function
    integer(c_int) function udgpu_masscorr(h, dt, rk3step, udef, vdef) bind(C, name="udgpu_masscorr")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h
      real(c_double), value :: dt
      integer(c_int), value :: rk3step
      type(c_ptr), value :: udef, vdef       ! c_null_ptr: no host synchronisation
    end function
    integer(c_int) function udgpu_ibm_set_points(h, kind, n, ijk, layout) bind(C, name="udgpu_ibm_set_points")
      import :: c_int, c_ptr
      type(c_ptr), value :: h
      integer(c_int), value :: kind, n, layout
      integer(c_int), intent(in) :: ijk(*)
    end function
    integer(c_int) function udgpu_ibm_commit(h) bind(C, name="udgpu_ibm_commit")
      import :: c_int, c_ptr
      type(c_ptr), value :: h
    end function
    integer(c_int) function udgpu_ibmnorm(h) bind(C, name="udgpu_ibmnorm")
      import :: c_int, c_ptr
      type(c_ptr), value :: h
    end function
    integer(c_int) function udgpu_ibm_diffcorr(h) bind(C, name="udgpu_ibm_diffcorr")
      import :: c_int, c_ptr
      type(c_ptr), value :: h
    end function
    integer(c_int) function udgpu_rk3_step_host(h, u0, v0, w0, pres0, dt, dtmax, ladaptive, courant, diffnr) &
        bind(C, name="udgpu_rk3_step_host")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h
      real(c_double), intent(inout) :: u0(*), v0(*), w0(*), pres0(*)
      real(c_double), intent(inout) :: dt
      real(c_double), value :: dtmax, courant, diffnr
      integer(c_int), value :: ladaptive
    end function
  end interface

contains

  !> the reference's error convention: message on unit 0, stop 1 (e.g. src/modpois.f90:896-898)
  subroutine chk(ierr, where)
    integer(c_int), intent(in) :: ierr
    character(len=*), intent(in) :: where
    character(kind=c_char), pointer :: msg(:)
    integer :: n
    if (ierr == 0) return
    call c_f_pointer(udgpu_last_error(), msg, [1024])
    n = 1
    do while (n < 1024 .and. msg(n) /= c_null_char)
      n = n + 1
    end do
    write (0, *) 'ERROR: libudales_gpu ', where, ' failed with code ', ierr, ': ', msg(1:n - 1)
    stop 1
  end subroutine chk

  integer(c_int) function l2i(l)
    logical, intent(in) :: l
    l2i = merge(1_c_int, 0_c_int, l)
  end function l2i

  !> after initpois (src/program.f90:89): hand the grid, metrics and switches to the device library
  subroutine gpu_init
    use modglobal, only: itot, jtot, ktot, imax, jmax, kmax, ih, jh, kh, ihc, jhc, khc, nsv, dx, dy, dzf, dzh, &
                         delta, BCxm, BCym, BCtopm, BCzp, ipoiss, iadv_mom, iadv_sv, iadv_thl, lles, ltempeq, lmoist, &
                         numol, prandtlmoli, Uinf, Vinf, e12min, ib, kb
    use modsubgriddata, only: lvreman, lsmagorinsky, loneeqn, prandtli, c_vreman, cs
    use modmpi, only: nprocx, nprocy, myidx, myidy, myid, comm3d, mpierr
    use decomp_2d, only: zstart
    use mpi
    type(udgpu_cfg) :: c
    character(kind=c_char), target :: uid(128)
    real(c_double), allocatable, target, save :: dzf_c(:), dzh_c(:), delta_c(:)

    allocate (dzf_c(ktot + 2*kh), dzh_c(ktot + kh), delta_c(ktot + kh))
    dzf_c = dzf(kb - kh:kb + ktot - 1 + kh)
    dzh_c = dzh(kb:kb + ktot - 1 + kh)
    delta_c = delta(ib, kb:kb + ktot - 1 + kh)

    c%abi_version = 2
    c%itot = itot; c%jtot = jtot; c%ktot = ktot
    c%imax = imax; c%jmax = jmax; c%kmax = kmax
    c%ih = ih; c%jh = jh; c%kh = kh
    c%ihc = ihc; c%jhc = jhc; c%khc = khc
    c%nsv = nsv
    c%zstart = zstart
    c%nprocx = nprocx; c%nprocy = nprocy; c%myidx = myidx; c%myidy = myidy
    c%BCxm = BCxm; c%BCym = BCym; c%BCtopm = BCtopm; c%BCzp = BCzp
    c%ipoiss = ipoiss; c%iadv_mom = iadv_mom
    c%iadv_sv = 7
    if (nsv > 0) then
      ! the device path runs one advection scheme for all scalars (the reference forces kappa for every sv anyway,
      ! src/modglobal.f90:556-559)
      if (any(iadv_sv(1:nsv) /= iadv_sv(1))) then
        write (0, *) 'ERROR: gpu_init: all scalars must use the same advection scheme, iadv_sv = ', iadv_sv(1:nsv)
        stop 1
      end if
      c%iadv_sv = iadv_sv(1)
    end if
    c%lles = l2i(lles); c%lvreman = l2i(lvreman); c%lsmagorinsky = l2i(lsmagorinsky); c%loneeqn = l2i(loneeqn)
    c%ltempeq = l2i(ltempeq); c%lmoist = l2i(lmoist)
    c%dx = dx; c%dy = dy
    c%dzf = c_loc(dzf_c); c%dzh = c_loc(dzh_c); c%delta = c_loc(delta_c)
    c%numol = numol; c%prandtlmoli = prandtlmoli; c%prandtli = prandtli
    c%c_vreman = c_vreman; c%cs = cs; c%Uinf = Uinf; c%Vinf = Vinf; c%e12min = e12min
    c%device = -1          ! LOCAL_RANK (one rank per GPU)
    c%flags = 0
    c%iadv_thl = iadv_thl

    if (nprocx*nprocy > 1) then
      if (myid == 0) call chk(udgpu_nccl_unique_id(uid), 'nccl_unique_id')
      call MPI_BCAST(uid, 128, MPI_CHARACTER, 0, comm3d, mpierr)
      call chk(udgpu_init(c, c_loc(uid), handle), 'init')
    else
      call chk(udgpu_init(c, c_null_ptr, handle), 'init')
    end if
    call gpu_push_state
    if (ltempeq) call gpu_thermo_init
  end subroutine gpu_init

  !> temperature, dry (ltempeq, lbuoyancy): namelist values and the radiative tendency profile go down once, thl0 / thlm
  !! follow, and the first thermodynamics (src/modstartup.f90 calls it before the time loop) sets thvh for forces
  subroutine gpu_thermo_init
    use modglobal, only: lbuoyancy, grav, BCtopT, BCbotT, lmoist
    use modfields, only: thl0, thlm, thlp, thlpcar
    use modglobal, only: prandtlturb, kb
    use modfields, only: thlprof
    use modsurfdata, only: thls, wtsurf, wttop, thl_top, z0h
    use modsubgriddata, only: lbuoycorr, Rigc
    if (lmoist) then
      write (0, *) 'ERROR: gpu_thermo_init: lmoist is outside the GPU path'
      stop 1
    end if
    if (BCbotT == 2) call chk(udgpu_set_wfuno(handle, z0h, prandtlturb, grav, thls, thlprof(kb)), 'set_wfuno')   ! wfuno case 92
    call chk(udgpu_set_thermo(handle, l2i(lbuoyancy), grav, thls, int(BCtopT, c_int), wttop, thl_top, int(BCbotT, c_int), &
                              wtsurf, thlpcar), 'set_thermo')
    call chk(udgpu_set_buoycorr(handle, l2i(lbuoycorr), Rigc), 'set_buoycorr')
    call chk(udgpu_push(handle, F_THL0, 0_c_int, thl0), 'push thl0')
    call chk(udgpu_push(handle, F_THLM, 0_c_int, thlm), 'push thlm')
    call chk(udgpu_push(handle, F_THLP, 0_c_int, thlp), 'push thlp')
    call chk(udgpu_thermodynamics(handle), 'thermodynamics')
  end subroutine gpu_thermo_init
  !> thermodynamics (src/modthermodynamics.f90:55, program.f90:212), dry: thl0av, thvh on the device; the host copies
  !! of the two profiles are refreshed for the statistics
  subroutine gpu_thermodynamics
    use modglobal, only: ltempeq
    use modfields, only: thl0av, thvh
    if (.not. ltempeq) return
    call chk(udgpu_thermodynamics(handle), 'thermodynamics')
    call chk(udgpu_thermo_profile(handle, 0_c_int, thl0av), 'thl0av')
    call chk(udgpu_thermo_profile(handle, 1_c_int, thvh), 'thvh')
  end subroutine gpu_thermodynamics

  subroutine gpu_exit
    if (c_associated(handle)) call chk(udgpu_finalize(handle), 'finalize')
    handle = c_null_ptr
  end subroutine gpu_exit

  !> residency control: the prognostic state the hot path reads (after readinitfiles / a host add-on)
  subroutine gpu_push_state
    use modfields, only: u0, v0, w0, um, vm, wm, pres0, up, vp, wp
    call chk(udgpu_push(handle, F_U0, 0_c_int, u0), 'push u0')
    call chk(udgpu_push(handle, F_V0, 0_c_int, v0), 'push v0')
    call chk(udgpu_push(handle, F_W0, 0_c_int, w0), 'push w0')
    call chk(udgpu_push(handle, F_UM, 0_c_int, um), 'push um')
    call chk(udgpu_push(handle, F_VM, 0_c_int, vm), 'push vm')
    call chk(udgpu_push(handle, F_WM, 0_c_int, wm), 'push wm')
    call chk(udgpu_push(handle, F_PRES0, 0_c_int, pres0), 'push pres0')
    call chk(udgpu_push(handle, F_UP, 0_c_int, up), 'push up')
    call chk(udgpu_push(handle, F_VP, 0_c_int, vp), 'push vp')
    call chk(udgpu_push(handle, F_WP, 0_c_int, wp), 'push wp')
    call chk(udgpu_sync(handle), 'sync')
  end subroutine gpu_push_state

  !> before writerestartfiles / fielddump / statistics (time-gated in program.f90:201-220)
  subroutine gpu_pull_state
    use modglobal, only: ltempeq
    use modfields, only: u0, v0, w0, um, vm, wm, pres0, thl0, thlm
    use modsubgriddata, only: ekm, ekh
    call chk(udgpu_pull(handle, F_U0, 0_c_int, u0), 'pull u0')
    call chk(udgpu_pull(handle, F_V0, 0_c_int, v0), 'pull v0')
    call chk(udgpu_pull(handle, F_W0, 0_c_int, w0), 'pull w0')
    call chk(udgpu_pull(handle, F_UM, 0_c_int, um), 'pull um')
    call chk(udgpu_pull(handle, F_VM, 0_c_int, vm), 'pull vm')
    call chk(udgpu_pull(handle, F_WM, 0_c_int, wm), 'pull wm')
    call chk(udgpu_pull(handle, F_PRES0, 0_c_int, pres0), 'pull pres0')
    call chk(udgpu_pull(handle, F_EKM, 0_c_int, ekm), 'pull ekm')
    call chk(udgpu_pull(handle, F_EKH, 0_c_int, ekh), 'pull ekh')
    if (ltempeq) then
      call chk(udgpu_pull(handle, F_THL0, 0_c_int, thl0), 'pull thl0')
      call chk(udgpu_pull(handle, F_THLM, 0_c_int, thlm), 'pull thlm')
    end if
  end subroutine gpu_pull_state

  !> single-field variants for host add-ons that touch the tendencies between subgrid and poisson
  !! (program.f90:152-191): call gpu_pull(F_UP, up) ... host routine ... call gpu_push(F_UP, up)
  subroutine gpu_push(field, a)
    integer(c_int), intent(in) :: field
    real(c_double), intent(in) :: a(*)
    call chk(udgpu_push(handle, field, 0_c_int, a), 'push')
  end subroutine gpu_push
  subroutine gpu_pull(field, a)
    integer(c_int), intent(in) :: field
    real(c_double), intent(inout) :: a(*)
    call chk(udgpu_pull(handle, field, 0_c_int, a), 'pull')
  end subroutine gpu_pull

  !> sparse residency for host add-ons that touch few cells (the facet wall functions): values of a resident field at
  !! the points (i,j,k) of a list, and additions to a resident tendency at such points.  ijk(n,3): Fortran indices of the
  !! reference arrays (halo cells allowed); klo = lower k bound of the array (kb-kh for fields, kb for tendencies)
  subroutine gpu_pull_points(field, n, ijk, klo, vals)
    use modglobal, only: ib, jb, ih, jh, imax, jmax
    integer(c_int), intent(in) :: field
    integer, intent(in) :: n, ijk(n, 3), klo
    real(c_double), intent(out) :: vals(n)
    integer(c_long_long) :: off(max(n, 1))
    integer :: q
    do q = 1, n
      off(q) = int(ijk(q, 1) - (ib - ih), c_long_long) + int(imax + 2*ih, c_long_long)*(int(ijk(q, 2) - (jb - jh), c_long_long) &
               + int(jmax + 2*jh, c_long_long)*int(ijk(q, 3) - klo, c_long_long))
    end do
    call chk(udgpu_pull_points(handle, field, 0_c_int, int(n, c_long_long), off, vals), 'pull_points')
  end subroutine gpu_pull_points
  subroutine gpu_add_points(field, n, ijk, klo, vals)
    use modglobal, only: ib, jb, ih, jh, imax, jmax
    integer(c_int), intent(in) :: field
    integer, intent(in) :: n, ijk(n, 3), klo
    real(c_double), intent(in) :: vals(n)
    integer(c_long_long) :: off(max(n, 1))
    integer :: q
    do q = 1, n
      off(q) = int(ijk(q, 1) - (ib - ih), c_long_long) + int(imax + 2*ih, c_long_long)*(int(ijk(q, 2) - (jb - jh), c_long_long) &
               + int(jmax + 2*jh, c_long_long)*int(ijk(q, 3) - klo, c_long_long))
    end do
    call chk(udgpu_add_points(handle, field, 0_c_int, int(n, c_long_long), off, vals), 'add_points')
  end subroutine gpu_add_points

  subroutine gpu_tstep_update
    use modglobal, only: dt, courant, diffnr, dtmax, ladaptive, rk3step, timee, timeleft, ntimee, ntrun, dt_lim
    real(c_double) :: ct, dn
    integer(c_int) :: rk
    rk = rk3step
    call chk(udgpu_tstep_update(handle, dt, courant, diffnr, dtmax, l2i(ladaptive), rk, ct, dn), 'tstep_update')
    rk3step = rk
    if (rk3step == 1) then          ! bookkeeping of src/modtstep.f90:136-147 stays on the host
      if (ladaptive) dt_lim = timeleft     ! :136 (before timeleft is advanced; the non-adaptive branch :142-147 leaves dt_lim alone)
      timeleft = timeleft - dt
      timee = timee + dt
      ntimee = ntimee + 1
      ntrun = ntrun + 1
    end if
  end subroutine gpu_tstep_update

  subroutine gpu_advection
    call chk(udgpu_advection(handle), 'advection')
  end subroutine
  subroutine gpu_subgrid
    call chk(udgpu_subgrid(handle), 'subgrid')
  end subroutine
  subroutine gpu_poisson
    use modglobal, only: dt, rk3step
    call chk(udgpu_poisson(handle, dt, int(rk3step, c_int)), 'poisson')
  end subroutine
  subroutine gpu_tstep_integrate
    use modglobal, only: dt, rk3step
    call chk(udgpu_tstep_integrate(handle, dt, int(rk3step, c_int)), 'tstep_integrate')
  end subroutine
  subroutine gpu_halos
    call chk(udgpu_halos(handle), 'halos')
  end subroutine
  subroutine gpu_boundary
    call chk(udgpu_boundary(handle), 'boundary')
  end subroutine
  !> chkdiv (src/modchecksim.f90:161): divmax, divtot from the resident fields
  subroutine gpu_chkdiv(divmax, divtot)
    real(c_double), intent(out) :: divmax, divtot
    real(c_double) :: divrms
    call chk(udgpu_divergence(handle, divmax, divtot, divrms), 'divergence')
  end subroutine
  !> forces (src/modforces.f90:46, neutral branch) on the resident tendencies; dpdxl/dpdyl (kb:ke+kh) are uploaded on
  !! every call (ktot+1 doubles each: negligible) so that fixuinf / time-dependent forcing on the host is picked up
  subroutine gpu_forces
    use modfields, only: dpdxl, dpdyl
    ! lbuoyancy: the buoyancy term and thlpcar are applied by the library (udgpu_set_thermo, gpu_thermo_init)
    call chk(udgpu_set_forcing(handle, dpdxl, dpdyl), 'set_forcing')
    call chk(udgpu_forces(handle), 'forces')
  end subroutine
  !> hand modibm's local point lists to the device (call once after initibm, src/program.f90); the lists are the
  !! (n,3) integer arrays solid_info_*%solpts_loc / bound_info_*%bndpts_loc exactly as they lie in memory (layout 1)
  subroutine gpu_ibm_init
    use modglobal, only: libm, nsv, ltempeq
    use modibm, only: solid_info_u, solid_info_v, solid_info_w, solid_info_c, &
                      bound_info_u, bound_info_v, bound_info_w, bound_info_c
    integer(c_int) :: dummy(3)
    if (.not. libm) return
    dummy = 1
    ! a rank may own no point of a kind: its list is then zero-sized or not allocated at all, so a dummy goes in its place
    call set_list(0_c_int, solid_info_u%nsolptsrank, solid_info_u%solpts_loc)
    call set_list(1_c_int, solid_info_v%nsolptsrank, solid_info_v%solpts_loc)
    call set_list(2_c_int, solid_info_w%nsolptsrank, solid_info_w%solpts_loc)
    call set_list(4_c_int, bound_info_u%nbndptsrank, bound_info_u%bndpts_loc)
    call set_list(5_c_int, bound_info_v%nbndptsrank, bound_info_v%bndpts_loc)
    call set_list(6_c_int, bound_info_w%nbndptsrank, bound_info_w%bndpts_loc)
    if (nsv > 0 .or. ltempeq) then
      call set_list(3_c_int, solid_info_c%nsolptsrank, solid_info_c%solpts_loc)
      call set_list(7_c_int, bound_info_c%nbndptsrank, bound_info_c%bndpts_loc)
    else
      call chk(udgpu_ibm_set_points(handle, 3_c_int, 0_c_int, dummy, 1_c_int), 'ibm solid_c')
      call chk(udgpu_ibm_set_points(handle, 7_c_int, 0_c_int, dummy, 1_c_int), 'ibm bound_c')
    end if
    call chk(udgpu_ibm_commit(handle), 'ibm commit')
  contains
    subroutine set_list(kind, n, pts)
      integer(c_int), intent(in) :: kind
      integer, intent(in) :: n
      integer, allocatable, intent(in) :: pts(:, :)
      if (n > 0 .and. allocated(pts)) then
        call chk(udgpu_ibm_set_points(handle, kind, int(n, c_int), pts, 1_c_int), 'ibm point list')
      else
        call chk(udgpu_ibm_set_points(handle, kind, 0_c_int, dummy, 1_c_int), 'ibm point list (empty)')
      end if
    end subroutine set_list
  end subroutine
  !> bottom (src/modibm.f90:1998, program.f90:152): wfuno case 91 (BCbotm = 2, the default) or wfmneutral case 91
  !! (BCbotm = 3), the temperature bottom (BCbotT = 1 flux / 2 wfuno case 92) and the zero-flux scalar bottom on the resident
  !! tendencies.  The namelist values go down once (first call)
  subroutine gpu_bottom
    use modglobal, only: lbottom, BCbotm, BCbotT, BCbots, fkar, grav, prandtlturb, kb
    use modsurfdata, only: z0, z0h, thls
    use modfields, only: thlprof
    logical, save :: first = .true.
    if (first) then
      ! wfuno's stability correction: wall temperature thls; without temperature equation thl0 stays at thlprof
      if (BCbotm == 2 .or. BCbotT == 2) call chk(udgpu_set_wfuno(handle, z0h, prandtlturb, grav, thls, thlprof(kb)), 'set_wfuno')
      call chk(udgpu_set_bottom(handle, l2i(lbottom), int(BCbotm, c_int), int(BCbots, c_int), z0, fkar), 'set_bottom')
      first = .false.
    end if
    call chk(udgpu_bottom(handle), 'bottom')
  end subroutine
  !> masscorr (src/modforces.f90:328, program.f90:169), volume-flow branches; the outflow-rate branches (luoutflowr /
  !! lvoutflowr) are for non-periodic domains and stay with the host.  udef / vdef are not needed by the host path.
  subroutine gpu_masscorr
    use modglobal, only: dt, rk3step, linoutflow, luoutflowr, lvoutflowr, luvolflowr, lvvolflowr, uflowrate, vflowrate
    logical, save :: first = .true.
    if (linoutflow) return
    if (luoutflowr .or. lvoutflowr) then
      write (0, *) 'ERROR: gpu_masscorr: luoutflowr / lvoutflowr are outside the GPU path (periodic domains use l[uv]volflowr)'
      stop 1
    end if
    if (first) then
      call chk(udgpu_set_masscorr(handle, l2i(luvolflowr), l2i(lvvolflowr), uflowrate, vflowrate), 'set_masscorr')
      first = .false.
    end if
    call chk(udgpu_masscorr(handle, dt, int(rk3step, c_int), c_null_ptr, c_null_ptr), 'masscorr')
  end subroutine
  !> ibmnorm (src/modibm.f90:697) and the diff*_corr part of ibmwallfun (:1211-1213,1240-1242) on the resident fields
  subroutine gpu_ibmnorm
    call chk(udgpu_ibmnorm(handle), 'ibmnorm')
  end subroutine
  subroutine gpu_ibm_diffcorr
    call chk(udgpu_ibm_diffcorr(handle), 'ibm_diffcorr')
  end subroutine
  !> one whole RK3 time step (three passes of program.f90:132-207) on the host-resident module arrays:
  !! the drop-in for a model whose other physics stay on the CPU and touch the fields once per time step
  subroutine gpu_rk3_step_host
    use modglobal, only: dt, dtmax, ladaptive, courant, diffnr
    use modfields, only: u0, v0, w0, pres0
    integer(c_int) :: lad
    lad = 0; if (ladaptive) lad = 1
    call chk(udgpu_rk3_step_host(handle, u0, v0, w0, pres0, dt, dtmax, lad, courant, diffnr), 'rk3_step_host')
  end subroutine
end module modgpu
