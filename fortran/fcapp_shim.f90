!***********************************************************************
! fcapp_shim.f90 -- ISO_C_BINDING layer between freeCappuccino's Fortran
! host code and libfcapp_cuda.so (C ABI: include/fcapp.h).
!
! Drop-in use: remove  sparse_matrix.f90's create_CSR_matrix_from_mesh_data
! body, calcp-multiple_correction_SIMPLE.f90, calcuvw.f90, PISO_multiple_correction.f90,
! PIMPLE_multiple_correction.f90, get_rAU_x_UEqnH.f90, dpcg.f90, iccg.f90,
! bicgstab.f90, fvm_laplacian.f90 (laplacian) and the 'gauss' branch of
! gradients.f90 from the Makefile's object list, add this file, and link
! with  -L<repo>/freecappuccino_b200 -lfcapp_cuda .  Every subroutine
! below keeps the reference's name and argument list, so no call site in
! main.f90 / calcuvw.f90 / init.f90 / poisson.f90 changes.
!
! Not compiled in this repository's image (it has no Fortran compiler);
! the same calls are exercised through ctypes (freecappuccino_b200/lib.py)
! and C++ (host/), which pass arguments exactly as gfortran does here:
! default INTEGER = c_int, REAL(dp) = c_double, arrays by reference,
! 1-based index values.
!***********************************************************************
module fcapp_c
  use iso_c_binding
  implicit none

  integer(c_int), parameter :: FC_DPCG = 0, FC_ICCG = 1, FC_BICGSTAB = 2
  ! field ids of include/fcapp.h
  integer(c_int), parameter :: FC_U=0, FC_V=1, FC_W=2, FC_P=3, FC_PP=4, FC_DEN=5, FC_FLMASS=6,      &
                               FC_APU=7, FC_APV=8, FC_APW=9, FC_DUDXI=10, FC_DVDXI=11, FC_DWDXI=12, &
                               FC_DPDXI=13, FC_A=14, FC_SU=15, FC_RES=16, FC_FMI=17, FC_FMO=18,      &
                               FC_APR=19, FC_FMPRO=20, FC_SCRATCH_T=21, FC_USER0=22,                  &
                               FC_VIS=26, FC_UO=27, FC_VO=28, FC_WO=29, FC_UOO=30, FC_VOO=31, FC_WOO=32, &
                               FC_T=33, FC_SV=34, FC_SW=35, FC_SPU=36, FC_SPV=37, FC_SP=38

  type, bind(C) :: fc_mesh_desc
    integer(c_int) :: numCells, numInnerFaces, numFaces, numTotal
    integer(c_int) :: npro
    integer(c_int) :: ninl, nout, nsym, nwal, npru, noc
    integer(c_int) :: iProcFacesStart, iInletFacesStart, iOutletFacesStart, iSymmetryFacesStart, &
                      iWallFacesStart, iPressOutletFacesStart, iOCFacesStart
    type(c_ptr) :: owner, neighbour
    type(c_ptr) :: xc, yc, zc, vol
    type(c_ptr) :: arx, ary, arz, xf, yf, zf
    type(c_ptr) :: facint, fpro
    integer(c_int) :: numConnections
    type(c_ptr) :: neighbProcNo, neighbProcOffset
    integer(c_int) :: gloCells
  end type

  type, bind(C) :: fc_solver_opts
    real(c_double) :: sor
    integer(c_int) :: nsw
    real(c_double) :: small
    real(c_double) :: tol
    integer(c_int) :: parallel
  end type

  type, bind(C) :: fc_solver_report
    real(c_double) :: res0, resl
    integer(c_int) :: iters
  end type

  type, bind(C) :: fc_calcp_opts
    integer(c_int) :: npcor, nigrad, nipgrad, pRefCell
    real(c_double) :: urf_p
    integer(c_int) :: solver, const_mflux
    real(c_double) :: flomas
    integer(c_int) :: lsq_flag, flux_variant
    type(fc_solver_opts) :: sol
  end type

  type, bind(C) :: fc_calcp_report
    type(fc_solver_report) :: rep(8)
    real(c_double) :: sumLocalContErr, globalContErr
  end type

  type, bind(C) :: fc_calcuvw_opts
    integer(c_int) :: nigrad, nipgrad, scheme, limiter
    real(c_double) :: gds
    real(c_double) :: urf(3), sor(3)
    integer(c_int) :: nsw(3)
    integer(c_int) :: bdf
    real(c_double) :: btime, timestep
    integer(c_int) :: cn, const_mflux
    real(c_double) :: gradPcmf
    integer(c_int) :: lbuoy, boussinesq
    real(c_double) :: beta, tref, densit, gravx, gravy, gravz, viscos
    type(fc_solver_opts) :: sol
  end type

  type, bind(C) :: fc_piso_opts
    integer(c_int) :: ncorr, npcor, nigrad, nipgrad, pRefCell, pimple
    real(c_double) :: urf_p
    integer(c_int) :: const_mflux
    real(c_double) :: flomas
    integer(c_int) :: bdf
    real(c_double) :: btime, timestep
    integer(c_int) :: cn, lbuoy, boussinesq
    real(c_double) :: beta, tref, densit, gravx, gravy, gravz
    type(fc_solver_opts) :: sol
  end type

  type, bind(C) :: fc_piso_report
    type(fc_solver_report) :: rep(16)
    integer(c_int) :: nsolves
    real(c_double) :: sumLocalContErr, globalContErr
  end type

  type, bind(C) :: fc_calcuvw_report
    type(fc_solver_report) :: rep(3)
  end type

  type(c_ptr), save :: fc_ctx = c_null_ptr   ! one context per rank / GPU

  interface
    integer(c_int) function fc_create(device, ctx) bind(C, name='fc_create')
      import; integer(c_int), value :: device; type(c_ptr) :: ctx
    end function
    integer(c_int) function fc_destroy(ctx) bind(C, name='fc_destroy')
      import; type(c_ptr), value :: ctx
    end function
    integer(c_int) function fc_comm_unique_id(id) bind(C, name='fc_comm_unique_id')
      import; character(kind=c_char) :: id(128)
    end function
    integer(c_int) function fc_comm_init(ctx, rank, nranks, id) bind(C, name='fc_comm_init')
      import; type(c_ptr), value :: ctx; integer(c_int), value :: rank, nranks; character(kind=c_char) :: id(128)
    end function
    integer(c_int) function fc_set_mesh(ctx, mesh) bind(C, name='fc_set_mesh')
      import; type(c_ptr), value :: ctx; type(fc_mesh_desc) :: mesh
    end function
    integer(c_int) function fc_create_csr(ctx, ioffset, ja, diag, icj, jci) bind(C, name='fc_create_csr')
      import; type(c_ptr), value :: ctx; integer(c_int) :: ioffset(*), ja(*), diag(*), icj(*), jci(*)
    end function
    integer(c_int) function fc_upload(ctx, field, host, n) bind(C, name='fc_upload')
      import; type(c_ptr), value :: ctx; integer(c_int), value :: field; real(c_double) :: host(*)
      integer(c_size_t), value :: n
    end function
    integer(c_int) function fc_download(ctx, field, host, n) bind(C, name='fc_download')
      import; type(c_ptr), value :: ctx; integer(c_int), value :: field; real(c_double) :: host(*)
      integer(c_size_t), value :: n
    end function
    integer(c_int) function fc_grad_gauss(ctx, phi_field, grad_field, nigrad) bind(C, name='fc_grad_gauss')
      import; type(c_ptr), value :: ctx; integer(c_int), value :: phi_field, grad_field, nigrad
    end function
    integer(c_int) function fc_grad_gauss_corrected(ctx, phi_field, grad_field, zero_seed) &
        bind(C, name='fc_grad_gauss_corrected')
      import; type(c_ptr), value :: ctx; integer(c_int), value :: phi_field, grad_field, zero_seed
    end function
    integer(c_int) function fc_laplacian(ctx, mu_field, phi_field) bind(C, name='fc_laplacian')
      import; type(c_ptr), value :: ctx; integer(c_int), value :: mu_field, phi_field
    end function
    integer(c_int) function fc_solve_host(ctx, solver, a, su, fi, res, o, rep) bind(C, name='fc_solve_host')
      import; type(c_ptr), value :: ctx; integer(c_int), value :: solver
      real(c_double) :: a(*), su(*), fi(*), res(*); type(fc_solver_opts) :: o; type(fc_solver_report) :: rep
    end function
    integer(c_int) function fc_solve(ctx, solver, fi_field, o, rep) bind(C, name='fc_solve')
      import; type(c_ptr), value :: ctx; integer(c_int), value :: solver, fi_field
      type(fc_solver_opts) :: o; type(fc_solver_report) :: rep
    end function
    integer(c_int) function fc_calcp_host(ctx, o, u, v, w, p, pp, apu, apv, apw, flmass, rep) &
        bind(C, name='fc_calcp_host')
      import; type(c_ptr), value :: ctx; type(fc_calcp_opts) :: o
      real(c_double) :: u(*), v(*), w(*), p(*), pp(*), apu(*), apv(*), apw(*), flmass(*)
      type(fc_calcp_report) :: rep
    end function
    integer(c_int) function fc_calcuvw_host(ctx, o, u, v, w, p, vis, flmass, apu, apv, apw, rep) &
        bind(C, name='fc_calcuvw_host')
      import; type(c_ptr), value :: ctx; type(fc_calcuvw_opts) :: o
      real(c_double) :: u(*), v(*), w(*), p(*), vis(*), flmass(*), apu(*), apv(*), apw(*)
      type(fc_calcuvw_report) :: rep
    end function
    ! the two halves of calcp around the solve, for a host that keeps its own linear solver (LIS solve_csr)
    integer(c_int) function fc_calcp_assemble(ctx, o) bind(C, name='fc_calcp_assemble')
      import; type(c_ptr), value :: ctx; type(fc_calcp_opts) :: o
    end function
    integer(c_int) function fc_calcp_correct(ctx, o, ipcorr, rep) bind(C, name='fc_calcp_correct')
      import; type(c_ptr), value :: ctx; type(fc_calcp_opts) :: o; integer(c_int), value :: ipcorr
      type(fc_calcp_report) :: rep
    end function
    integer(c_int) function fc_piso(ctx, o, rep) bind(C, name='fc_piso')
      import; type(c_ptr), value :: ctx; type(fc_piso_opts) :: o; type(fc_piso_report) :: rep
    end function
    integer(c_int) function fc_exchange(ctx, field) bind(C, name='fc_exchange')
      import; type(c_ptr), value :: ctx; integer(c_int), value :: field
    end function
    integer(c_int) function fc_global_sum(ctx, x) bind(C, name='fc_global_sum')
      import; type(c_ptr), value :: ctx; real(c_double) :: x
    end function
    function fc_last_error(ctx) bind(C, name='fc_last_error') result(msg)
      import; type(c_ptr), value :: ctx; type(c_ptr) :: msg
    end function
  end interface

contains

  subroutine fc_check(ierr, who)   ! the reference has no status codes: any failure stops the run
    integer(c_int), intent(in) :: ierr
    character(len=*), intent(in) :: who
    if (ierr /= 0) then
      write(*,'(3a,i0)') '  libfcapp_cuda: ', who, ' failed with status ', ierr
      stop
    end if
  end subroutine

  function solver_opts(ifi) result(o)   ! sor(ifi), nsw(ifi), small of module parameters
    use parameters
    integer, intent(in) :: ifi
    type(fc_solver_opts) :: o
    o%sor = sor(ifi); o%nsw = nsw(ifi); o%small = small; o%tol = 1e-13; o%parallel = 0
  end function

  ! Called once after mesh_geometry (main.f90:56-62): hands the arrays of module geometry to the GPU.
  subroutine fcapp_init(device)
    use geometry
    integer, intent(in) :: device
    type(fc_mesh_desc) :: m
    call fc_check(fc_create(int(device, c_int), fc_ctx), 'fc_create')
    m%numCells = numCells; m%numInnerFaces = numInnerFaces; m%numFaces = numFaces; m%numTotal = numTotal
    m%npro = 0; m%ninl = ninl; m%nout = nout; m%nsym = nsym; m%nwal = nwal; m%npru = npru; m%noc = noc
    m%iProcFacesStart = 0; m%iInletFacesStart = iInletFacesStart; m%iOutletFacesStart = iOutletFacesStart
    m%iSymmetryFacesStart = iSymmetryFacesStart; m%iWallFacesStart = iWallFacesStart
    m%iPressOutletFacesStart = iPressOutletFacesStart; m%iOCFacesStart = iOCFacesStart
    m%owner = c_loc(owner); m%neighbour = c_loc(neighbour)
    m%xc = c_loc(xc); m%yc = c_loc(yc); m%zc = c_loc(zc); m%vol = c_loc(vol)
    m%arx = c_loc(arx); m%ary = c_loc(ary); m%arz = c_loc(arz)
    m%xf = c_loc(xf); m%yf = c_loc(yf); m%zf = c_loc(zf); m%facint = c_loc(facint)
    m%fpro = c_null_ptr; m%numConnections = 0; m%neighbProcNo = c_null_ptr; m%neighbProcOffset = c_null_ptr
    m%gloCells = numCells
    call fc_check(fc_set_mesh(fc_ctx, m), 'fc_set_mesh')
  end subroutine

end module fcapp_c

!***********************************************************************
! Replacement bodies: same names and dummy arguments as the reference.
!***********************************************************************

! src/sparse_matrix.f90:42-172 -- keeps the allocations, replaces the sort / search by the device build
subroutine create_CSR_matrix_from_mesh_data
  use fcapp_c
  use geometry, only: nnz, numCells, numInnerFaces, noc
  use sparse_matrix
  implicit none
  allocate( ioffset(numCells+1), ja(nnz), diag(numCells), a(nnz) )
  allocate( su(numCells), sv(numCells), sw(numCells), spu(numCells), spv(numCells), sp(numCells) )
  allocate( res(numCells), apu(numCells), apv(numCells), apw(numCells), al(noc), ar(noc) )
  allocate( icell_jcell_csr_value_index(numInnerFaces), jcell_icell_csr_value_index(numInnerFaces) )
  call fc_check(fc_create_csr(fc_ctx, ioffset, ja, diag, icell_jcell_csr_value_index, &
                              jcell_icell_csr_value_index), 'fc_create_csr')
end subroutine

! src/dpcg.f90:3 -- matrix and rhs are the module arrays a, su (uploaded inside the call)
subroutine dpcg(fi,ifi)
  use types
  use parameters
  use geometry, only: numTotal
  use sparse_matrix
  use title_mod
  use fcapp_c
  implicit none
  integer, intent(in) :: ifi
  real(dp), dimension(numTotal), intent(inout) :: fi
  type(fc_solver_report) :: rep
  type(fc_solver_opts) :: o
  o = solver_opts(ifi)
  call fc_check(fc_solve_host(fc_ctx, FC_DPCG, a, su, fi, res, o, rep), 'fc_solve_host(dpcg)')
  if (rep%iters > 0) resor(ifi) = rep%res0                       ! dpcg.f90:139
  write(6,'(3a,1PE10.3,a,1PE10.3,a,I0)') 'PCG(Jacobi):  Solving for ',trim(chvarSolver(ifi)), &
  ', Initial residual = ',rep%res0,', Final residual = ',rep%resl,', No Iterations ',rep%iters
end subroutine

! src/iccg.f90:3
subroutine iccg(fi,ifi)
  use types
  use parameters
  use geometry, only: numTotal
  use sparse_matrix
  use title_mod
  use fcapp_c
  implicit none
  integer, intent(in) :: ifi
  real(dp), dimension(numTotal), intent(inout) :: fi
  type(fc_solver_report) :: rep
  type(fc_solver_opts) :: o
  o = solver_opts(ifi)
  call fc_check(fc_solve_host(fc_ctx, FC_ICCG, a, su, fi, res, o, rep), 'fc_solve_host(iccg)')
  if (rep%iters > 0) resor(ifi) = rep%res0
  write(6,'(3a,1PE10.3,a,1PE10.3,a,I0)') '  PCG(IC0):  Solving for ',trim(chvarSolver(ifi)), &
  ', Initial residual = ',rep%res0,', Final residual = ',rep%resl,', No Iterations ',rep%iters
end subroutine

! src/bicgstab.f90:1
subroutine bicgstab(fi,ifi)
  use types
  use parameters
  use geometry, only: numTotal
  use sparse_matrix
  use title_mod
  use fcapp_c
  implicit none
  integer, intent(in) :: ifi
  real(dp), dimension(numTotal), intent(inout) :: fi
  type(fc_solver_report) :: rep
  type(fc_solver_opts) :: o
  o = solver_opts(ifi)
  call fc_check(fc_solve_host(fc_ctx, FC_BICGSTAB, a, su, fi, res, o, rep), 'fc_solve_host(bicgstab)')
  if (rep%iters > 0) resor(ifi) = rep%res0
  write(6,'(3a,1PE10.3,a,1PE10.3,a,I0)') '  BiCGStab(ILU(0)):  Solving for ',trim(chvarSolver(ifi)), &
  ', Initial residual = ',rep%res0,', Final residual = ',rep%resl,', No Iterations ',rep%iters
end subroutine

! src/fvm_laplacian.f90:1 -- fills the module arrays a and su (wall part) through the device
subroutine laplacian(mu,phi)
  use types
  use geometry, only: numCells, numTotal, nnz
  use sparse_matrix
  use fcapp_c
  implicit none
  real(dp), dimension(numCells), intent(in) :: mu
  real(dp), dimension(numTotal), intent(in) :: phi
  call fc_check(fc_upload(fc_ctx, FC_APU, mu, int(numCells, c_size_t)), 'upload mu')
  call fc_check(fc_upload(fc_ctx, FC_SCRATCH_T, phi, int(numTotal, c_size_t)), 'upload phi')
  call fc_check(fc_upload(fc_ctx, FC_SU, su, int(numCells, c_size_t)), 'upload su')
  call fc_check(fc_laplacian(fc_ctx, FC_APU, FC_SCRATCH_T), 'fc_laplacian')
  call fc_check(fc_download(fc_ctx, FC_A, a, int(nnz, c_size_t)), 'download a')
  call fc_check(fc_download(fc_ctx, FC_SU, su, int(numCells, c_size_t)), 'download su')
end subroutine

! src/gradients.f90:95 (grad_scalar_field, Gauss branch): dPhidxi(3,numCells) is xyz-interleaved = the device layout
subroutine grad_gauss_device(phi, dPhidxi)
  use types
  use parameters, only: nigrad
  use geometry, only: numCells, numTotal
  use fcapp_c
  implicit none
  real(dp), dimension(numTotal), intent(in) :: phi
  real(dp), dimension(3,numCells), intent(inout) :: dPhidxi
  call fc_check(fc_upload(fc_ctx, FC_SCRATCH_T, phi, int(numTotal, c_size_t)), 'upload phi')
  call fc_check(fc_grad_gauss(fc_ctx, FC_SCRATCH_T, FC_DPDXI, int(nigrad, c_int)), 'fc_grad_gauss')
  call fc_check(fc_download(fc_ctx, FC_DPDXI, dPhidxi, int(3*numCells, c_size_t)), 'download gradient')
end subroutine

! src/calcp-multiple_correction_SIMPLE.f90:3 -- `call calcp`, all state in the modules
subroutine calcp
  use types
  use parameters
  use geometry
  use sparse_matrix
  use variables
  use gradients, only: lstsq_qr, lstsq_dm
  use title_mod
  use fcapp_c
  implicit none
  type(fc_calcp_opts) :: o
  type(fc_calcp_report) :: rep
  integer :: k
  o%npcor = npcor; o%nigrad = nigrad; o%nipgrad = nipgrad; o%pRefCell = pRefCell
  o%urf_p = urf(ip); o%solver = FC_ICCG                           ! calcp :119
  o%const_mflux = merge(1, 0, const_mflux); o%flomas = flomas
  o%lsq_flag = merge(1, 0, lstsq_qr .or. lstsq_dm); o%flux_variant = 0
  o%sol = solver_opts(ip)
  ! den, fmi and the incoming pressure gradient only change outside calcp
  call fc_check(fc_upload(fc_ctx, FC_DEN, den, int(numTotal, c_size_t)), 'upload den')
  if (ninl > 0) call fc_check(fc_upload(fc_ctx, FC_FMI, fmi, int(ninl, c_size_t)), 'upload fmi')
  call fc_check(fc_upload(fc_ctx, FC_DPDXI, dPdxi, int(3*numCells, c_size_t)), 'upload dPdxi')
  call fc_check(fc_calcp_host(fc_ctx, o, u, v, w, p, pp, apu, apv, apw, flmass, rep), 'fc_calcp_host')
  call fc_check(fc_download(fc_ctx, FC_DPDXI, dPdxi, int(3*numCells, c_size_t)), 'download dPdxi')
  call fc_check(fc_download(fc_ctx, FC_SU, su, int(numCells, c_size_t)), 'download su')
  do k = 1, npcor
    if (rep%rep(k)%iters > 0) resor(ip) = rep%rep(k)%res0
    write(6,'(3a,1PE10.3,a,1PE10.3,a,I0)') '  PCG(IC0):  Solving for ',trim(chvarSolver(ip)), &
    ', Initial residual = ',rep%rep(k)%res0,', Final residual = ',rep%rep(k)%resl,', No Iterations ',rep%rep(k)%iters
  end do
  sumLocalContErr = rep%sumLocalContErr                            ! continuityErrors.h
  globalContErr = rep%globalContErr
  cumulativeContErr = cumulativeContErr + globalContErr
  write(6,'(3(a,es10.3))') "  time step continuity errors : sum local = ", sumLocalContErr, &
 &                          ", global = ", globalContErr, ", cumulative = ", cumulativeContErr
end subroutine

! src/calcuvw.f90:3 -- `call calcuvw` (laminar, serial): the momentum predictor on the device.  With calcp
! above also on the device, u, v, w, p, ap*, flmass and the gradients could stay resident between the two
! calls (fc_calcuvw + fc_calcp instead of the *_host forms); the host forms keep every module array current
! for the routines that are still Fortran (turbulence, scalars, output).
subroutine calcuvw
  use types
  use parameters
  use geometry
  use sparse_matrix
  use variables
  use title_mod
  use fcapp_c
  implicit none
  type(fc_calcuvw_opts) :: o
  type(fc_calcuvw_report) :: rep
  integer :: k
  character(len=1), parameter :: nm(3) = (/ 'U', 'V', 'W' /)
  if (lturb) stop 'fcapp calcuvw: turbulent stresses (calcstress) are not on the GPU path'
  o%nigrad = nigrad; o%nipgrad = nipgrad
  ! face_value dispatch (interpolation.f90:36-57) in the order of its if-chain
  o%limiter = 7
  if (lcds) then;           o%scheme = 0
  elseif (lcdsc) then;      o%scheme = 1
  elseif (lcds_flnt) then;  o%scheme = 2
  elseif (l2nd_flnt) then;  o%scheme = 3
  elseif (lmuscl_flnt) then; o%scheme = 4
  elseif (flux_limiter) then
    o%scheme = 5
    if (lsmart) then;      o%limiter = 0
    elseif (lavl) then;    o%limiter = 1
    elseif (lmuscl) then;  o%limiter = 2
    elseif (lumist) then;  o%limiter = 3
    elseif (lkoren) then;  o%limiter = 4
    elseif (lcharm) then;  o%limiter = 5
    elseif (lospre) then;  o%limiter = 6
    endif
  else;                     o%scheme = 4
  endif
  o%gds = gds(iu)
  o%urf = (/ urf(iu), urf(iv), urf(iw) /)
  o%sor = (/ sor(iu), sor(iv), sor(iw) /)
  o%nsw = (/ nsw(iu), nsw(iv), nsw(iw) /)
  o%bdf = merge(1, 0, bdf); o%btime = btime; o%timestep = timestep; o%cn = merge(1, 0, cn)
  o%const_mflux = merge(1, 0, const_mflux); o%gradPcmf = gradPcmf
  o%lbuoy = merge(1, 0, lcal(ien) .and. lbuoy); o%boussinesq = merge(1, 0, boussinesq)
  o%beta = beta; o%tref = tref; o%densit = densit; o%gravx = gravx; o%gravy = gravy; o%gravz = gravz
  o%viscos = viscos
  o%sol = solver_opts(iu)
  call fc_check(fc_upload(fc_ctx, FC_DEN, den, int(numTotal, c_size_t)), 'upload den')
  if (ninl > 0) call fc_check(fc_upload(fc_ctx, FC_FMI, fmi, int(ninl, c_size_t)), 'upload fmi')
  if (nout > 0) call fc_check(fc_upload(fc_ctx, FC_FMO, fmo, int(nout, c_size_t)), 'upload fmo')
  ! `a` is not uploaded: the U row sums read the stale diagonal of the previous solve (calcuvw.f90:423), and every
  ! solve of this build runs on the device, so the device copy of `a` is the current one by construction
  if (bdf .or. cn) then
    call fc_check(fc_upload(fc_ctx, FC_UO, uo, int(numTotal, c_size_t)), 'upload uo')
    call fc_check(fc_upload(fc_ctx, FC_VO, vo, int(numTotal, c_size_t)), 'upload vo')
    call fc_check(fc_upload(fc_ctx, FC_WO, wo, int(numTotal, c_size_t)), 'upload wo')
    call fc_check(fc_upload(fc_ctx, FC_UOO, uoo, int(numTotal, c_size_t)), 'upload uoo')
    call fc_check(fc_upload(fc_ctx, FC_VOO, voo, int(numTotal, c_size_t)), 'upload voo')
    call fc_check(fc_upload(fc_ctx, FC_WOO, woo, int(numTotal, c_size_t)), 'upload woo')
  end if
  if (lcal(ien) .and. lbuoy) call fc_check(fc_upload(fc_ctx, FC_T, t, int(numTotal, c_size_t)), 'upload t')
  call fc_check(fc_calcuvw_host(fc_ctx, o, u, v, w, p, vis, flmass, apu, apv, apw, rep), 'fc_calcuvw_host')
  ! module arrays other Fortran routines read afterwards
  call fc_check(fc_download(fc_ctx, FC_DUDXI, dUdxi, int(3*numCells, c_size_t)), 'download dUdxi')
  call fc_check(fc_download(fc_ctx, FC_DVDXI, dVdxi, int(3*numCells, c_size_t)), 'download dVdxi')
  call fc_check(fc_download(fc_ctx, FC_DWDXI, dWdxi, int(3*numCells, c_size_t)), 'download dWdxi')
  call fc_check(fc_download(fc_ctx, FC_DPDXI, dPdxi, int(3*numCells, c_size_t)), 'download dPdxi')
  call fc_check(fc_download(fc_ctx, FC_A, a, int(nnz, c_size_t)), 'download a')
  do k = 1, 3
    if (rep%rep(k)%iters > 0) resor(iu+k-1) = rep%rep(k)%res0     ! bicgstab.f90: resor(ifi) = res0
    write(6,'(3a,1PE10.3,a,1PE10.3,a,I0)') '  BiCGStab(ILU(0)):  Solving for ',nm(k), &
    ', Initial residual = ',rep%rep(k)%res0,', Final residual = ',rep%rep(k)%resl,', No Iterations ',rep%rep(k)%iters
  end do
end subroutine

! src/PISO_multiple_correction.f90:2 and src/PIMPLE_multiple_correction.f90:2 -- run directly after calcuvw: the
! device still holds the momentum matrix (backed up as h = a), ap*, u, v, w, the old time levels and flmass
subroutine fcapp_piso(pimple)
  use types
  use parameters
  use geometry
  use sparse_matrix
  use variables
  use title_mod
  use fcapp_c
  implicit none
  logical, intent(in) :: pimple
  type(fc_piso_opts) :: o
  type(fc_piso_report) :: rep
  integer :: k
  o%ncorr = ncorr; o%npcor = npcor; o%nigrad = nigrad; o%nipgrad = nipgrad; o%pRefCell = pRefCell
  o%pimple = merge(1, 0, pimple); o%urf_p = urf(ip)
  o%const_mflux = merge(1, 0, const_mflux); o%flomas = flomas
  o%bdf = merge(1, 0, bdf); o%btime = btime; o%timestep = timestep; o%cn = merge(1, 0, cn)
  o%lbuoy = merge(1, 0, lcal(ien) .and. lbuoy); o%boussinesq = merge(1, 0, boussinesq)
  o%beta = beta; o%tref = tref; o%densit = densit; o%gravx = gravx; o%gravy = gravy; o%gravz = gravz
  o%sol = solver_opts(ip)
  call fc_check(fc_upload(fc_ctx, FC_PP, pp, int(numTotal, c_size_t)), 'upload pp')   ! pp is not reset (PISO :199)
  call fc_check(fc_piso(fc_ctx, o, rep), 'fc_piso')
  call fc_check(fc_download(fc_ctx, FC_U, u, int(numTotal, c_size_t)), 'download u')
  call fc_check(fc_download(fc_ctx, FC_V, v, int(numTotal, c_size_t)), 'download v')
  call fc_check(fc_download(fc_ctx, FC_W, w, int(numTotal, c_size_t)), 'download w')
  call fc_check(fc_download(fc_ctx, FC_P, p, int(numTotal, c_size_t)), 'download p')
  call fc_check(fc_download(fc_ctx, FC_PP, pp, int(numTotal, c_size_t)), 'download pp')
  call fc_check(fc_download(fc_ctx, FC_FLMASS, flmass, int(numInnerFaces, c_size_t)), 'download flmass')
  call fc_check(fc_download(fc_ctx, FC_DPDXI, dPdxi, int(3*numCells, c_size_t)), 'download dPdxi')
  do k = 1, min(rep%nsolves, 16)
    if (rep%rep(k)%iters > 0) resor(ip) = rep%rep(k)%res0
    write(6,'(3a,1PE10.3,a,1PE10.3,a,I0)') '  PCG(IC0):  Solving for ',trim(chvarSolver(ip)), &
    ', Initial residual = ',rep%rep(k)%res0,', Final residual = ',rep%rep(k)%resl,', No Iterations ',rep%rep(k)%iters
  end do
  sumLocalContErr = rep%sumLocalContErr
  globalContErr = rep%globalContErr
  cumulativeContErr = cumulativeContErr + globalContErr
  write(6,'(3(a,es10.3))') "  time step continuity errors : sum local = ", sumLocalContErr, &
 &                          ", global = ", globalContErr, ", cumulative = ", cumulativeContErr
end subroutine

subroutine PISO_multiple_correction
  implicit none
  call fcapp_piso(.false.)
end subroutine

subroutine PIMPLE_multiple_correction
  implicit none
  call fcapp_piso(.true.)
end subroutine

! src-parallel/exchange.f90:3 and global_sum_mpi.f90:4 (MPI build: fc_comm_init after MPI_Init, the
! 128-byte id broadcast with MPI_BCAST from rank 0)
subroutine exchange(phi)
  use types
  use geometry, only: numTotal
  use fcapp_c
  implicit none
  real(dp), intent(inout) :: phi(numTotal)
  call fc_check(fc_upload(fc_ctx, FC_SCRATCH_T, phi, int(numTotal, c_size_t)), 'upload phi')
  call fc_check(fc_exchange(fc_ctx, FC_SCRATCH_T), 'fc_exchange')
  call fc_check(fc_download(fc_ctx, FC_SCRATCH_T, phi, int(numTotal, c_size_t)), 'download phi')
end subroutine

subroutine global_sum(phi)
  use types
  use fcapp_c
  implicit none
  real(dp) :: phi
  call fc_check(fc_global_sum(fc_ctx, phi), 'fc_global_sum')
end subroutine
