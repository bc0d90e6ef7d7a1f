!***************************************************************************************************
!  lsqr_b200_shim.F90 -- Fortran 2008 host layer of the B200-native LSQR engine.
!
!  Keeps the reference's module / type / binding names AND argument lists (jacobwilliams/LSQR
!  src/lsqr.f90:16-82, src/lsqr_kinds.F90:16-28, src/lsqrblas.f90:8-16) so that README.md:24-51,
!  test/lsqrtest_ez.f90 and test/lsqrtest_module.f90 compile unchanged against this file +
!  liblsqr_b200.so instead of the reference's src/*.f90:
!
!      use lsqr_kinds
!      use lsqr_module, only: lsqr_solver_ez
!      type(lsqr_solver_ez) :: solver
!      call solver%initialize(m,n,a,irow,icol)          ! src/lsqr.f90:91   -> lsqr_b200_ez_initialize
!      call solver%solve(b,damp,x,istop)                ! src/lsqr.f90:207  -> lsqr_b200_ez_solve
!
!      use lsqpblas_module                              ! test/lsqrtest_module.f90:26
!      type,extends(lsqr_solver) :: test_solver ; procedure :: aprod => my_host_aprod ; end type
!      call s%acheck(m,n,nout,eps,v,w,x,y,inform)       ! src/lsqr.f90:908  -> lsqr_b200_acheck_host
!      call s%lsqr(m,n,damp,wantse,u,v,w,x,se,...)      ! src/lsqr.f90:432  -> lsqr_b200_lsqr_host
!      call s%xcheck(m,n,nout,anorm,damp,eps,b,u,v,w,x,inform,test1,test2,test3)     ! :1015 -> lsqr_b200_xcheck_host
!
!  Everything numerical happens on the GPU behind the C ABI of include/lsqr_b200.h, which this file
!  binds through iso_c_binding.  An operator written as HOST code (the reference's aprod_func) costs a
!  host round trip per product; a device operator (type lsqr_solver_dev below: device addresses + CUDA
!  stream) is the fast path.  Where the reference executes  error stop '<message>'  the shim does the
!  same with the same literal message (the C ABI returns the code, the message comes from
!  lsqr_b200_error_message).
!
!  NOT COMPILED IN THE BUILD IMAGE: it has no Fortran compiler (no gfortran/flang/nvfortran), so this
!  file is delivered as source; the host layer that is compiled and tested is the C++ mirror
!  include/lsqr_b200.hpp and the Python mirror lsqr_b200/solver.py (same names, same argument order),
!  and tests/test_host_logic.py checks the bind(C) struct below field by field against the C header.
!  Build where a compiler exists:
!      gfortran -O2 -c fortran/lsqr_b200_shim.F90
!      gfortran -O2 my_program.f90 lsqr_b200_shim.o -Llsqr_b200/lib -llsqr_b200 -Wl,-rpath,$PWD/lsqr_b200/lib
!***************************************************************************************************

module lsqr_kinds                                   ! src/lsqr_kinds.F90:16-28
   use iso_fortran_env, only: real64
   implicit none
   private
   ! The engine computes in IEEE binary64 only (north_star: FP64 SpMV); the reference's REAL32 / REAL128 builds
   ! (src/lsqr_kinds.F90:16-22) have no counterpart and -DREAL32 / -DREAL128 are rejected at compile time.
#if defined(REAL32) || defined(REAL128)
#error "lsqr_b200 is an FP64 engine: build without -DREAL32 / -DREAL128"
#endif
   integer,parameter,public :: wp = real64
   real(wp),parameter,public :: zero = 0.0_wp
   real(wp),parameter,public :: one  = 1.0_wp
end module lsqr_kinds

!> iso_c_binding interfaces of include/lsqr_b200.h (one per exported entry point that the shim uses)
module lsqr_b200_c
   use iso_c_binding
   implicit none
   public

   type,bind(C) :: lsqr_b200_options                ! struct lsqr_b200_options (field order = the C header's)
      real(c_double)  :: atol = 0.0_c_double
      real(c_double)  :: btol = 0.0_c_double
      real(c_double)  :: conlim = 0.0_c_double
      integer(c_int32_t) :: itnlim = 100_c_int32_t
      integer(c_int32_t) :: device = -1_c_int32_t
      type(c_ptr)     :: stream = c_null_ptr
      type(c_funptr)  :: log = c_null_funptr
      type(c_ptr)     :: log_user = c_null_ptr
      type(c_funptr)  :: iter = c_null_funptr
      type(c_ptr)     :: iter_user = c_null_ptr
      integer(c_int32_t) :: engine = 0
      integer(c_int32_t) :: use_graph = 1
      integer(c_int32_t) :: profile = 0
      integer(c_int32_t) :: spmv_variant = 0
      integer(c_int32_t) :: world_size = 1
      integer(c_int32_t) :: rank = 0
      type(c_ptr)     :: nccl_unique_id = c_null_ptr
      integer(c_int64_t) :: m_global = 0
   end type lsqr_b200_options

   interface
      function lsqr_b200_error_message(code) bind(C,name='lsqr_b200_error_message') result(msg)
         import :: c_int, c_ptr
         integer(c_int),value :: code
         type(c_ptr) :: msg
      end function
      subroutine lsqr_b200_default_options(opts) bind(C,name='lsqr_b200_default_options')
         import :: lsqr_b200_options
         type(lsqr_b200_options),intent(out) :: opts
      end subroutine
      function lsqr_b200_ez_initialize(me,m,n,size_a,a,size_irow,irow,size_icol,icol,opts) &
               bind(C,name='lsqr_b200_ez_initialize') result(rc)
         import :: c_ptr, c_int, c_int32_t, c_int64_t, c_double, lsqr_b200_options
         type(c_ptr),intent(out) :: me
         integer(c_int32_t),value :: m, n
         integer(c_int64_t),value :: size_a, size_irow, size_icol
         real(c_double),intent(in) :: a(*)
         integer(c_int32_t),intent(in) :: irow(*), icol(*)
         type(lsqr_b200_options),intent(in) :: opts
         integer(c_int) :: rc
      end function
      function lsqr_b200_ez_solve(me,b,damp,x,istop,se,itn,anorm,acond,rnorm,arnorm,xnorm) &
               bind(C,name='lsqr_b200_ez_solve') result(rc)
         import :: c_ptr, c_int, c_int32_t, c_double
         type(c_ptr),value :: me
         real(c_double),intent(in) :: b(*)
         real(c_double),value :: damp
         real(c_double),intent(out) :: x(*)
         integer(c_int32_t),intent(out) :: istop
         type(c_ptr),value :: se                    ! c_null_ptr = wantse .false.
         integer(c_int32_t),intent(out) :: itn
         real(c_double),intent(out) :: anorm, acond, rnorm, arnorm, xnorm
         integer(c_int) :: rc
      end function
      function lsqr_b200_ez_aprod(me,mode,m,n,x,y) bind(C,name='lsqr_b200_ez_aprod') result(rc)
         import :: c_ptr, c_int, c_int32_t, c_double
         type(c_ptr),value :: me
         integer(c_int32_t),value :: mode, m, n
         real(c_double),intent(inout) :: x(*), y(*)     ! host (or device) arrays
         integer(c_int) :: rc
      end function
      function lsqr_b200_ez_aprod_device(h,mode,m,n,x,y,stream) bind(C,name='lsqr_b200_ez_aprod_device') result(rc)
         import :: c_ptr, c_int, c_int32_t
         type(c_ptr),value :: h, x, y, stream
         integer(c_int32_t),value :: mode, m, n
         integer(c_int) :: rc
      end function
      subroutine lsqr_b200_ez_destroy(me) bind(C,name='lsqr_b200_ez_destroy')
         import :: c_ptr
         type(c_ptr),value :: me
      end subroutine
      ! ---- abstract class, the reference's signatures: host arrays + host operator (src/lsqr.f90:432, :908, :1015)
      function lsqr_b200_lsqr_host(aprod,aprod_user,m,n,damp,wantse,u,v,w,x,se,atol,btol,conlim,itnlim,opts, &
                                   istop,itn,anorm,acond,rnorm,arnorm,xnorm) bind(C,name='lsqr_b200_lsqr_host') result(rc)
         import :: c_ptr, c_funptr, c_int, c_int32_t, c_double, lsqr_b200_options
         type(c_funptr),value :: aprod
         type(c_ptr),value :: aprod_user
         integer(c_int32_t),value :: m, n, wantse, itnlim
         real(c_double),value :: damp, atol, btol, conlim
         real(c_double),intent(inout) :: u(*), v(*), w(*), x(*)
         type(c_ptr),value :: se
         type(lsqr_b200_options),intent(in) :: opts
         integer(c_int32_t),intent(out) :: istop, itn
         real(c_double),intent(out) :: anorm, acond, rnorm, arnorm, xnorm
         integer(c_int) :: rc
      end function
      function lsqr_b200_acheck_host(aprod,aprod_user,m,n,eps,v,w,x,y,opts,inform,relerr) &
               bind(C,name='lsqr_b200_acheck_host') result(rc)
         import :: c_ptr, c_funptr, c_int, c_int32_t, c_double, lsqr_b200_options
         type(c_funptr),value :: aprod
         type(c_ptr),value :: aprod_user
         integer(c_int32_t),value :: m, n
         real(c_double),value :: eps
         real(c_double),intent(inout) :: v(*), w(*), x(*), y(*)
         type(lsqr_b200_options),intent(in) :: opts
         integer(c_int32_t),intent(out) :: inform
         real(c_double),intent(out) :: relerr
         integer(c_int) :: rc
      end function
      function lsqr_b200_xcheck_host(aprod,aprod_user,m,n,anorm,damp,eps,b,u,v,w,x,opts,inform,test1,test2,test3,norms) &
               bind(C,name='lsqr_b200_xcheck_host') result(rc)
         import :: c_ptr, c_funptr, c_int, c_int32_t, c_double, lsqr_b200_options
         type(c_funptr),value :: aprod
         type(c_ptr),value :: aprod_user
         integer(c_int32_t),value :: m, n
         real(c_double),value :: anorm, damp, eps
         real(c_double),intent(in) :: b(*), x(*)
         real(c_double),intent(inout) :: u(*), v(*), w(*)
         type(lsqr_b200_options),intent(in) :: opts
         integer(c_int32_t),intent(out) :: inform
         real(c_double),intent(out) :: test1, test2, test3
         type(c_ptr),value :: norms
         integer(c_int) :: rc
      end function
      ! ---- the same with a device operator (device addresses; the fast path)
      function lsqr_b200_lsqr(aprod,aprod_user,m,n,damp,wantse,u,v,w,x,se,atol,btol,conlim,itnlim,opts, &
                              istop,itn,anorm,acond,rnorm,arnorm,xnorm) bind(C,name='lsqr_b200_lsqr') result(rc)
         import :: c_ptr, c_funptr, c_int, c_int32_t, c_double, lsqr_b200_options
         type(c_funptr),value :: aprod
         type(c_ptr),value :: aprod_user
         integer(c_int32_t),value :: m, n, wantse, itnlim
         real(c_double),value :: damp, atol, btol, conlim
         type(c_ptr),value :: u, v, w, x, se        ! DEVICE pointers
         type(lsqr_b200_options),intent(in) :: opts
         integer(c_int32_t),intent(out) :: istop, itn
         real(c_double),intent(out) :: anorm, acond, rnorm, arnorm, xnorm
         integer(c_int) :: rc
      end function
      ! ---- BLAS-1 (src/lsqrblas.f90); host or device arrays, stride 1
      function lsqr_b200_dnrm2(n,x,res,stream) bind(C,name='lsqr_b200_dnrm2') result(rc)
         import :: c_int, c_int64_t, c_double, c_ptr
         integer(c_int64_t),value :: n
         real(c_double),intent(in) :: x(*)
         real(c_double),intent(out) :: res
         type(c_ptr),value :: stream
         integer(c_int) :: rc
      end function
      function lsqr_b200_ddot(n,x,y,res,stream) bind(C,name='lsqr_b200_ddot') result(rc)
         import :: c_int, c_int64_t, c_double, c_ptr
         integer(c_int64_t),value :: n
         real(c_double),intent(in) :: x(*), y(*)
         real(c_double),intent(out) :: res
         type(c_ptr),value :: stream
         integer(c_int) :: rc
      end function
      function lsqr_b200_dscal(n,da,x,stream) bind(C,name='lsqr_b200_dscal') result(rc)
         import :: c_int, c_int64_t, c_double, c_ptr
         integer(c_int64_t),value :: n
         real(c_double),value :: da
         real(c_double),intent(inout) :: x(*)
         type(c_ptr),value :: stream
         integer(c_int) :: rc
      end function
      function lsqr_b200_dcopy(n,x,y,stream) bind(C,name='lsqr_b200_dcopy') result(rc)
         import :: c_int, c_int64_t, c_double, c_ptr
         integer(c_int64_t),value :: n
         real(c_double),intent(in) :: x(*)
         real(c_double),intent(out) :: y(*)
         type(c_ptr),value :: stream
         integer(c_int) :: rc
      end function
   end interface

contains

   !> the C string returned by lsqr_b200_error_message as a Fortran string
   function error_text(code) result(s)
      integer(c_int),intent(in) :: code
      character(len=:),allocatable :: s
      character(kind=c_char),pointer :: p(:)
      type(c_ptr) :: cp
      integer :: i, n
      cp = lsqr_b200_error_message(code)
      call c_f_pointer(cp, p, [256])
      n = 0
      do while (n < 256)
         if (p(n+1) == c_null_char) exit
         n = n + 1
      end do
      allocate(character(len=n) :: s)
      do i = 1, n
         s(i:i) = p(i)
      end do
   end function error_text

   subroutine stop_on(rc)
      integer(c_int),intent(in) :: rc
      if (rc /= 0) error stop error_text(rc)     ! same literal messages as src/lsqr.f90:109-111,152,197
   end subroutine stop_on

   !> nout log callback: `user` points at the default integer holding the Fortran unit number
   subroutine log_to_unit(user, line) bind(C)
      type(c_ptr),value :: user, line
      integer,pointer :: nout
      character(kind=c_char),pointer :: p(:)
      integer :: n
      call c_f_pointer(user, nout)
      call c_f_pointer(line, p, [512])
      n = 0
      do while (n < 512)
         if (p(n+1) == c_null_char) exit
         n = n + 1
      end do
      write(nout,'(*(a))') p(1:n)
   end subroutine log_to_unit

end module lsqr_b200_c

!> src/lsqrblas.f90:8-16 [sic: the reference spells the module `lsqpblas_module`].  Same names and argument lists
!> (n, dx, incx [, dy, incy]); the arithmetic runs on the GPU (host arrays are staged).  Strided arguments are packed
!> into contiguous temporaries by Fortran array sections; negative increments follow the BLAS convention of the
!> reference (the vector is traversed backwards from element 1 + (1-n)*inc).
module lsqpblas_module
   use lsqr_kinds
   use lsqr_b200_c
   use iso_c_binding
   implicit none
   private
   public :: dcopy, ddot, dnrm2, dscal

contains

   pure function first_index(n, inc) result(i0)
      integer,intent(in) :: n, inc
      integer :: i0
      i0 = 1
      if (inc < 0) i0 = (-n+1)*inc + 1           ! src/lsqrblas.f90:43-44,93-94
   end function first_index

   subroutine dcopy(n,dx,incx,dy,incy)            ! src/lsqrblas.f90:25-67
      integer  :: incx, incy, n
      real(wp) :: dx(*), dy(*)
      real(wp),allocatable :: t(:)
      integer :: i, ix, iy
      if (n <= 0) return
      if (incx == 1 .and. incy == 1) then
         call stop_on(lsqr_b200_dcopy(int(n,c_int64_t), dx, dy, c_null_ptr))
         return
      end if
      allocate(t(n))
      ix = first_index(n, incx); iy = first_index(n, incy)
      do i = 1, n
         t(i) = dx(ix + (i-1)*incx)
      end do
      do i = 1, n
         dy(iy + (i-1)*incy) = t(i)
      end do
   end subroutine dcopy

   real(wp) function ddot(n,dx,incx,dy,incy)      ! src/lsqrblas.f90:74-116
      integer  :: incx, incy, n
      real(wp) :: dx(*), dy(*)
      real(wp),allocatable :: tx(:), ty(:)
      integer :: i, ix, iy
      ddot = zero
      if (n <= 0) return
      if (incx == 1 .and. incy == 1) then
         call stop_on(lsqr_b200_ddot(int(n,c_int64_t), dx, dy, ddot, c_null_ptr))
         return
      end if
      allocate(tx(n), ty(n))
      ix = first_index(n, incx); iy = first_index(n, incy)
      do i = 1, n
         tx(i) = dx(ix + (i-1)*incx); ty(i) = dy(iy + (i-1)*incy)
      end do
      call stop_on(lsqr_b200_ddot(int(n,c_int64_t), tx, ty, ddot, c_null_ptr))
   end function ddot

   real(wp) function dnrm2(n,x,incx)              ! src/lsqrblas.f90:123-159 (scaled: no overflow / underflow)
      integer  :: incx, n
      real(wp) :: x(*)
      real(wp),allocatable :: t(:)
      integer :: i
      dnrm2 = zero
      if (n < 1 .or. incx < 1) return             ! :131
      if (incx == 1) then
         call stop_on(lsqr_b200_dnrm2(int(n,c_int64_t), x, dnrm2, c_null_ptr))
         return
      end if
      allocate(t(n))
      do i = 1, n
         t(i) = x(1 + (i-1)*incx)
      end do
      call stop_on(lsqr_b200_dnrm2(int(n,c_int64_t), t, dnrm2, c_null_ptr))
   end function dnrm2

   subroutine dscal(n,da,dx,incx)                 ! src/lsqrblas.f90:166-201
      integer  :: incx, n
      real(wp) :: da, dx(*)
      real(wp),allocatable :: t(:)
      integer :: i
      if (n <= 0 .or. incx <= 0) return           ! :174
      if (incx == 1) then
         call stop_on(lsqr_b200_dscal(int(n,c_int64_t), da, dx, c_null_ptr))
         return
      end if
      allocate(t(n))
      do i = 1, n
         t(i) = dx(1 + (i-1)*incx)
      end do
      call stop_on(lsqr_b200_dscal(int(n,c_int64_t), da, t, c_null_ptr))
      do i = 1, n
         dx(1 + (i-1)*incx) = t(i)
      end do
   end subroutine dscal

end module lsqpblas_module

module lsqr_module
   use lsqr_kinds
   use lsqr_b200_c
   use iso_c_binding
   implicit none
   private

   !> src/lsqr.f90:16-30, identical public interface: a deferred host `aprod` and the three public procedures.
   type,abstract,public :: lsqr_solver
      private
   contains
      private
      procedure(aprod_func),deferred,public :: aprod   !! User function to access the sparse matrix `A` (host arrays).
      procedure,public :: lsqr                          !! src/lsqr.f90:432
      procedure,public :: acheck                        !! src/lsqr.f90:908
      procedure,public :: xcheck                        !! src/lsqr.f90:1015
   end type lsqr_solver

   abstract interface
      subroutine aprod_func ( me, mode, m, n, x, y )    ! src/lsqr.f90:67-82, verbatim interface
         import :: wp, lsqr_solver
         implicit none
         class(lsqr_solver),intent(inout) :: me
         integer,intent(in) :: mode          !! 1: y = y + A*x ; 2: x = x + A'*y
         integer,intent(in) :: m
         integer,intent(in) :: n
         real(wp),dimension(:),intent(inout) :: x
         real(wp),dimension(:),intent(inout) :: y
      end subroutine aprod_func
   end interface

   !> The fast path (no reference counterpart): an operator that works on DEVICE vectors.  x and y arrive as device
   !> addresses and the work must be enqueued on `stream` (cudaStream_t) without synchronising -- e.g. CUDA Fortran,
   !> OpenACC host_data, or a CUDA C routine.  The host `aprod` of such a type may simply error stop.
   type,abstract,public,extends(lsqr_solver) :: lsqr_solver_dev
   contains
      procedure(aprod_dev_func),deferred,public :: aprod_dev
      procedure,public :: lsqr_dev                      !! LSQR on device vectors u(m), v(n), w(n), x(n), se(n)
   end type lsqr_solver_dev

   abstract interface
      subroutine aprod_dev_func(me, mode, m, n, x, y, stream)
         import :: lsqr_solver_dev, c_ptr
         class(lsqr_solver_dev),intent(inout) :: me
         integer,intent(in) :: mode, m, n
         type(c_ptr),value  :: x, y          !! DEVICE addresses of x(n), y(m)
         type(c_ptr),value  :: stream
      end subroutine aprod_dev_func
   end interface

   !> src/lsqr.f90:32-65: same public bindings (initialize, solve, aprod) with the reference's argument lists.
   type,public,extends(lsqr_solver) :: lsqr_solver_ez
      private
      type(c_ptr) :: handle = c_null_ptr
      integer :: m = 0, n = 0
      integer,pointer :: nout => null()   ! unit number handed to the log callback
   contains
      private
      procedure,public :: initialize => initialize_ez  !! Constructor. Must be call first.
      procedure,public :: solve      => solve_ez
      procedure,public :: aprod      => aprod_ez        !! src/lsqr.f90:134-143: (me,mode,m,n,x,y), host arrays
      procedure,public :: aprod_dev  => aprod_ez_dev    !! the same product on device addresses, enqueued on a stream
      final :: destroy_ez
   end type lsqr_solver_ez

   type :: hook_box                      ! what the C trampolines need to reach the Fortran object
      class(lsqr_solver),pointer :: obj => null()
      class(lsqr_solver_dev),pointer :: dev => null()
   end type hook_box

contains

   ! ------------------------------------------------------------------ initialize_ez, src/lsqr.f90:91-127
   subroutine initialize_ez(me,m,n,a,irow,icol,atol,btol,conlim,itnlim,nout)
      class(lsqr_solver_ez),intent(out) :: me
      integer,intent(in)                :: m, n
      real(wp),dimension(:),intent(in)  :: a
      integer,dimension(:),intent(in)   :: irow, icol
      real(wp),intent(in),optional      :: atol, btol, conlim
      integer,intent(in),optional       :: itnlim, nout
      type(lsqr_b200_options) :: o
      call lsqr_b200_default_options(o)
      if (present(atol))   o%atol   = atol
      if (present(btol))   o%btol   = btol
      if (present(conlim)) o%conlim = conlim
      if (present(itnlim)) o%itnlim = int(itnlim, c_int32_t)
      if (present(nout)) then
         if (nout /= 0) then
            allocate(me%nout); me%nout = nout
            o%log = c_funloc(log_to_unit)
            o%log_user = c_loc(me%nout)
         end if
      end if
      me%m = m; me%n = n
      call stop_on(lsqr_b200_ez_initialize(me%handle, int(m,c_int32_t), int(n,c_int32_t), &
                                           int(size(a),c_int64_t), a, int(size(irow),c_int64_t), int(irow,c_int32_t), &
                                           int(size(icol),c_int64_t), int(icol,c_int32_t), o))
   end subroutine initialize_ez

   ! ------------------------------------------------------------------ solve_ez, src/lsqr.f90:207-259
   subroutine solve_ez(me,b,damp,x,istop,se,itn,anorm,acond,rnorm,arnorm,xnorm)
      class(lsqr_solver_ez),intent(inout) :: me
      real(wp),dimension(:),intent(in)    :: b
      real(wp),intent(in)                 :: damp
      real(wp),dimension(:),intent(out)   :: x
      integer,intent(out)                 :: istop
      real(wp),dimension(:),intent(out),optional,target :: se
      integer,intent(out),optional        :: itn
      real(wp),intent(out),optional       :: anorm, acond, rnorm, arnorm, xnorm
      integer(c_int32_t) :: istop_, itn_
      real(c_double) :: anorm_, acond_, rnorm_, arnorm_, xnorm_
      type(c_ptr) :: se_
      if (.not. c_associated(me%handle)) error stop 'lsqr_solver_ez class not properly initialized'
      se_ = c_null_ptr
      if (present(se)) se_ = c_loc(se)
      call stop_on(lsqr_b200_ez_solve(me%handle, b, damp, x, istop_, se_, itn_, anorm_, acond_, rnorm_, arnorm_, xnorm_))
      istop = istop_
      if (present(itn))    itn    = itn_
      if (present(anorm))  anorm  = anorm_
      if (present(acond))  acond  = acond_
      if (present(rnorm))  rnorm  = rnorm_
      if (present(arnorm)) arnorm = arnorm_
      if (present(xnorm))  xnorm  = xnorm_
   end subroutine solve_ez

   ! ------------------------------------------------------------------ aprod_ez, src/lsqr.f90:134-200 (reference signature)
   subroutine aprod_ez(me, mode, m, n, x, y)
      class(lsqr_solver_ez),intent(inout) :: me
      integer,intent(in) :: mode, m, n
      real(wp),dimension(:),intent(inout) :: x, y
      call stop_on(lsqr_b200_ez_aprod(me%handle, int(mode,c_int32_t), int(m,c_int32_t), int(n,c_int32_t), x, y))
   end subroutine aprod_ez

   !> the same operator on device addresses (what a lsqr_solver_dev extension would forward to)
   subroutine aprod_ez_dev(me, mode, m, n, x, y, stream)
      class(lsqr_solver_ez),intent(inout) :: me
      integer,intent(in) :: mode, m, n
      type(c_ptr),value  :: x, y, stream
      call stop_on(lsqr_b200_ez_aprod_device(me%handle, int(mode,c_int32_t), int(m,c_int32_t), int(n,c_int32_t), x, y, stream))
   end subroutine aprod_ez_dev

   subroutine destroy_ez(me)
      type(lsqr_solver_ez),intent(inout) :: me
      if (c_associated(me%handle)) call lsqr_b200_ez_destroy(me%handle)
      me%handle = c_null_ptr
      if (associated(me%nout)) deallocate(me%nout)
   end subroutine destroy_ez

   ! ------------------------------------------------------------------ trampolines
   !> lsqr_b200_aprod_host_fn -> the Fortran object's host aprod (src/lsqr.f90:67-82)
   function host_trampoline(user, mode, m, n, x, y) bind(C) result(rc)
      type(c_ptr),value :: user
      integer(c_int32_t),value :: mode, m, n
      real(c_double),intent(inout),target :: x(*), y(*)
      integer(c_int) :: rc
      type(hook_box),pointer :: box
      real(wp),pointer :: xp(:), yp(:)
      call c_f_pointer(user, box)
      call c_f_pointer(c_loc(x), xp, [int(n)])
      call c_f_pointer(c_loc(y), yp, [int(m)])
      call box%obj%aprod(int(mode), int(m), int(n), xp, yp)
      rc = 0
   end function host_trampoline

   !> lsqr_b200_aprod_fn -> the Fortran object's device aprod
   function dev_trampoline(user, mode, m, n, x, y, stream) bind(C) result(rc)
      type(c_ptr),value :: user, x, y, stream
      integer(c_int32_t),value :: mode, m, n
      integer(c_int) :: rc
      type(hook_box),pointer :: box
      call c_f_pointer(user, box)
      call box%dev%aprod_dev(int(mode), int(m), int(n), x, y, stream)
      rc = 0
   end function dev_trampoline

   subroutine log_options(o, nout, unit_)
      type(lsqr_b200_options),intent(inout) :: o
      integer,intent(in) :: nout
      integer,intent(inout),target :: unit_
      call lsqr_b200_default_options(o)
      unit_ = nout
      if (nout /= 0) then
         o%log = c_funloc(log_to_unit)
         o%log_user = c_loc(unit_)
      end if
   end subroutine log_options

   ! ------------------------------------------------------------------ LSQR, src/lsqr.f90:432-435 (reference argument list)
   subroutine lsqr(me, m, n, damp, wantse, u, v, w, x, se, atol, btol, conlim, itnlim, nout, &
                   istop, itn, anorm, acond, rnorm, arnorm, xnorm)
      class(lsqr_solver),intent(inout),target :: me
      integer,intent(in)    :: m, n
      real(wp),intent(in)   :: damp
      logical,intent(in)    :: wantse
      real(wp),intent(inout):: u(m)
      real(wp),intent(inout):: v(n)
      real(wp),intent(inout):: w(n)
      real(wp),intent(out)  :: x(n)
      real(wp),dimension(*),intent(out),target :: se
      real(wp),intent(in)   :: atol, btol, conlim
      integer,intent(in)    :: itnlim, nout
      integer,intent(out)   :: istop, itn
      real(wp),intent(out)  :: anorm, acond, rnorm, arnorm, xnorm
      type(lsqr_b200_options) :: o
      type(hook_box),target :: box
      integer,target :: unit_
      integer(c_int32_t) :: istop_, itn_
      type(c_ptr) :: se_
      call log_options(o, nout, unit_)
      box%obj => me
      se_ = c_null_ptr
      if (wantse) se_ = c_loc(se)                   ! wantse = .false.: se is not touched and may be any length (:478-480)
      call stop_on(lsqr_b200_lsqr_host(c_funloc(host_trampoline), c_loc(box), int(m,c_int32_t), int(n,c_int32_t), damp, &
                                       merge(1_c_int32_t, 0_c_int32_t, wantse), u, v, w, x, se_, atol, btol, conlim, &
                                       int(itnlim,c_int32_t), o, istop_, itn_, anorm, acond, rnorm, arnorm, xnorm))
      istop = istop_
      itn = itn_
   end subroutine lsqr

   ! ------------------------------------------------------------------ acheck, src/lsqr.f90:908-909
   subroutine acheck(me, m, n, nout, eps, v, w, x, y, inform)
      class(lsqr_solver),intent(inout),target :: me
      integer,intent(in)   :: m, n, nout
      integer,intent(out)  :: inform
      real(wp),intent(in)  :: eps
      real(wp)             :: v(n), w(m), x(n), y(m)
      type(lsqr_b200_options) :: o
      type(hook_box),target :: box
      integer,target :: unit_
      integer(c_int32_t) :: inform_
      real(c_double) :: relerr
      call log_options(o, nout, unit_)
      box%obj => me
      call stop_on(lsqr_b200_acheck_host(c_funloc(host_trampoline), c_loc(box), int(m,c_int32_t), int(n,c_int32_t), eps, &
                                         v, w, x, y, o, inform_, relerr))
      inform = inform_
   end subroutine acheck

   ! ------------------------------------------------------------------ xcheck, src/lsqr.f90:1015-1017
   subroutine xcheck(me, m, n, nout, anorm, damp, eps, b, u, v, w, x, inform, test1, test2, test3)
      class(lsqr_solver),intent(inout),target :: me
      integer,intent(in)   :: m, n, nout
      integer,intent(out)  :: inform
      real(wp),intent(in)  :: anorm, damp, eps
      real(wp),intent(out) :: test1, test2, test3
      real(wp),intent(in)  :: b(m)
      real(wp),intent(out) :: u(m), v(n), w(n)
      real(wp),intent(in)  :: x(n)
      type(lsqr_b200_options) :: o
      type(hook_box),target :: box
      integer,target :: unit_
      integer(c_int32_t) :: inform_
      call log_options(o, nout, unit_)
      box%obj => me
      call stop_on(lsqr_b200_xcheck_host(c_funloc(host_trampoline), c_loc(box), int(m,c_int32_t), int(n,c_int32_t), &
                                         anorm, damp, eps, b, u, v, w, x, o, inform_, test1, test2, test3, c_null_ptr))
      inform = inform_
   end subroutine xcheck

   ! ------------------------------------------------------------------ LSQR with a device operator on device vectors
   subroutine lsqr_dev(me, m, n, damp, wantse, u, v, w, x, se, atol, btol, conlim, itnlim, nout, &
                       istop, itn, anorm, acond, rnorm, arnorm, xnorm, stream)
      class(lsqr_solver_dev),intent(inout),target :: me
      integer,intent(in)  :: m, n, itnlim, nout
      real(wp),intent(in) :: damp, atol, btol, conlim
      logical,intent(in)  :: wantse
      type(c_ptr),value   :: u, v, w, x, se         !! DEVICE addresses of u(m), v(n), w(n), x(n), se(n)
      integer,intent(out) :: istop, itn
      real(wp),intent(out):: anorm, acond, rnorm, arnorm, xnorm
      type(c_ptr),value,optional :: stream          !! cudaStream_t the vectors were produced on (default: stream 0 ordering)
      type(lsqr_b200_options) :: o
      type(hook_box),target :: box
      integer,target :: unit_
      integer(c_int32_t) :: istop_, itn_
      call log_options(o, nout, unit_)
      if (present(stream)) o%stream = stream
      box%dev => me
      call stop_on(lsqr_b200_lsqr(c_funloc(dev_trampoline), c_loc(box), int(m,c_int32_t), int(n,c_int32_t), damp, &
                                  merge(1_c_int32_t, 0_c_int32_t, wantse), u, v, w, x, se, atol, btol, conlim, &
                                  int(itnlim,c_int32_t), o, istop_, itn_, anorm, acond, rnorm, arnorm, xnorm))
      istop = istop_
      itn = itn_
   end subroutine lsqr_dev

end module lsqr_module
