!***************************************************************************************************
!  lsqr_b200_shim.f90 -- Fortran 2008 host layer of the B200-native LSQR engine.
!
!  Keeps the reference's module / type / binding names (jacobwilliams/LSQR src/lsqr.f90:16-82,
!  src/lsqr_kinds.F90:16-28) so that README.md:24-51 and test/lsqrtest_ez.f90 compile unchanged
!  against this file + liblsqr_b200.so instead of the reference's src/lsqr.f90:
!
!      use lsqr_kinds
!      use lsqr_module, only: lsqr_solver_ez
!      type(lsqr_solver_ez) :: solver
!      call solver%initialize(m,n,a,irow,icol)          ! src/lsqr.f90:91   -> lsqr_b200_ez_initialize
!      call solver%solve(b,damp,x,istop)                ! src/lsqr.f90:207  -> lsqr_b200_ez_solve
!
!  Everything numerical happens on the GPU behind the C ABI of include/lsqr_b200.h, which this file
!  binds through iso_c_binding.  Where the reference executes  error stop '<message>'  the shim
!  does the same with the same literal message (the C ABI returns the code, the message comes from
!  lsqr_b200_error_message).
!
!  NOT COMPILED IN THE BUILD IMAGE: it has no Fortran compiler (no gfortran/flang/nvfortran), so this
!  file is delivered as source; the host layer that is compiled and tested is the C++ mirror
!  include/lsqr_b200.hpp (same names, same argument order).  Build where a compiler exists:
!      gfortran -O2 -c fortran/lsqr_b200_shim.f90
!      gfortran -O2 my_program.f90 lsqr_b200_shim.o -Llsqr_b200/lib -llsqr_b200 -Wl,-rpath,$PWD/lsqr_b200/lib
!***************************************************************************************************

module lsqr_kinds                                   ! src/lsqr_kinds.F90:16-28 (wp = real64 only)
   use iso_fortran_env, only: real64
   implicit none
   private
   integer,parameter,public :: wp = real64
   real(wp),parameter,public :: zero = 0.0_wp
   real(wp),parameter,public :: one  = 1.0_wp
end module lsqr_kinds

!> iso_c_binding interfaces of include/lsqr_b200.h (one per exported entry point that the shim uses)
module lsqr_b200_c
   use iso_c_binding
   implicit none
   public

   type,bind(C) :: lsqr_b200_options                ! struct lsqr_b200_options
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
      integer(c_int32_t) :: engine = 0, use_graph = 1, profile = 0, spmv_variant = 0
      integer(c_int32_t) :: world_size = 1, rank = 0
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
         real(c_double),intent(inout) :: x(*), y(*)
         integer(c_int) :: rc
      end function
      subroutine lsqr_b200_ez_destroy(me) bind(C,name='lsqr_b200_ez_destroy')
         import :: c_ptr
         type(c_ptr),value :: me
      end subroutine
      ! low-level solver with a device-pointer operator (src/lsqr.f90:432, :67-82)
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

module lsqr_module
   use lsqr_kinds
   use lsqr_b200_c
   use iso_c_binding
   implicit none
   private

   !> src/lsqr.f90:16-30.  The operator works on DEVICE vectors: x and y arrive as device addresses and the work
   !> must be enqueued on `stream` (cudaStream_t) without synchronising -- e.g. by CUDA Fortran / OpenACC
   !> host_data code, or by calling a CUDA C routine.
   type,abstract,public :: lsqr_solver
   contains
      procedure(aprod_func),deferred,public :: aprod
      procedure,public :: lsqr
   end type lsqr_solver

   abstract interface
      subroutine aprod_func(me, mode, m, n, x, y, stream)
         import :: lsqr_solver, c_ptr
         class(lsqr_solver),intent(inout) :: me
         integer,intent(in) :: mode          !! 1: y = y + A*x ; 2: x = x + A'*y   (src/lsqr.f90:71-76)
         integer,intent(in) :: m, n
         type(c_ptr),value  :: x, y          !! DEVICE addresses of x(n), y(m)
         type(c_ptr),value  :: stream
      end subroutine aprod_func
   end interface

   !> src/lsqr.f90:32-65
   type,public,extends(lsqr_solver) :: lsqr_solver_ez
      private
      type(c_ptr) :: handle = c_null_ptr
      integer :: m = 0, n = 0
      integer,pointer :: nout => null()   ! unit number handed to the log callback
   contains
      procedure,public :: initialize => initialize_ez
      procedure,public :: solve      => solve_ez
      procedure,public :: aprod      => aprod_ez_dev
      procedure,public :: aprod_host => aprod_ez
      final :: destroy_ez
   end type lsqr_solver_ez

   type :: hook_box                      ! what the C trampoline needs to reach the Fortran object
      class(lsqr_solver),pointer :: obj => null()
   end type hook_box

contains

   subroutine stop_on(rc)
      integer(c_int),intent(in) :: rc
      if (rc /= 0) error stop error_text(rc)     ! same literal messages as src/lsqr.f90:109-111,152,197
   end subroutine stop_on

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

   ! ------------------------------------------------------------------ aprod_ez, src/lsqr.f90:134-200 (host vectors)
   subroutine aprod_ez(me, mode, m, n, x, y)
      class(lsqr_solver_ez),intent(inout) :: me
      integer,intent(in) :: mode, m, n
      real(wp),dimension(:),intent(inout) :: x, y
      call stop_on(lsqr_b200_ez_aprod(me%handle, int(mode,c_int32_t), int(m,c_int32_t), int(n,c_int32_t), x, y))
   end subroutine aprod_ez

   !> the same operator as the deferred `aprod` of the low-level class (device addresses)
   subroutine aprod_ez_dev(me, mode, m, n, x, y, stream)
      class(lsqr_solver_ez),intent(inout) :: me
      integer,intent(in) :: mode, m, n
      type(c_ptr),value  :: x, y, stream
      interface
         function lsqr_b200_ez_aprod_device(h,mode,m,n,x,y,stream) bind(C,name='lsqr_b200_ez_aprod_device') result(rc)
            import :: c_ptr, c_int, c_int32_t
            type(c_ptr),value :: h, x, y, stream
            integer(c_int32_t),value :: mode, m, n
            integer(c_int) :: rc
         end function
      end interface
      call stop_on(lsqr_b200_ez_aprod_device(me%handle, int(mode,c_int32_t), int(m,c_int32_t), int(n,c_int32_t), x, y, stream))
   end subroutine aprod_ez_dev

   subroutine destroy_ez(me)
      type(lsqr_solver_ez),intent(inout) :: me
      if (c_associated(me%handle)) call lsqr_b200_ez_destroy(me%handle)
      me%handle = c_null_ptr
      if (associated(me%nout)) deallocate(me%nout)
   end subroutine destroy_ez

   ! ------------------------------------------------------------------ LSQR, src/lsqr.f90:432-882 (device vectors)
   function trampoline(user, mode, m, n, x, y, stream) bind(C) result(rc)
      type(c_ptr),value :: user, x, y, stream
      integer(c_int32_t),value :: mode, m, n
      integer(c_int) :: rc
      type(hook_box),pointer :: box
      call c_f_pointer(user, box)
      call box%obj%aprod(int(mode), int(m), int(n), x, y, stream)
      rc = 0
   end function trampoline

   subroutine lsqr(me, m, n, damp, wantse, u, v, w, x, se, atol, btol, conlim, itnlim, nout, &
                   istop, itn, anorm, acond, rnorm, arnorm, xnorm)
      class(lsqr_solver),intent(inout),target :: me
      integer,intent(in)  :: m, n, itnlim, nout
      real(wp),intent(in) :: damp, atol, btol, conlim
      logical,intent(in)  :: wantse
      type(c_ptr),value   :: u, v, w, x, se         !! DEVICE addresses of u(m), v(n), w(n), x(n), se(n)
      integer,intent(out) :: istop, itn
      real(wp),intent(out):: anorm, acond, rnorm, arnorm, xnorm
      type(lsqr_b200_options) :: o
      type(hook_box),target :: box
      integer,target :: unit_
      integer(c_int32_t) :: istop_, itn_
      call lsqr_b200_default_options(o)
      unit_ = nout
      if (nout /= 0) then
         o%log = c_funloc(log_to_unit)
         o%log_user = c_loc(unit_)
      end if
      box%obj => me
      call stop_on(lsqr_b200_lsqr(c_funloc(trampoline), c_loc(box), int(m,c_int32_t), int(n,c_int32_t), damp, &
                                  merge(1_c_int32_t, 0_c_int32_t, wantse), u, v, w, x, se, atol, btol, conlim, &
                                  int(itnlim,c_int32_t), o, istop_, itn_, anorm, acond, rnorm, arnorm, xnorm))
      istop = istop_
      itn = itn_
   end subroutine lsqr

end module lsqr_module
