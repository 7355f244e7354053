! fitpack_b200_curves.f90 -- a Fortran `fitpack_curve` on top of libfitpack_b200.so
!
! The reference's object layer (src/fitpack_curves.f90:54-131) is a derived type whose methods call the Fortran core.
! This module offers a type with the SAME method names -- new_points, new_fit, fit, interpolate, least_squares, eval,
! dfdx, dfdx_all, integral, fourier_coefficients, smoothing, mse, degree, bc, comm_size / comm_pack / comm_expand,
! destroy -- whose methods call the handle C-ABI of this library (include/fitpack_b200.h, the same functions as the
! reference's include/fitpack_curves_c.h:45-63), i.e. the CUDA kernels.  Host code switches with one line:
!
!     use fitpack_b200_curves, only: fitpack_curve => fitpack_curve_b200
!
! State (data, knots, coefficients, work arrays, iopt) lives behind the handle, in the library's mirror of the type
! (fitpack_b200/csrc/fpb_curve.cpp: sort of the data, nest rule, default smoothing trajectory, iopt promotion, bc
! default as in the reference).  Components the reference exposes as public data (t, knots, c, ...) are reached through
! knots() / get_t() / get_c().
!
! NOTE: the build container has no Fortran compiler: this file is shipped as source only, like fitpack_b200.f90
! (not compiled or tested here).  It contains no arithmetic.
module fitpack_b200_curves
   use, intrinsic :: iso_c_binding
   implicit none
   private

   integer, parameter, public :: FP_REAL = c_double
   integer, parameter, public :: FP_SIZE = c_int32_t
   integer, parameter, public :: FP_FLAG = c_int32_t
   integer, parameter, public :: FP_BOOL = c_bool

   !> The opaque handle of include/fitpack_curves_c.h:37-42
   type, bind(C) :: fitpack_curve_c
      type(c_ptr) :: cptr = c_null_ptr
   end type fitpack_curve_c

   type, public :: fitpack_curve_b200
      type(fitpack_curve_c) :: h
   contains
      procedure :: destroy
      procedure :: new_points
      procedure :: new_fit
      procedure :: fit
      procedure :: interpolate
      procedure :: least_squares
      procedure, private :: eval_one
      procedure, private :: eval_many
      generic   :: eval => eval_one, eval_many
      procedure :: dfdx
      procedure :: dfdx_all
      procedure :: integral
      procedure :: fourier_coefficients
      procedure :: smoothing
      procedure :: mse
      procedure :: degree
      procedure :: set_bc
      procedure :: get_bc
      procedure :: knots
      procedure :: get_t
      procedure :: get_c
      procedure :: comm_size
      procedure :: comm_pack
      procedure :: comm_expand
      procedure, private :: assign_curve
      generic   :: assignment(=) => assign_curve      ! deep copy through fitpack_curve_c_copy (the handle owns the data)
      final     :: finalize
   end type fitpack_curve_b200

   interface
      subroutine fitpack_curve_c_allocate(self) bind(C,name="fitpack_curve_c_allocate")
         import
         type(fitpack_curve_c), intent(inout) :: self
      end subroutine fitpack_curve_c_allocate
      subroutine fitpack_curve_c_destroy(self) bind(C,name="fitpack_curve_c_destroy")
         import
         type(fitpack_curve_c), intent(inout) :: self
      end subroutine fitpack_curve_c_destroy
      subroutine fitpack_curve_c_copy(self,other) bind(C,name="fitpack_curve_c_copy")
         import
         type(fitpack_curve_c), intent(inout) :: self
         type(fitpack_curve_c), intent(in)    :: other
      end subroutine fitpack_curve_c_copy
      subroutine fitpack_curve_c_new_points(self,npts,x,y,w) bind(C,name="fitpack_curve_c_new_points")
         import
         type(fitpack_curve_c), intent(inout) :: self
         integer(FP_SIZE), intent(in), value  :: npts
         real(FP_REAL), intent(in)            :: x(npts),y(npts)
         type(c_ptr), value                   :: w          ! optional weights: c_null_ptr = none
      end subroutine fitpack_curve_c_new_points
      integer(FP_FLAG) function fitpack_curve_c_new_fit(self,npts,x,y,w,smoothing) bind(C,name="fitpack_curve_c_new_fit")
         import
         type(fitpack_curve_c), intent(inout) :: self
         integer(FP_SIZE), intent(in), value  :: npts
         real(FP_REAL), intent(in)            :: x(npts),y(npts)
         type(c_ptr), value                   :: w,smoothing ! optional: c_null_ptr = absent
      end function fitpack_curve_c_new_fit
      integer(FP_FLAG) function fitpack_curve_c_fit(self,smoothing,order) bind(C,name="fitpack_curve_c_fit")
         import
         type(fitpack_curve_c), intent(inout) :: self
         type(c_ptr), value                   :: smoothing,order
      end function fitpack_curve_c_fit
      integer(FP_FLAG) function fitpack_curve_c_interpolating(self,order) bind(C,name="fitpack_curve_c_interpolating")
         import
         type(fitpack_curve_c), intent(inout) :: self
         type(c_ptr), value                   :: order
      end function fitpack_curve_c_interpolating
      integer(FP_FLAG) function fitpack_curve_c_least_squares(self,smoothing,reset_knots) &
                                bind(C,name="fitpack_curve_c_least_squares")
         import
         type(fitpack_curve_c), intent(inout) :: self
         type(c_ptr), value                   :: smoothing,reset_knots
      end function fitpack_curve_c_least_squares
      real(FP_REAL) function fitpack_curve_c_eval_one(self,x,ierr) bind(C,name="fitpack_curve_c_eval_one")
         import
         type(fitpack_curve_c), intent(inout) :: self
         real(FP_REAL), intent(in), value     :: x
         type(c_ptr), value                   :: ierr
      end function fitpack_curve_c_eval_one
      subroutine fitpack_curve_c_eval_many(self,npts,x,y,ierr) bind(C,name="fitpack_curve_c_eval_many")
         import
         type(fitpack_curve_c), intent(inout) :: self
         integer(FP_SIZE), intent(in), value  :: npts
         real(FP_REAL), intent(in)            :: x(npts)
         real(FP_REAL), intent(out)           :: y(npts)
         type(c_ptr), value                   :: ierr
      end subroutine fitpack_curve_c_eval_many
      real(FP_REAL) function fitpack_curve_c_integral(self,from,to) bind(C,name="fitpack_curve_c_integral")
         import
         type(fitpack_curve_c), intent(inout) :: self
         real(FP_REAL), intent(in), value     :: from,to
      end function fitpack_curve_c_integral
      subroutine fitpack_curve_c_fourier(self,nparm,alpha,A,B,ierr) bind(C,name="fitpack_curve_c_fourier")
         import
         type(fitpack_curve_c), intent(inout) :: self
         integer(FP_SIZE), intent(in), value  :: nparm
         real(FP_REAL), intent(in)            :: alpha(nparm)
         real(FP_REAL), intent(out)           :: A(nparm),B(nparm)
         type(c_ptr), value                   :: ierr
      end subroutine fitpack_curve_c_fourier
      real(FP_REAL) function fitpack_curve_c_derivative(self,x,order,ierr) bind(C,name="fitpack_curve_c_derivative")
         import
         type(fitpack_curve_c), intent(inout) :: self
         real(FP_REAL), intent(in), value     :: x
         integer(FP_SIZE), intent(in), value  :: order
         type(c_ptr), value                   :: ierr
      end function fitpack_curve_c_derivative
      integer(FP_FLAG) function fitpack_curve_c_all_derivatives(self,x,ddx) bind(C,name="fitpack_curve_c_all_derivatives")
         import
         type(fitpack_curve_c), intent(inout) :: self
         real(FP_REAL), intent(in), value     :: x
         real(FP_REAL), intent(out)           :: ddx(*)
      end function fitpack_curve_c_all_derivatives
      real(FP_REAL) function fitpack_curve_c_smoothing(self) bind(C,name="fitpack_curve_c_smoothing")
         import
         type(fitpack_curve_c), intent(inout) :: self
      end function fitpack_curve_c_smoothing
      real(FP_REAL) function fitpack_curve_c_mse(self) bind(C,name="fitpack_curve_c_mse")
         import
         type(fitpack_curve_c), intent(inout) :: self
      end function fitpack_curve_c_mse
      integer(FP_SIZE) function fitpack_curve_c_degree(self) bind(C,name="fitpack_curve_c_degree")
         import
         type(fitpack_curve_c), intent(inout) :: self
      end function fitpack_curve_c_degree
      subroutine fitpack_curve_c_set_bc(self,bc) bind(C,name="fitpack_curve_c_set_bc")
         import
         type(fitpack_curve_c), intent(inout) :: self
         integer(FP_FLAG), intent(in), value  :: bc
      end subroutine fitpack_curve_c_set_bc
      integer(FP_FLAG) function fitpack_curve_c_get_bc(self) bind(C,name="fitpack_curve_c_get_bc")
         import
         type(fitpack_curve_c), intent(inout) :: self
      end function fitpack_curve_c_get_bc
      integer(FP_SIZE) function fitpack_curve_c_knots(self) bind(C,name="fitpack_curve_c_knots")
         import
         type(fitpack_curve_c), intent(inout) :: self
      end function fitpack_curve_c_knots
      integer(FP_SIZE) function fitpack_curve_c_get_t(self,nmax,t) bind(C,name="fitpack_curve_c_get_t")
         import
         type(fitpack_curve_c), intent(inout) :: self
         integer(FP_SIZE), intent(in), value  :: nmax
         real(FP_REAL), intent(out)           :: t(nmax)
      end function fitpack_curve_c_get_t
      integer(FP_SIZE) function fitpack_curve_c_get_c(self,nmax,c) bind(C,name="fitpack_curve_c_get_c")
         import
         type(fitpack_curve_c), intent(inout) :: self
         integer(FP_SIZE), intent(in), value  :: nmax
         real(FP_REAL), intent(out)           :: c(nmax)
      end function fitpack_curve_c_get_c
      integer(FP_SIZE) function fitpack_curve_c_comm_size(self) bind(C,name="fitpack_curve_c_comm_size")
         import
         type(fitpack_curve_c), intent(inout) :: self
      end function fitpack_curve_c_comm_size
      subroutine fitpack_curve_c_comm_pack(self,buffer) bind(C,name="fitpack_curve_c_comm_pack")
         import
         type(fitpack_curve_c), intent(inout) :: self
         real(FP_REAL), intent(out)           :: buffer(*)
      end subroutine fitpack_curve_c_comm_pack
      subroutine fitpack_curve_c_comm_expand(self,buffer) bind(C,name="fitpack_curve_c_comm_expand")
         import
         type(fitpack_curve_c), intent(inout) :: self
         real(FP_REAL), intent(in)            :: buffer(*)
      end subroutine fitpack_curve_c_comm_expand
   end interface

contains

   !> c_loc of an optional scalar / array, c_null_ptr when absent (the C-ABI's "nullable pointer = optional")
   type(c_ptr) function opt_real(v) result(p)
      real(FP_REAL), intent(in), optional, target :: v
      p = c_null_ptr
      if (present(v)) p = c_loc(v)
   end function opt_real
   type(c_ptr) function opt_reals(v) result(p)
      real(FP_REAL), intent(in), optional, target, contiguous :: v(:)
      p = c_null_ptr
      if (present(v)) p = c_loc(v)
   end function opt_reals
   type(c_ptr) function opt_int(v) result(p)
      integer(FP_SIZE), intent(in), optional, target :: v
      p = c_null_ptr
      if (present(v)) p = c_loc(v)
   end function opt_int

   subroutine destroy(this)
      class(fitpack_curve_b200), intent(inout) :: this
      if (c_associated(this%h%cptr)) call fitpack_curve_c_destroy(this%h)
      this%h%cptr = c_null_ptr
   end subroutine destroy

   subroutine finalize(this)
      type(fitpack_curve_b200), intent(inout) :: this
      call this%destroy()
   end subroutine finalize

   subroutine assign_curve(this,other)
      class(fitpack_curve_b200), intent(inout) :: this
      type(fitpack_curve_b200), intent(in)     :: other
      call fitpack_curve_c_copy(this%h,other%h)
   end subroutine assign_curve

   !> curve%new_points(x,y,w)  (src/fitpack_curves.f90:215-263)
   subroutine new_points(this,x,y,w)
      class(fitpack_curve_b200), intent(inout) :: this
      real(FP_REAL), intent(in), contiguous :: x(:),y(:)
      real(FP_REAL), intent(in), optional, contiguous, target :: w(:)
      call fitpack_curve_c_new_points(this%h,int(size(x),FP_SIZE),x,y,opt_reals(w))
   end subroutine new_points

   !> ierr = curve%new_fit(x,y,w,smoothing,order)  (:168-179)
   integer(FP_FLAG) function new_fit(this,x,y,w,smoothing,order) result(ierr)
      class(fitpack_curve_b200), intent(inout) :: this
      real(FP_REAL), intent(in), contiguous :: x(:),y(:)
      real(FP_REAL), intent(in), optional, contiguous, target :: w(:)
      real(FP_REAL), intent(in), optional, target :: smoothing
      integer(FP_SIZE), intent(in), optional, target :: order
      if (present(order)) then
         call this%new_points(x,y,w)
         ierr = this%fit(smoothing,order)
      else
         ierr = fitpack_curve_c_new_fit(this%h,int(size(x),FP_SIZE),x,y,opt_reals(w),opt_real(smoothing))
      end if
   end function new_fit

   !> ierr = curve%fit(smoothing,order)  (:431-488)
   integer(FP_FLAG) function fit(this,smoothing,order) result(ierr)
      class(fitpack_curve_b200), intent(inout) :: this
      real(FP_REAL), intent(in), optional, target :: smoothing
      integer(FP_SIZE), intent(in), optional, target :: order
      ierr = fitpack_curve_c_fit(this%h,opt_real(smoothing),opt_int(order))
   end function fit

   !> ierr = curve%interpolate(order)  (:396-408)
   integer(FP_FLAG) function interpolate(this,order) result(ierr)
      class(fitpack_curve_b200), intent(inout) :: this
      integer(FP_SIZE), intent(in), optional, target :: order
      ierr = fitpack_curve_c_interpolating(this%h,opt_int(order))
   end function interpolate

   !> ierr = curve%least_squares(smoothing)  (:411-428)
   integer(FP_FLAG) function least_squares(this,smoothing) result(ierr)
      class(fitpack_curve_b200), intent(inout) :: this
      real(FP_REAL), intent(in), optional, target :: smoothing
      ierr = fitpack_curve_c_least_squares(this%h,opt_real(smoothing),c_null_ptr)
   end function least_squares

   !> y = curve%eval(x,ierr)  (:273-393)
   real(FP_REAL) function eval_one(this,x,ierr) result(y)
      class(fitpack_curve_b200), intent(inout) :: this
      real(FP_REAL), intent(in) :: x
      integer(FP_FLAG), intent(out), optional, target :: ierr
      type(c_ptr) :: pe
      pe = c_null_ptr
      if (present(ierr)) pe = c_loc(ierr)
      y = fitpack_curve_c_eval_one(this%h,x,pe)
   end function eval_one
   function eval_many(this,x,ierr) result(y)
      class(fitpack_curve_b200), intent(inout) :: this
      real(FP_REAL), intent(in), contiguous :: x(:)
      integer(FP_FLAG), intent(out), optional, target :: ierr
      real(FP_REAL) :: y(size(x))
      type(c_ptr) :: pe
      pe = c_null_ptr
      if (present(ierr)) pe = c_loc(ierr)
      call fitpack_curve_c_eval_many(this%h,int(size(x),FP_SIZE),x,y,pe)
   end function eval_many

   !> curve%dfdx(x,order,ierr), curve%dfdx_all(x,ierr)  (splder / spalde)
   real(FP_REAL) function dfdx(this,x,order,ierr) result(d)
      class(fitpack_curve_b200), intent(inout) :: this
      real(FP_REAL), intent(in) :: x
      integer(FP_SIZE), intent(in) :: order
      integer(FP_FLAG), intent(out), optional, target :: ierr
      type(c_ptr) :: pe
      pe = c_null_ptr
      if (present(ierr)) pe = c_loc(ierr)
      d = fitpack_curve_c_derivative(this%h,x,order,pe)
   end function dfdx
   function dfdx_all(this,x,ierr) result(ddx)
      class(fitpack_curve_b200), intent(inout) :: this
      real(FP_REAL), intent(in) :: x
      integer(FP_FLAG), intent(out), optional :: ierr
      real(FP_REAL), allocatable :: ddx(:)
      integer(FP_FLAG) :: ier
      allocate(ddx(this%degree()+1))
      ier = fitpack_curve_c_all_derivatives(this%h,x,ddx)
      if (present(ierr)) ierr = ier
   end function dfdx_all

   !> curve%integral(from,to)  (splint)
   real(FP_REAL) function integral(this,from,to) result(v)
      class(fitpack_curve_b200), intent(inout) :: this
      real(FP_REAL), intent(in) :: from,to
      v = fitpack_curve_c_integral(this%h,from,to)
   end function integral

   !> call curve%fourier_coefficients(alpha,A,B,ierr)  (:650-668, fourco)
   subroutine fourier_coefficients(this,alpha,A,B,ierr)
      class(fitpack_curve_b200), intent(inout) :: this
      real(FP_REAL), intent(in), contiguous :: alpha(:)
      real(FP_REAL), intent(out), contiguous :: A(:),B(:)
      integer(FP_FLAG), intent(out), optional, target :: ierr
      type(c_ptr) :: pe
      pe = c_null_ptr
      if (present(ierr)) pe = c_loc(ierr)
      call fitpack_curve_c_fourier(this%h,int(size(alpha),FP_SIZE),alpha,A,B,pe)
   end subroutine fourier_coefficients

   real(FP_REAL) function smoothing(this) result(s)
      class(fitpack_curve_b200), intent(inout) :: this
      s = fitpack_curve_c_smoothing(this%h)
   end function smoothing
   real(FP_REAL) function mse(this) result(v)
      class(fitpack_curve_b200), intent(inout) :: this
      v = fitpack_curve_c_mse(this%h)
   end function mse
   integer(FP_SIZE) function degree(this) result(k)
      class(fitpack_curve_b200), intent(inout) :: this
      k = fitpack_curve_c_degree(this%h)
   end function degree
   subroutine set_bc(this,bc)
      class(fitpack_curve_b200), intent(inout) :: this
      integer(FP_FLAG), intent(in) :: bc
      call fitpack_curve_c_set_bc(this%h,bc)
   end subroutine set_bc
   integer(FP_FLAG) function get_bc(this) result(bc)
      class(fitpack_curve_b200), intent(inout) :: this
      bc = fitpack_curve_c_get_bc(this%h)
   end function get_bc

   !> number of knots, knots t(1:n) and coefficients c(1:n) of the current fit
   integer(FP_SIZE) function knots(this) result(n)
      class(fitpack_curve_b200), intent(inout) :: this
      n = fitpack_curve_c_knots(this%h)
   end function knots
   function get_t(this) result(t)
      class(fitpack_curve_b200), intent(inout) :: this
      real(FP_REAL), allocatable :: t(:)
      integer(FP_SIZE) :: n
      n = this%knots()
      allocate(t(max(n,1)))
      n = fitpack_curve_c_get_t(this%h,int(size(t),FP_SIZE),t)
   end function get_t
   function get_c(this) result(c)
      class(fitpack_curve_b200), intent(inout) :: this
      real(FP_REAL), allocatable :: c(:)
      integer(FP_SIZE) :: n
      n = this%knots()
      allocate(c(max(n,1)))
      n = fitpack_curve_c_get_c(this%h,int(size(c),FP_SIZE),c)
   end function get_c

   !> parallel communication (src/fitpack_curves.f90:821-895): same wire format as the reference's objects
   integer(FP_SIZE) function comm_size(this) result(n)
      class(fitpack_curve_b200), intent(inout) :: this
      n = fitpack_curve_c_comm_size(this%h)
   end function comm_size
   subroutine comm_pack(this,buffer)
      class(fitpack_curve_b200), intent(inout) :: this
      real(FP_REAL), intent(out) :: buffer(:)
      call fitpack_curve_c_comm_pack(this%h,buffer)
   end subroutine comm_pack
   subroutine comm_expand(this,buffer)
      class(fitpack_curve_b200), intent(inout) :: this
      real(FP_REAL), intent(in) :: buffer(:)
      call fitpack_curve_c_comm_expand(this%h,buffer)
   end subroutine comm_expand

end module fitpack_b200_curves
