! fitpack_b200.f90 -- iso_c_binding interface module for libfitpack_b200.so
!
! Host code stays Modern Fortran (BASELINE.json north_star): this module is the thin shim
! through which it reaches the CUDA kernels.  It binds
!   (A) the reference's own C symbols for the path (same names/arguments as
!       src/fitpack_core_c.f90:64-74,145-154 of perazz/fitpack), and
!   (B) the new batch entry points of include/fitpack_b200.h.
! It contains no arithmetic.  NOTE: the build container has no Fortran compiler, so this file
! is shipped as source only (not compiled or tested here); INTEGRATION.md shows how the
! reference's fitpack_curve%fit / %eval call sites switch over to it.
module fitpack_b200
   use, intrinsic :: iso_c_binding
   implicit none
   private

   integer, parameter, public :: FP_REAL = c_double
   integer, parameter, public :: FP_SIZE = c_int32_t
   integer, parameter, public :: FP_FLAG = c_int32_t
   integer, parameter, public :: FP_BOOL = c_bool

   integer(FP_FLAG), parameter, public :: FPB_SPLEV_EXACT = 0, FPB_SPLEV_FAST = 1
   integer(c_int32_t), parameter, public :: FPB_STATUS_OK = 0, FPB_ERR_CUDA = -100, &
                                            FPB_ERR_ARG = -101, FPB_ERR_ALLOC = -102

   public :: fp_curfit_c, fp_splev_c, fp_curfit_batch_c, fp_splev_batch_c
   public :: fp_regrid_c, fp_bispev_c, fp_splder_c, fp_spalde_c, fp_splint_c, fp_fourco_c, fp_cualde_c, fp_insert_c, fp_sproot_c
   public :: fp_parcur_c, fp_curev_c, fp_parcur_batch_c, fp_percur_c, fp_percur_batch_c
   public :: fitpack_curve_c_comm_size, fitpack_curve_c_comm_pack, fitpack_curve_c_comm_expand
   public :: fp_curfit_batch_dev_c, fp_curfit_shared_dev_c, fp_splev_batch_dev_c
   public :: fpb_engine_create, fpb_engine_destroy, fpb_engine_synchronize, fpb_device_count
   public :: curfit_batch, splev_batch

   interface

      ! (A) drop-in: same interface as the reference's curfit_c / splev_c
      subroutine fp_curfit_c(iopt,m,x,y,w,xb,xe,k,s,nest,n,t,c,fp,wrk,lwrk,iwrk,ier) bind(C,name="fp_curfit_c")
         import
         real   (FP_REAL), intent(in),    value :: xb,xe,s
         real   (FP_REAL), intent(inout)        :: fp
         integer(FP_SIZE), intent(in),    value :: iopt,m,k,nest,lwrk
         integer(FP_FLAG), intent(out)          :: ier
         integer(FP_SIZE), intent(inout)        :: n
         real   (FP_REAL), intent(in)           :: x(m),y(m),w(m)
         real   (FP_REAL), intent(inout)        :: t(nest),c(nest),wrk(lwrk)
         integer(FP_SIZE), intent(inout)        :: iwrk(nest)
      end subroutine fp_curfit_c

      subroutine fp_splev_c(t,n,c,k,x,y,m,e,ier) bind(C,name="fp_splev_c")
         import
         integer(FP_SIZE), intent(in), value :: n,k,m
         real   (FP_REAL), intent(in)        :: t(n),c(n),x(m)
         real   (FP_REAL), intent(out)       :: y(m)
         integer(FP_FLAG), intent(in), value :: e
         integer(FP_FLAG), intent(out)       :: ier
      end subroutine fp_splev_c

      ! gridded surfaces and spline calculus: same interfaces as the reference's regrid_c, bispev_c,
      ! splder_c, spalde_c, splint_c (src/fitpack_core_c.f90:253-290, 404-415, 157-175, 207-213)
      subroutine fp_regrid_c(iopt,mx,x,my,y,z,xb,xe,yb,ye,kx,ky,s,nxest,nyest, &
                             nx,tx,ny,ty,c,fp,wrk,lwrk,iwrk,kwrk,ier) bind(C,name="fp_regrid_c")
         import
         real   (FP_REAL), intent(in), value :: xb,xe,yb,ye,s
         real   (FP_REAL), intent(out)       :: fp
         integer(FP_SIZE), intent(in), value :: iopt,mx,my,kx,ky,nxest,nyest,lwrk,kwrk
         integer(FP_SIZE), intent(inout)     :: nx,ny
         integer(FP_FLAG), intent(out)       :: ier
         real   (FP_REAL), intent(in)        :: x(mx),y(my),z(mx*my)
         real   (FP_REAL), intent(inout)     :: tx(nxest),ty(nyest),c((nxest-kx-1)*(nyest-ky-1)),wrk(lwrk)
         integer(FP_SIZE), intent(inout)     :: iwrk(kwrk)
      end subroutine fp_regrid_c

      integer(FP_FLAG) function fp_bispev_c(tx,nx,ty,ny,c,kx,ky,x,mx,y,my,z,wrk,lwrk,iwrk,kwrk) &
                                bind(C,name="fp_bispev_c")
         import
         integer(FP_SIZE), intent(in), value :: nx,ny,kx,ky,mx,my,lwrk,kwrk
         integer(FP_SIZE), intent(inout)     :: iwrk(kwrk)
         real   (FP_REAL), intent(in)        :: tx(nx),ty(ny),c((nx-kx-1)*(ny-ky-1)),x(mx),y(my)
         real   (FP_REAL), intent(out)       :: z(mx*my)
         real   (FP_REAL), intent(inout)     :: wrk(lwrk)
      end function fp_bispev_c

      ! parametric curves: same interfaces as the reference's parcur_c / curev_c
      ! (src/fitpack_core_c.f90:90-100, 178-184); fp_parcur_batch_c is the batch form (new)
      subroutine fp_parcur_c(iopt,ipar,idim,m,u,mx,x,w,ub,ue,k,s,nest,n,t,nc,c,fp,wrk,lwrk,iwrk,ier) &
                             bind(C,name="fp_parcur_c")
         import
         real   (FP_REAL), intent(inout)       :: ub,ue,s
         real   (FP_REAL), intent(out)         :: fp
         integer(FP_SIZE), intent(in),   value :: iopt,ipar,idim,m,mx,k,nest,lwrk,nc
         integer(FP_SIZE), intent(inout)       :: n
         integer(FP_FLAG), intent(out)         :: ier
         real   (FP_REAL), intent(in)          :: x(idim,m)
         real   (FP_REAL), intent(inout)       :: u(m),w(m),t(nest),c(nc),wrk(lwrk)
         integer(FP_SIZE), intent(inout)       :: iwrk(nest)
      end subroutine fp_parcur_c

      subroutine fp_curev_c(idim,t,n,c,nc,k,u,m,x,mx,ier) bind(C,name="fp_curev_c")
         import
         integer(FP_SIZE), intent(in), value :: idim,n,nc,k,m,mx
         integer(FP_FLAG), intent(out)       :: ier
         real   (FP_REAL), intent(in)        :: t(n),c(nc),u(m)
         real   (FP_REAL), intent(out)       :: x(mx)
      end subroutine fp_curev_c

      ! x(idim,m,nbatch), u(m,nbatch), w(m,nbatch) (c_null_ptr-style optional: pass w as a
      ! type(c_ptr) interface variant if unit weights are wanted), c(nest*idim,nbatch)
      integer(c_int32_t) function fp_parcur_batch_c(nbatch,iopt,ipar,idim,m,u,x,w,ub,ue,k,s,nest, &
                                                    n,t,c,fp,ier,fpint,nrdata,ngpus) bind(C,name="fp_parcur_batch_c")
         import
         integer(FP_SIZE), intent(in), value :: nbatch,iopt,ipar,idim,m,k,nest,ngpus
         real   (FP_REAL), intent(in), value :: s
         real   (FP_REAL), intent(inout)     :: u(m,nbatch),ub(nbatch),ue(nbatch)
         real   (FP_REAL), intent(in)        :: x(idim,m,nbatch),w(m,nbatch)
         integer(FP_SIZE), intent(inout)     :: n(nbatch),nrdata(nest,nbatch)
         real   (FP_REAL), intent(inout)     :: t(nest,nbatch),c(nest*idim,nbatch),fp(nbatch),fpint(nest,nbatch)
         integer(FP_FLAG), intent(out)       :: ier(nbatch)
      end function fp_parcur_batch_c

      ! periodic curves: same interface as the reference's percur_c (src/fitpack_core_c.f90:77-87);
      ! fp_percur_batch_c is the batch form (new)
      subroutine fp_percur_c(iopt,m,x,y,w,k,s,nest,n,t,c,fp,wrk,lwrk,iwrk,ier) bind(C,name="fp_percur_c")
         import
         real   (FP_REAL), intent(in),    value :: s
         real   (FP_REAL), intent(inout)        :: fp
         integer(FP_SIZE), intent(inout)        :: n
         integer(FP_FLAG), intent(inout)        :: ier
         integer(FP_SIZE), intent(in),    value :: iopt,m,k,nest,lwrk
         real   (FP_REAL), intent(in)           :: x(m),y(m),w(m)
         real   (FP_REAL), intent(inout)        :: t(nest),c(nest),wrk(lwrk)
         integer(FP_SIZE), intent(inout)        :: iwrk(nest)
      end subroutine fp_percur_c

      integer(c_int32_t) function fp_percur_batch_c(nbatch,iopt,m,x,y,w,k,s,nest,n,t,c,fp,ier,fpint,nrdata,ngpus) &
                                                    bind(C,name="fp_percur_batch_c")
         import
         integer(FP_SIZE), intent(in), value :: nbatch,iopt,m,k,nest,ngpus
         real   (FP_REAL), intent(in), value :: s
         real   (FP_REAL), intent(in)        :: x(m,nbatch),y(m,nbatch),w(m,nbatch)
         integer(FP_SIZE), intent(inout)     :: n(nbatch),nrdata(nest,nbatch)
         real   (FP_REAL), intent(inout)     :: t(nest,nbatch),c(nest,nbatch),fp(nbatch),fpint(nest,nbatch)
         integer(FP_FLAG), intent(out)       :: ier(nbatch)
      end function fp_percur_batch_c

      ! comm buffers of a curve handle, same wire format as curve%comm_pack / comm_expand
      ! (src/fitpack_curves.f90:821-895): `self` is the type(c_ptr) wrapper struct of the handle ABI
      integer(FP_SIZE) function fitpack_curve_c_comm_size(self) bind(C,name="fitpack_curve_c_comm_size")
         import
         type(c_ptr), intent(inout) :: self
      end function fitpack_curve_c_comm_size
      subroutine fitpack_curve_c_comm_pack(self,buffer) bind(C,name="fitpack_curve_c_comm_pack")
         import
         type(c_ptr), intent(inout) :: self
         real(FP_REAL), intent(out) :: buffer(*)
      end subroutine fitpack_curve_c_comm_pack
      subroutine fitpack_curve_c_comm_expand(self,buffer) bind(C,name="fitpack_curve_c_comm_expand")
         import
         type(c_ptr), intent(inout) :: self
         real(FP_REAL), intent(in)  :: buffer(*)
      end subroutine fitpack_curve_c_comm_expand

      subroutine fp_splder_c(t,n,c,k,nu,x,y,m,e,wrk,ier) bind(C,name="fp_splder_c")
         import
         integer(FP_SIZE), intent(in), value :: n,k,nu,m
         integer(FP_FLAG), intent(in), value :: e
         integer(FP_FLAG), intent(out)       :: ier
         real   (FP_REAL), intent(in)        :: t(n),c(n),x(m)
         real   (FP_REAL), intent(out)       :: y(m)
         real   (FP_REAL), intent(inout)     :: wrk(n)
      end subroutine fp_splder_c

      subroutine fp_spalde_c(t,n,c,k1,x,d,ier) bind(C,name="fp_spalde_c")
         import
         integer(FP_SIZE), intent(in), value :: n,k1
         integer(FP_FLAG), intent(out)       :: ier
         real   (FP_REAL), intent(in), value :: x
         real   (FP_REAL), intent(in)        :: t(n),c(n)
         real   (FP_REAL), intent(out)       :: d(k1)
      end subroutine fp_spalde_c

      real(FP_REAL) function fp_splint_c(t,n,c,k,a,b,wrk) bind(C,name="fp_splint_c")
         import
         real   (FP_REAL), intent(in), value :: a,b
         integer(FP_SIZE), intent(in), value :: n,k
         real   (FP_REAL), intent(in)        :: t(n),c(n)
         real   (FP_REAL), intent(inout)     :: wrk(n)
      end function fp_splint_c

      ! Fourier integrals of a cubic spline (same argument list as the reference's fp_fourco_c)
      subroutine fp_fourco_c(t,n,c,alfa,m,ress,resc,wrk1,wrk2,ier) bind(C,name="fp_fourco_c")
         import
         integer(FP_SIZE), intent(in), value :: n,m
         integer(FP_FLAG), intent(out)       :: ier
         real   (FP_REAL), intent(in)        :: t(n),c(n),alfa(m)
         real   (FP_REAL), intent(inout)     :: wrk1(n),wrk2(n)
         real   (FP_REAL), intent(out)       :: ress(m),resc(m)
      end subroutine fp_fourco_c

      ! all derivatives of a parametric curve at u / knot insertion in place (the reference's argument lists)
      subroutine fp_cualde_c(idim,t,n,c,nc,k1,u,d,nd,ier) bind(C,name="fp_cualde_c")
         import
         integer(FP_SIZE), intent(in), value :: idim,n,nc,k1,nd
         integer(FP_FLAG), intent(out)       :: ier
         real   (FP_REAL), intent(in), value :: u
         real   (FP_REAL), intent(in)        :: t(n),c(nc)
         real   (FP_REAL), intent(out)       :: d(nd)
      end subroutine fp_cualde_c
      subroutine fp_insert_c(iopt,t,n,c,k,x,nest,ier) bind(C,name="fp_insert_c")
         import
         integer(FP_SIZE), intent(in), value :: iopt,k,nest
         real   (FP_REAL), intent(inout)     :: t(nest),c(nest)
         real   (FP_REAL), intent(in), value :: x
         integer(FP_SIZE), intent(inout)     :: n
         integer(FP_FLAG), intent(out)       :: ier
      end subroutine fp_insert_c
      subroutine fp_sproot_c(t,n,c,zeros,mest,m,ier) bind(C,name="fp_sproot_c")
         import
         integer(FP_SIZE), intent(in), value :: n,mest
         integer(FP_SIZE), intent(out)       :: m
         integer(FP_FLAG), intent(out)       :: ier
         real   (FP_REAL), intent(in)        :: t(n),c(n)
         real   (FP_REAL), intent(out)       :: zeros(mest)
      end subroutine fp_sproot_c

      ! (B) batches of independent curves, host arrays.  Fortran arrays x(m,nbatch) are exactly
      ! the "curve-major" [nbatch][m] C layout the library expects.
      integer(c_int32_t) function fp_curfit_batch_c(nbatch,iopt,m,x,y,w,shared_x,xb,xe,k,s,nest, &
                                                    n,t,c,fp,ier,ngpus) bind(C,name="fp_curfit_batch_c")
         import
         integer(FP_SIZE), intent(in), value :: nbatch,iopt,m,shared_x,k,nest,ngpus
         real   (FP_REAL), intent(in)        :: x(*),y(m,nbatch)
         type(c_ptr),      intent(in), value :: w          ! c_null_ptr => unit weights
         type(c_ptr),      intent(in), value :: xb,xe      ! c_null_ptr => data range per curve
         real   (FP_REAL), intent(in), value :: s
         integer(FP_SIZE), intent(inout)     :: n(nbatch)
         real   (FP_REAL), intent(inout)     :: t(nest,nbatch),c(nest,nbatch),fp(nbatch)
         integer(FP_FLAG), intent(out)       :: ier(nbatch)
      end function fp_curfit_batch_c

      integer(c_int32_t) function fp_splev_batch_c(nspl,t,n,c,ldt,k,x,y,m,e,mode,ier,ngpus) &
                                                   bind(C,name="fp_splev_batch_c")
         import
         integer(FP_SIZE), intent(in), value :: nspl,ldt,k,m,mode,ngpus
         integer(FP_FLAG), intent(in), value :: e
         real   (FP_REAL), intent(in)        :: t(ldt,nspl),c(ldt,nspl),x(m,nspl)
         integer(FP_SIZE), intent(in)        :: n(nspl)
         real   (FP_REAL), intent(out)       :: y(m,nspl)
         integer(FP_FLAG), intent(out)       :: ier(nspl)
      end function fp_splev_batch_c

      ! device-pointer variants (type(c_ptr) = device addresses, e.g. from cudaMalloc / OpenACC
      ! host_data / CUDA Fortran c_devloc); see include/fitpack_b200.h for the stride convention
      integer(c_int32_t) function fp_curfit_batch_dev_c(eng,nbatch,iopt,m,x,x_sb,x_si,y,y_sb,y_si, &
                              w,w_sb,w_si,xb,xe,xbe_stride,k,s,s_stride,nest,n,t,c,fp,ier,       &
                              fpint,nrdata,stats) bind(C,name="fp_curfit_batch_dev_c")
         import
         type(c_ptr),        intent(in), value :: eng,x,y,w,xb,xe,s,n,t,c,fp,ier,fpint,nrdata,stats
         integer(FP_SIZE),   intent(in), value :: nbatch,iopt,m,xbe_stride,k,s_stride,nest
         integer(c_int64_t), intent(in), value :: x_sb,x_si,y_sb,y_si,w_sb,w_si
      end function fp_curfit_batch_dev_c

      integer(c_int32_t) function fp_curfit_shared_dev_c(eng,nbatch,m,x,w,y,y_sb,y_si,xb,xe,k,n,t, &
                              c,ldc,fp,ier) bind(C,name="fp_curfit_shared_dev_c")
         import
         type(c_ptr),        intent(in), value :: eng,x,w,y,t,c,fp,ier
         integer(FP_SIZE),   intent(in), value :: nbatch,m,k,n
         integer(c_int64_t), intent(in), value :: y_sb,y_si,ldc
         real(FP_REAL),      intent(in), value :: xb,xe
      end function fp_curfit_shared_dev_c

      integer(c_int32_t) function fp_splev_batch_dev_c(eng,nspl,t,n,c,ldt,k,x,y,m,e,mode,nshared, &
                              ier,lidx) bind(C,name="fp_splev_batch_dev_c")
         import
         type(c_ptr),      intent(in), value :: eng,t,n,c,x,y,ier,lidx
         integer(FP_SIZE), intent(in), value :: nspl,ldt,k,m,mode,nshared
         integer(FP_FLAG), intent(in), value :: e
      end function fp_splev_batch_dev_c

      integer(c_int32_t) function fpb_engine_create(device,eng) bind(C,name="fpb_engine_create")
         import
         integer(c_int32_t), intent(in), value :: device
         type(c_ptr), intent(out) :: eng
      end function fpb_engine_create

      subroutine fpb_engine_destroy(eng) bind(C,name="fpb_engine_destroy")
         import
         type(c_ptr), intent(in), value :: eng
      end subroutine fpb_engine_destroy

      integer(c_int32_t) function fpb_engine_synchronize(eng) bind(C,name="fpb_engine_synchronize")
         import
         type(c_ptr), intent(in), value :: eng
      end function fpb_engine_synchronize

      ! arithmetic of fp_curev_batch_dev_c: 0 exact (reference arithmetic), 1 fast (local polynomials + Horner)
      integer(c_int32_t) function fpb_engine_set_curev_mode(eng,mode) bind(C,name="fpb_engine_set_curev_mode")
         import :: c_ptr, c_int32_t
         type(c_ptr), value :: eng
         integer(c_int32_t), value :: mode
      end function fpb_engine_set_curev_mode
      ! batch of one from host memory: 1 warp-per-curve kernel (default), 0 batch kernels; returns the previous setting
      integer(c_int32_t) function fpb_set_single_curve_kernel(on) bind(C,name="fpb_set_single_curve_kernel")
         import :: c_int32_t
         integer(c_int32_t), value :: on
      end function fpb_set_single_curve_kernel
      ! ranks / slices assumed to share this host (LOCAL_WORLD_SIZE, ... or FPB_LOCAL_RANKS)
      integer(c_int32_t) function fpb_host_sharers() bind(C,name="fpb_host_sharers")
         import :: c_int32_t
      end function fpb_host_sharers
      ! form of the host-pointer batch fit for nbatch curves per GPU slice (C string: "warp-per-curve", "wavefront", "chunks")
      type(c_ptr) function fpb_host_form(nbatch) bind(C,name="fpb_host_form")
         import :: c_ptr, c_int64_t
         integer(c_int64_t), value :: nbatch
      end function fpb_host_form
      integer(c_int32_t) function fpb_engine_set_bispev_mode(eng,mode) bind(C,name="fpb_engine_set_bispev_mode")
         import :: c_ptr, c_int32_t
         type(c_ptr), value :: eng
         integer(c_int32_t), value :: mode
      end function fpb_engine_set_bispev_mode
      ! kernels launched so far by the default engines of a device (the ones the host-pointer calls run on)
      integer(c_int64_t) function fpb_default_engine_launch_count(device) bind(C,name="fpb_default_engine_launch_count")
         import :: c_int32_t, c_int64_t
         integer(c_int32_t), value :: device
      end function fpb_default_engine_launch_count
      integer(c_int32_t) function fpb_device_count() bind(C,name="fpb_device_count")
         import
      end function fpb_device_count

   end interface

contains

   !> Fit nbatch independent curves y(:,b) on abscissae x(:,b) (the batched analogue of
   !> curve%new_fit): same argument vocabulary as curfit (iopt, k, s, nest, n, t, c, fp, ier).
   integer(FP_FLAG) function curfit_batch(x,y,k,s,nest,n,t,c,fp,ier,w,ngpus) result(status)
      real(FP_REAL), intent(in), target   :: x(:,:), y(:,:)
      integer(FP_SIZE), intent(in)        :: k, nest
      real(FP_REAL), intent(in)           :: s
      integer(FP_SIZE), intent(inout)     :: n(:)
      real(FP_REAL), intent(inout)        :: t(:,:), c(:,:), fp(:)
      integer(FP_FLAG), intent(out)       :: ier(:)
      real(FP_REAL), intent(in), optional, target :: w(:,:)
      integer(FP_SIZE), intent(in), optional :: ngpus
      type(c_ptr) :: wp
      integer(FP_SIZE) :: ng
      wp = c_null_ptr; if (present(w)) wp = c_loc(w)
      ng = 0;          if (present(ngpus)) ng = ngpus
      status = fp_curfit_batch_c(size(y,2,FP_SIZE), 0_FP_SIZE, size(y,1,FP_SIZE), x, y, wp, &
                                 0_FP_SIZE, c_null_ptr, c_null_ptr, k, s, nest, n, t, c, fp, ier, ng)
   end function curfit_batch

   !> Evaluate nspl splines (t(:,j), n(j), c(:,j)) at x(:,j): batched analogue of curve%eval.
   integer(FP_FLAG) function splev_batch(t,n,c,k,x,y,e,ier,mode,ngpus) result(status)
      real(FP_REAL), intent(in)       :: t(:,:), c(:,:), x(:,:)
      integer(FP_SIZE), intent(in)    :: n(:), k
      real(FP_REAL), intent(out)      :: y(:,:)
      integer(FP_FLAG), intent(in)    :: e
      integer(FP_FLAG), intent(out)   :: ier(:)
      integer(FP_SIZE), intent(in), optional :: mode, ngpus
      integer(FP_SIZE) :: md, ng
      md = FPB_SPLEV_EXACT; if (present(mode)) md = mode
      ng = 0;               if (present(ngpus)) ng = ngpus
      status = fp_splev_batch_c(size(t,2,FP_SIZE), t, n, c, size(t,1,FP_SIZE), k, x, y, &
                                size(x,1,FP_SIZE), e, md, ier, ng)
   end function splev_batch

end module fitpack_b200
