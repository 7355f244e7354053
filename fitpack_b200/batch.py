"""Engine: batched device-pointer calls on torch CUDA tensors.

torch is plumbing only (device memory, streams, torch.distributed in bench.py); every
kernel launched here is one of this repo's own (fp_*_dev_c in include/fitpack_b200.h).
"""
import ctypes as C

import torch

from ._lib import check, lib


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _chk_f64(t, name):
    if t is not None and (t.dtype != torch.float64 or not t.is_cuda):
        raise TypeError("%s must be a CUDA float64 tensor" % name)


class Engine:
    """One per GPU: owns the stream-bound workspace of the kernels."""

    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("fitpack_b200 needs a CUDA device; there is no CPU fallback")
        idx = None if device is None else torch.device(device).index
        # an index-less "cuda" means the CURRENT device (one process per GPU sets it to its local rank)
        self.device = torch.device("cuda", torch.cuda.current_device() if idx is None else idx)
        h = C.c_void_p()
        check(lib().fpb_engine_create(self.device.index, C.byref(h)), "fpb_engine_create")
        self._h = h
        self.use_torch_stream()

    def close(self):
        if getattr(self, "_h", None):
            lib().fpb_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def use_torch_stream(self):
        s = torch.cuda.current_stream(self.device)
        check(lib().fpb_engine_set_stream(self._h, C.c_void_p(s.cuda_stream)), "set_stream")

    def synchronize(self):
        check(lib().fpb_engine_synchronize(self._h), "synchronize")

    @property
    def handle(self):
        """the fpb_engine* of the C-ABI (for the fitpack_b200.core calls that take engine=)"""
        return self._h

    @property
    def grid_fpgrre_calls(self):
        return lib().fpb_engine_grid_fpgrre_calls(self._h)

    @property
    def launches(self):
        return lib().fpb_engine_launch_count(self._h)

    def probe_fp64_gops(self, fused=False):
        """measured FP64 instruction rate (1e9/s): DMUL+DADD pairs (fused=False / 0), DFMA (True / 1),
        IEEE divisions (2) or IEEE square roots (3)"""
        self.use_torch_stream()
        return lib().fpb_probe_fp64_gops(self._h, int(fused))

    def last_kernel_ms(self):
        return lib().fpb_engine_last_kernel_ms(self._h)

    # ------------------------------------------------------------------ fit
    def curfit_batch(self, x, y, w=None, k=3, s=0.0, nest=None, iopt=0, xb=None, xe=None,
                     layout="point_major", n=None, t=None, fpint=None, nrdata=None,
                     want_stats=False, out=None):
        """Batched curfit on device tensors.

        layout "point_major": y is [m][B] (SoA, coalesced); "curve_major": y is [B][m].
        x / w may be 1-D [m] (shared by all curves) or shaped like y.  s: float or [B] tensor.
        xb/xe: None (data range of each curve), floats, or [B] tensors.
        Returns dict(n, t, c, fp, ier[, stats]) of device tensors (t, c: [B][nest]).
        """
        _chk_f64(y, "y"); _chk_f64(x, "x"); _chk_f64(w, "w")
        if layout == "point_major":
            m, B = y.shape
            sb, si = 1, y.stride(0)
        elif layout == "curve_major":
            B, m = y.shape
            sb, si = y.stride(0), 1
        else:
            raise ValueError("layout must be point_major or curve_major")

        def strides(a, name):
            if a is None:
                return 0, 0
            if a.dim() == 1:
                if a.numel() != m:
                    raise ValueError("shared %s must have m entries" % name)
                return 0, a.stride(0)
            if a.shape != y.shape:
                raise ValueError("%s must be [m] or shaped like y" % name)
            return (1, a.stride(0)) if layout == "point_major" else (a.stride(0), 1)

        x_sb, x_si = strides(x, "x")
        w_sb, w_si = strides(w, "w")
        nest = int(nest if nest is not None else m + k + 1)
        dev = y.device
        f64, i32 = torch.float64, torch.int32
        if torch.is_tensor(s):
            s_t, s_stride = s.to(device=dev, dtype=f64).contiguous(), 1
        else:
            s_t, s_stride = torch.full((1,), float(s), device=dev, dtype=f64), 0
        if xb is None:
            xb_t = xe_t = None
            xbe_stride = 0
        elif torch.is_tensor(xb):
            xb_t, xe_t, xbe_stride = xb.contiguous(), xe.contiguous(), 1
        else:
            xb_t = torch.full((1,), float(xb), device=dev, dtype=f64)
            xe_t = torch.full((1,), float(xe), device=dev, dtype=f64)
            xbe_stride = 0
        if out is None:
            out = {}
        n_t = n if n is not None else out.get("n")
        if n_t is None:
            n_t = torch.zeros(B, device=dev, dtype=i32)
        t_t = t if t is not None else out.get("t")
        if t_t is None:
            t_t = torch.zeros((B, nest), device=dev, dtype=f64)
        c_t = out.get("c")
        if c_t is None:
            c_t = torch.zeros((B, nest), device=dev, dtype=f64)
        fp_t = out.get("fp")
        if fp_t is None:
            fp_t = torch.zeros(B, device=dev, dtype=f64)
        ier_t = out.get("ier")
        if ier_t is None:
            ier_t = torch.zeros(B, device=dev, dtype=i32)
        stats_t = None
        if want_stats:
            stats_t = out.get("stats")
            if stats_t is None:
                stats_t = torch.zeros((B, 2), device=dev, dtype=i32)
        self.use_torch_stream()
        rc = lib().fp_curfit_batch_dev_c(
            self._h, B, int(iopt), m, _ptr(x), x_sb, x_si, _ptr(y), sb, si, _ptr(w), w_sb, w_si,
            _ptr(xb_t), _ptr(xe_t), xbe_stride, int(k), _ptr(s_t), s_stride, nest, _ptr(n_t),
            _ptr(t_t), _ptr(c_t), _ptr(fp_t), _ptr(ier_t), _ptr(fpint), _ptr(nrdata),
            _ptr(stats_t))
        check(rc, "fp_curfit_batch_dev_c")
        res = dict(n=n_t, t=t_t, c=c_t, fp=fp_t, ier=ier_t, _keep=(s_t, xb_t, xe_t))
        if want_stats:
            res["stats"] = stats_t
        return res

    def parcur_batch(self, x, w=None, k=3, s=0.0, nest=None, iopt=0, ipar=0, u=None, ub=None, ue=None,
                     n=None, t=None, fpint=None, nrdata=None, want_stats=False):
        """Batched parcur on device tensors, point-major: x is [m][idim][B] (coordinate d of point i
        of curve b at x[i, d, b]), u / w are [m][B].  ipar=0: u is computed on the device (returned).
        ub/ue: None (data range) or [B] tensors.  Returns dict(n, t, c, fp, ier, u[, stats]) with
        c [B][nest*idim] (coordinate d at c[b, d*n[b] : d*n[b]+n[b]-k-1])."""
        _chk_f64(x, "x"); _chk_f64(w, "w"); _chk_f64(u, "u")
        m, idim, B = x.shape
        if not x.is_contiguous():
            raise ValueError("x must be a contiguous [m][idim][B] tensor")
        dev = x.device
        f64, i32 = torch.float64, torch.int32
        if u is None:
            u = torch.zeros((m, B), device=dev, dtype=f64)
        nest = int(nest if nest is not None else m + k + 1)
        if torch.is_tensor(s):
            s_t, s_stride = s.to(device=dev, dtype=f64).contiguous(), 1
        else:
            s_t, s_stride = torch.full((1,), float(s), device=dev, dtype=f64), 0
        n_t = n if n is not None else torch.zeros(B, device=dev, dtype=i32)
        t_t = t if t is not None else torch.zeros((B, nest), device=dev, dtype=f64)
        c_t = torch.zeros((B, nest * idim), device=dev, dtype=f64)
        fp_t = torch.zeros(B, device=dev, dtype=f64)
        ier_t = torch.zeros(B, device=dev, dtype=i32)
        stats_t = torch.zeros((B, 2), device=dev, dtype=i32) if want_stats else None
        self.use_torch_stream()
        rc = lib().fp_parcur_batch_dev_c(
            self._h, B, int(iopt), int(ipar), idim, m, _ptr(u), 1, u.stride(0), _ptr(x), 1, x.stride(0),
            x.stride(1), _ptr(w), 1, 0 if w is None else w.stride(0), _ptr(ub), _ptr(ue), 1, int(k),
            _ptr(s_t), s_stride, nest, _ptr(n_t), _ptr(t_t), _ptr(c_t), _ptr(fp_t), _ptr(ier_t),
            _ptr(fpint), _ptr(nrdata), _ptr(stats_t))
        check(rc, "fp_parcur_batch_dev_c")
        res = dict(n=n_t, t=t_t, c=c_t, fp=fp_t, ier=ier_t, u=u, _keep=(s_t,))
        if want_stats:
            res["stats"] = stats_t
        return res

    def curev_batch(self, idim, t, n, c, k, u, out=None, mode=0):
        """curev for many curves: t [B][ldt], c [B][ldc], u [B][m] -> (x [B][m][idim], ier [B]);
        out = (x, ier) reuses the result tensors of an earlier call.  mode 0: exact (reference arithmetic,
        bit-identical), 1: fast (local polynomials + Horner, tolerance 1e-12 * max(1, max|c|))."""
        _chk_f64(t, "t"); _chk_f64(c, "c"); _chk_f64(u, "u")
        B, m = u.shape
        if out is not None:
            x, ier = out
        else:
            x = torch.zeros((B, m, idim), device=u.device, dtype=torch.float64)
            ier = torch.zeros(B, device=u.device, dtype=torch.int32)
        self.use_torch_stream()
        check(lib().fpb_engine_set_curev_mode(self._h, int(mode)), "set_curev_mode")
        rc = lib().fp_curev_batch_dev_c(self._h, B, int(idim), _ptr(t), _ptr(n), _ptr(c), t.stride(0),
                                        c.stride(0), int(k), _ptr(u.contiguous()), _ptr(x), m, _ptr(ier))
        check(rc, "fp_curev_batch_dev_c")
        return x, ier

    def percur_batch(self, x, y, w=None, k=3, s=0.0, nest=None, iopt=0, n=None, t=None, fpint=None,
                     nrdata=None):
        """Batched percur on device tensors, point-major: x, y, w are [m][B].  Returns dict(n, t, c,
        fp, ier) (t, c: [B][nest]); evaluate with splev_batch."""
        _chk_f64(x, "x"); _chk_f64(y, "y"); _chk_f64(w, "w")
        m, B = y.shape
        dev = y.device
        f64, i32 = torch.float64, torch.int32
        nest = int(nest if nest is not None else m + 2 * k)
        if torch.is_tensor(s):
            s_t, s_stride = s.to(device=dev, dtype=f64).contiguous(), 1
        else:
            s_t, s_stride = torch.full((1,), float(s), device=dev, dtype=f64), 0
        n_t = n if n is not None else torch.zeros(B, device=dev, dtype=i32)
        t_t = t if t is not None else torch.zeros((B, nest), device=dev, dtype=f64)
        c_t = torch.zeros((B, nest), device=dev, dtype=f64)
        fp_t = torch.zeros(B, device=dev, dtype=f64)
        ier_t = torch.zeros(B, device=dev, dtype=i32)
        self.use_torch_stream()
        rc = lib().fp_percur_batch_dev_c(
            self._h, B, int(iopt), m, _ptr(x), 1, x.stride(0), _ptr(y), 1, y.stride(0), _ptr(w), 1,
            0 if w is None else w.stride(0), int(k), _ptr(s_t), s_stride, nest, _ptr(n_t), _ptr(t_t),
            _ptr(c_t), _ptr(fp_t), _ptr(ier_t), _ptr(fpint), _ptr(nrdata))
        check(rc, "fp_percur_batch_dev_c")
        return dict(n=n_t, t=t_t, c=c_t, fp=fp_t, ier=ier_t, _keep=(s_t,))

    def curfit_shared(self, x, y, t, k=3, w=None, xb=None, xe=None, layout="point_major",
                      out=None, mode=0):
        """Shared-abscissa least squares (iopt=-1): x [m], t [n] (interior knots in place),
        y [m][B] (point_major) or [B][m].  Returns dict(c [B][n-k-1], fp [B], ier [1], t)."""
        _chk_f64(y, "y"); _chk_f64(x, "x"); _chk_f64(w, "w"); _chk_f64(t, "t")
        if layout == "point_major":
            m, B = y.shape
            sb, si = 1, y.stride(0)
        else:
            B, m = y.shape
            sb, si = y.stride(0), 1
        n = t.numel()
        nk1 = n - k - 1
        dev = y.device
        if xb is None:
            xb, xe = float(x[0]), float(x[-1])
        out = out or {}
        c_t = out.get("c")
        if c_t is None:
            c_t = torch.zeros((B, max(nk1, 1)), device=dev, dtype=torch.float64)
        fp_t = out.get("fp")
        if fp_t is None:
            fp_t = torch.zeros(B, device=dev, dtype=torch.float64)
        ier_t = out.get("ier")
        if ier_t is None:
            ier_t = torch.zeros(1, device=dev, dtype=torch.int32)
        self.use_torch_stream()
        check(lib().fpb_engine_set_shared_mode(self._h, int(mode)), "set_shared_mode")
        rc = lib().fp_curfit_shared_dev_c(self._h, B, m, _ptr(x), _ptr(w), _ptr(y), sb, si,
                                          float(xb), float(xe), int(k), n, _ptr(t), _ptr(c_t),
                                          c_t.stride(0), _ptr(fp_t), _ptr(ier_t))
        check(rc, "fp_curfit_shared_dev_c")
        return dict(c=c_t, fp=fp_t, ier=ier_t, t=t)

    # ----------------------------------------------------------------- eval
    def splev_batch(self, t, n, c, k, x, e=0, mode=0, shared=False, want_idx=False, out=None):
        """t, c: [nspl][ldt]; n: [nspl] int32; x: [nspl][m].  shared=True: one spline (row 0)
        for all rows of x.  Returns dict(y, ier[, lidx])."""
        _chk_f64(t, "t"); _chk_f64(c, "c"); _chk_f64(x, "x")
        nrows, m = x.shape
        ldt = t.stride(0) if t.dim() == 2 else t.numel()
        dev = x.device
        out = out or {}
        y = out.get("y")
        if y is None:
            y = torch.zeros_like(x)
        ier = out.get("ier")
        if ier is None:
            ier = torch.zeros(nrows, device=dev, dtype=torch.int32)
        lidx = torch.zeros(x.shape, device=dev, dtype=torch.int32) if want_idx else None
        self.use_torch_stream()
        rc = lib().fp_splev_batch_dev_c(self._h, nrows, _ptr(t), _ptr(n), _ptr(c), int(ldt), int(k),
                                        _ptr(x), _ptr(y), m, int(e), int(mode),
                                        1 if shared else 0, _ptr(ier), _ptr(lidx))
        check(rc, "fp_splev_batch_dev_c")
        res = dict(y=y, ier=ier)
        if want_idx:
            res["lidx"] = lidx
        return res

    def transpose(self, src):
        """[rows][cols] -> [cols][rows] with the engine's tile-transpose kernel."""
        _chk_f64(src, "src")
        rows, cols = src.shape
        dst = torch.empty((cols, rows), device=src.device, dtype=torch.float64)
        self.use_torch_stream()
        check(lib().fpb_transpose_dev(self._h, _ptr(src), src.stride(0), _ptr(dst), dst.stride(0),
                                      rows, cols), "fpb_transpose_dev")
        return dst
