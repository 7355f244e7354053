"""Flat entry points with the reference's argument conventions (fitpack_core curfit/splev:
iopt, k, s, nest, n, t, c, fp, wrk, iwrk, ier) on host numpy arrays.  Each call goes through
the C-ABI (fp_curfit_c / fp_splev_c / fp_*_batch_c) and therefore through the CUDA kernels;
without a GPU the calls report an error -- there is no CPU path here.
"""
import ctypes as C

import numpy as np

from ._lib import FitpackB200Error, check, lib

OUTSIDE_EXTRAPOLATE, OUTSIDE_ZERO, OUTSIDE_NOT_ALLOWED, OUTSIDE_NEAREST_BND = 0, 1, 2, 3
IOPT_NEW_LEASTSQUARES, IOPT_NEW_SMOOTHING, IOPT_OLD_FIT = -1, 0, 1
FPB_ERR_CUDA = -100


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def FITPACK_SUCCESS(ier):
    """reference core:393-396"""
    return bool(lib().FITPACK_SUCCESS_c(int(ier)))


def FITPACK_MESSAGE(ier):
    """reference core:315-339 via fitpack_message_c"""
    buf = C.create_string_buffer(128)
    lib().fitpack_message_c(int(ier), buf)
    return buf.value.decode()


def lwrk_min(m, k, nest):
    return m * (k + 1) + nest * (7 + 3 * k)


def curfit(iopt, x, y, w, xb, xe, k, s, nest, n=0, t=None, c=None, fp=0.0, wrk=None, iwrk=None,
           lwrk=None):
    """curfit (reference core:1240-1299) through fp_curfit_c.

    t, c, wrk, iwrk are updated IN PLACE when given (pass them back for iopt=1);
    returns dict(n, t, c, fp, ier, wrk, iwrk).
    """
    x = _f64(x)
    y = _f64(y)
    m = x.size
    w = np.ones(m) if w is None else _f64(w)
    if wrk is None:
        wrk = np.zeros(max(lwrk_min(m, max(k, 0), max(nest, 0)) if lwrk is None else lwrk, 1))
    if lwrk is None:
        lwrk = wrk.size
    t = np.zeros(max(nest, 1)) if t is None else t
    c = np.zeros(max(nest, 1)) if c is None else c
    iwrk = np.zeros(max(nest, 1), dtype=np.int32) if iwrk is None else iwrk
    for name, arr, dt in (("t", t, np.float64), ("c", c, np.float64), ("wrk", wrk, np.float64),
                          ("iwrk", iwrk, np.int32)):
        if arr.dtype != dt or not arr.flags.c_contiguous:
            raise TypeError("%s must be a contiguous %s array" % (name, dt.__name__))
    n_ = C.c_int32(int(n))
    fp_ = C.c_double(float(fp))
    ier_ = C.c_int32(0)
    lib().fp_curfit_c(int(iopt), m, _p(x), _p(y), _p(w), float(xb), float(xe), int(k), float(s),
                      int(nest), C.addressof(n_), _p(t), _p(c), C.addressof(fp_), _p(wrk),
                      int(lwrk), _p(iwrk), C.addressof(ier_))
    if ier_.value <= FPB_ERR_CUDA:
        raise FitpackB200Error("fp_curfit_c: %s" % lib().fpb_last_error().decode())
    return dict(n=n_.value, t=t, c=c, fp=fp_.value, ier=ier_.value, wrk=wrk, iwrk=iwrk)


def splev(t, n, c, k, x, e=OUTSIDE_EXTRAPOLATE, y=None):
    """splev (reference core:17826-17896) through fp_splev_c; returns (y, ier)."""
    t = _f64(t)
    c = _f64(c)
    x = _f64(x)
    m = x.size
    if y is None:
        y = np.zeros(max(m, 1))
    ier_ = C.c_int32(0)
    lib().fp_splev_c(_p(t), int(n), _p(c), int(k), _p(x), _p(y), m, int(e), C.addressof(ier_))
    if ier_.value <= FPB_ERR_CUDA:
        raise FitpackB200Error("fp_splev_c: %s" % lib().fpb_last_error().decode())
    return y[:m], ier_.value


def curfit_batch(iopt, x, y, w, k, s, nest, xb=None, xe=None, shared_x=False, n=None, t=None,
                 ngpus=1):
    """Many independent curfit problems, host arrays, curve-major (fp_curfit_batch_c).

    y: [B][m]; x, w: [B][m] or [m] when shared_x.  Returns dict(n, t, c, fp, ier).
    """
    y = _f64(y)
    B, m = y.shape
    x = _f64(x)
    w = None if w is None else _f64(w)
    n_arr = np.zeros(B, dtype=np.int32) if n is None else np.ascontiguousarray(n, dtype=np.int32)
    t_arr = np.zeros((B, nest)) if t is None else _f64(t)
    c_arr = np.zeros((B, nest))
    fp = np.zeros(B)
    ier = np.zeros(B, dtype=np.int32)
    xb_ = None if xb is None else C.c_double(float(xb))
    xe_ = None if xe is None else C.c_double(float(xe))
    rc = lib().fp_curfit_batch_c(B, int(iopt), m, _p(x), _p(y), _p(w), 1 if shared_x else 0,
                                 None if xb_ is None else C.addressof(xb_),
                                 None if xe_ is None else C.addressof(xe_), int(k), float(s),
                                 int(nest), _p(n_arr), _p(t_arr), _p(c_arr), _p(fp), _p(ier),
                                 int(ngpus))
    check(rc, "fp_curfit_batch_c")
    return dict(n=n_arr, t=t_arr, c=c_arr, fp=fp, ier=ier)


def splev_batch(t, n, c, k, x, e=OUTSIDE_EXTRAPOLATE, mode=0, ngpus=1):
    """Many splines, m points each, host arrays (fp_splev_batch_c).  Returns (y, ier)."""
    t = _f64(t)
    c = _f64(c)
    x = _f64(x)
    n = np.ascontiguousarray(n, dtype=np.int32)
    nspl, ldt = t.shape
    m = x.shape[1]
    y = np.zeros_like(x)
    ier = np.zeros(nspl, dtype=np.int32)
    rc = lib().fp_splev_batch_c(nspl, _p(t), _p(n), _p(c), ldt, int(k), _p(x), _p(y), m, int(e),
                                int(mode), _p(ier), int(ngpus))
    check(rc, "fp_splev_batch_c")
    return y, ier


# ------------------------------------------------------------------ gridded surfaces (dims = 2)
def regrid_lwrk(mx, my, kx, ky, nxest, nyest):
    """minimum (lwrk, kwrk) of regrid at dims=2 (reference core:16789-16797)"""
    maxm, maxnest, maxk1 = max(mx, my), max(nxest, nyest), max(kx, ky) + 1
    maxk2 = maxk1 + 1
    nkx, nky = nxest - kx - 1, nyest - ky - 1
    mm = max(maxnest, maxm, my, nkx, nky)
    lwrk = 4 + maxnest * 2 + maxm * maxk1 * 2 + 2 * maxnest * maxk2 * 2 + mm + 2 * nkx * my
    kwrk = 3 + maxm * 2 + maxnest * 2
    return lwrk, kwrk


def regrid(iopt, x, y, z, xb, xe, yb, ye, kx, ky, s, nxest, nyest, nx=0, ny=0, tx=None, ty=None,
           c=None, wrk=None, iwrk=None, engine=None):
    """regrid (reference core:16732-16839, dims=2) through fp_regrid_c, or fp_regrid_dev_c when
    `z` is a CUDA torch tensor (then `engine` is a fitpack_b200.batch.Engine).
    tx, ty, c, wrk, iwrk are updated IN PLACE when given (pass them back for iopt=1).
    Returns dict(nx, tx, ny, ty, c, fp, ier, wrk, iwrk)."""
    x = _f64(x)
    y = _f64(y)
    mx, my = x.size, y.size
    lw, kw = regrid_lwrk(mx, my, max(kx, 0), max(ky, 0), max(nxest, 0), max(nyest, 0))
    tx = np.zeros(max(nxest, 1)) if tx is None else tx
    ty = np.zeros(max(nyest, 1)) if ty is None else ty
    c = np.zeros(max((nxest - kx - 1) * (nyest - ky - 1), 1)) if c is None else c
    wrk = np.zeros(max(lw, 8)) if wrk is None else wrk
    iwrk = np.zeros(max(kw, 8), dtype=np.int32) if iwrk is None else iwrk
    nx_, ny_ = C.c_int32(int(nx)), C.c_int32(int(ny))
    fp_, ier_ = C.c_double(0.0), C.c_int32(0)
    tail = (float(xb), float(xe), float(yb), float(ye), int(kx), int(ky), float(s), int(nxest),
            int(nyest), C.addressof(nx_), _p(tx), C.addressof(ny_), _p(ty), _p(c), C.addressof(fp_),
            _p(wrk), int(wrk.size), _p(iwrk), int(iwrk.size), C.addressof(ier_))
    if hasattr(z, "data_ptr"):  # torch tensor in HBM
        if engine is None:
            raise ValueError("regrid with a device tensor needs engine=")
        assert z.is_cuda and z.is_contiguous() and z.numel() == mx * my
        rc = lib().fp_regrid_dev_c(engine.handle, int(iopt), mx, _p(x), my, _p(y),
                                   C.c_void_p(z.data_ptr()), *tail)
        check(rc, "fp_regrid_dev_c")
    else:
        z = _f64(z).ravel()
        assert z.size == mx * my
        lib().fp_regrid_c(int(iopt), mx, _p(x), my, _p(y), _p(z), *tail)
    if ier_.value <= FPB_ERR_CUDA:
        raise FitpackB200Error("fp_regrid_c: %s" % lib().fpb_last_error().decode())
    return dict(nx=nx_.value, tx=tx, ny=ny_.value, ty=ty, c=c, fp=fp_.value, ier=ier_.value, wrk=wrk,
                iwrk=iwrk)


def bispev(tx, nx, ty, ny, c, kx, ky, x, y, out=None, engine=None, mode=0):
    """bispev (reference core:424-458) through fp_bispev_c; `out`: optional CUDA torch tensor
    [mx, my] to receive the values in HBM (fp_bispev_dev_c; mode 0 exact / 1 fast = FMA sums).
    Returns (z, ier)."""
    tx, ty, c, x, y = _f64(tx), _f64(ty), _f64(c), _f64(x), _f64(y)
    mx, my = x.size, y.size
    if out is not None:
        ier_ = C.c_int32(0)
        check(lib().fpb_engine_set_bispev_mode(engine.handle, int(mode)), "set_bispev_mode")
        rc = lib().fp_bispev_dev_c(engine.handle, _p(tx), int(nx), _p(ty), int(ny), _p(c), int(kx),
                                   int(ky), _p(x), mx, _p(y), my, C.c_void_p(out.data_ptr()),
                                   C.addressof(ier_))
        check(rc, "fp_bispev_dev_c")
        return out, ier_.value
    z = np.zeros(max(mx * my, 1))
    lwrk = (kx + 1) * mx + (ky + 1) * my
    wrk = np.zeros(max(lwrk, 1))
    iwrk = np.zeros(max(mx + my, 1), dtype=np.int32)
    ier = lib().fp_bispev_c(_p(tx), int(nx), _p(ty), int(ny), _p(c), int(kx), int(ky), _p(x), mx,
                            _p(y), my, _p(z), _p(wrk), lwrk, _p(iwrk), mx + my)
    if ier <= FPB_ERR_CUDA:
        raise FitpackB200Error("fp_bispev_c: %s" % lib().fpb_last_error().decode())
    return z[:mx * my].reshape(mx, my), int(ier)


# ------------------------------------------------------------------ spline calculus
def splder(t, n, c, k, nu, x, e=OUTSIDE_EXTRAPOLATE):
    """splder (reference core:17699-17802) through fp_splder_c; returns (y, ier, wrk)."""
    t, c, x = _f64(t), _f64(c), _f64(x)
    m = x.size
    y = np.zeros(max(m, 1))
    wrk = np.zeros(max(int(n), 1))
    ier_ = C.c_int32(0)
    lib().fp_splder_c(_p(t), int(n), _p(c), int(k), int(nu), _p(x), _p(y), m, int(e), _p(wrk),
                      C.addressof(ier_))
    if ier_.value <= FPB_ERR_CUDA:
        raise FitpackB200Error("fp_splder_c: %s" % lib().fpb_last_error().decode())
    return y[:m], ier_.value, wrk


def spalde(t, n, c, k1, x):
    """spalde (reference core:16865-16895) through fp_spalde_c; returns (d[k1], ier)."""
    t, c = _f64(t), _f64(c)
    d = np.zeros(max(int(k1), 1))
    ier_ = C.c_int32(0)
    lib().fp_spalde_c(_p(t), int(n), _p(c), int(k1), float(x), _p(d), C.addressof(ier_))
    if ier_.value <= FPB_ERR_CUDA:
        raise FitpackB200Error("fp_spalde_c: %s" % lib().fpb_last_error().decode())
    return d, ier_.value


def splint(t, n, c, k, a, b):
    """splint (reference core:17919-17939) through fp_splint_c; returns (integral, wrk)."""
    t, c = _f64(t), _f64(c)
    wrk = np.zeros(max(int(n), 1))
    v = lib().fp_splint_c(_p(t), int(n), _p(c), int(k), float(a), float(b), _p(wrk))
    if v != v and lib().fpb_device_count() == 0:
        raise FitpackB200Error("fp_splint_c: %s" % lib().fpb_last_error().decode())
    return float(v), wrk


def cualde(idim, t, n, c, k1, u):
    """cualde (reference core:1062-1101) through fp_cualde_c; returns (d [k1][idim], ier)."""
    t, c = _f64(t), _f64(c)
    d = np.zeros(max(k1 * idim, 1))
    ier_ = C.c_int32(0)
    lib().fp_cualde_c(int(idim), _p(t), int(n), _p(c), c.size, int(k1), float(u), _p(d), d.size, C.addressof(ier_))
    if ier_.value <= FPB_ERR_CUDA:
        raise FitpackB200Error("fp_cualde_c: %s" % lib().fpb_last_error().decode())
    return d[:k1 * idim].reshape(k1, idim), ier_.value


def insert(iopt, t, n, c, k, x, nest):
    """insert (reference insert_inplace core:15138-15159) through fp_insert_c; returns (t, n, c, ier) -- copies of
    length nest with the knot x inserted (unchanged when ier = 10)."""
    tt, cc = np.zeros(nest), np.zeros(nest)
    tt[:min(nest, len(t))] = np.asarray(t, dtype=np.float64)[:nest]
    cc[:min(nest, len(c))] = np.asarray(c, dtype=np.float64)[:nest]
    n_, ier_ = C.c_int32(int(n)), C.c_int32(0)
    lib().fp_insert_c(int(iopt), _p(tt), C.addressof(n_), _p(cc), int(k), float(x), int(nest), C.addressof(ier_))
    if ier_.value <= FPB_ERR_CUDA:
        raise FitpackB200Error("fp_insert_c: %s" % lib().fpb_last_error().decode())
    return tt, n_.value, cc, ier_.value


def sproot(t, n, c, mest):
    """sproot (reference core:17962-18109) through fp_sproot_c; returns (zeros[:m], m, ier)."""
    t, c = _f64(t), _f64(c)
    z = np.zeros(max(int(mest), 1))
    m_, ier_ = C.c_int32(0), C.c_int32(0)
    lib().fp_sproot_c(_p(t), int(n), _p(c), _p(z), int(mest), C.addressof(m_), C.addressof(ier_))
    if ier_.value <= FPB_ERR_CUDA:
        raise FitpackB200Error("fp_sproot_c: %s" % lib().fpb_last_error().decode())
    return z[:min(m_.value, int(mest))].copy(), m_.value, ier_.value


def fourco(t, n, c, alfa):
    """fourco (reference core:1474-1520) through fp_fourco_c: the integrals of a cubic spline against sin(alfa x) and
    cos(alfa x) over its support; returns (ress, resc, ier, wrk1, wrk2)."""
    t, c, alfa = _f64(t), _f64(c), _f64(alfa)
    m = alfa.size
    ress, resc = np.zeros(max(m, 1)), np.zeros(max(m, 1))
    wrk1, wrk2 = np.zeros(max(int(n), 1)), np.zeros(max(int(n), 1))
    ier_ = C.c_int32(0)
    lib().fp_fourco_c(_p(t), int(n), _p(c), _p(alfa), m, _p(ress), _p(resc), _p(wrk1), _p(wrk2), C.addressof(ier_))
    if ier_.value <= FPB_ERR_CUDA:
        raise FitpackB200Error("fp_fourco_c: %s" % lib().fpb_last_error().decode())
    return ress[:m], resc[:m], ier_.value, wrk1, wrk2


# ------------------------------------------------------------------ parametric curves
def parcur_lwrk(m, k, nest, idim):
    return m * (k + 1) + nest * (6 + idim + 3 * k)


def parcur(iopt, ipar, idim, x, w, k, s, nest, u=None, ub=0.0, ue=1.0, n=0, t=None, wrk=None, iwrk=None,
           c=None):
    """parcur (reference core:15201-15285) through fp_parcur_c.  x is [m][idim] (memory order of
    Fortran's x(idim,m)).  Returns dict(n,t,c,fp,ier,u,ub,ue,wrk,iwrk); pass t/n/u/ub/ue/wrk/iwrk
    back for an iopt=1 continuation."""
    x = _f64(x)
    m = x.shape[0]
    w = np.ones(m) if w is None else _f64(w)
    u = np.zeros(m) if u is None else np.array(u, dtype=np.float64)
    idim_, k_, nest_ = max(idim, 1), max(k, 0), max(nest, 0)
    if wrk is None:
        wrk = np.zeros(max(parcur_lwrk(m, k_, nest_, idim_), 1))
    t = np.zeros(max(nest, 1)) if t is None else np.array(t, dtype=np.float64)
    c = np.zeros(max(nest_ * idim_, 1)) if c is None else np.array(c, dtype=np.float64)
    iwrk = np.zeros(max(nest, 1), dtype=np.int32) if iwrk is None else iwrk
    n_, fp_, ier_ = C.c_int32(int(n)), C.c_double(0.0), C.c_int32(0)
    ub_, ue_, s_ = C.c_double(ub), C.c_double(ue), C.c_double(s)
    lib().fp_parcur_c(int(iopt), int(ipar), int(idim), m, _p(u), x.size, _p(x), _p(w), C.addressof(ub_),
                      C.addressof(ue_), int(k), C.addressof(s_), int(nest), C.addressof(n_), _p(t),
                      c.size, _p(c), C.addressof(fp_), _p(wrk), wrk.size, _p(iwrk), C.addressof(ier_))
    if ier_.value <= FPB_ERR_CUDA:
        raise FitpackB200Error("fp_parcur_c: %s" % lib().fpb_last_error().decode())
    return dict(n=n_.value, t=t, c=c, fp=fp_.value, ier=ier_.value, u=u, ub=ub_.value, ue=ue_.value,
                wrk=wrk, iwrk=iwrk)


def curev(idim, t, n, c, k, u):
    """curev (reference core:1126-1182) through fp_curev_c; returns (x [m][idim], ier)."""
    t, c, u = _f64(t), _f64(c), _f64(u)
    m = u.size
    x = np.zeros(max(m * max(idim, 1), 1))
    ier_ = C.c_int32(0)
    lib().fp_curev_c(int(idim), _p(t), int(n), _p(c), c.size, int(k), _p(u), m, _p(x), m * idim,
                     C.addressof(ier_))
    if ier_.value <= FPB_ERR_CUDA:
        raise FitpackB200Error("fp_curev_c: %s" % lib().fpb_last_error().decode())
    return x[:m * idim].reshape(m, idim), ier_.value


def parcur_batch(iopt, ipar, idim, x, w, k, s, nest, u=None, ub=None, ue=None, n=None, t=None,
                 fpint=None, nrdata=None, ngpus=1, c=None):
    """Many independent parametric curves, host arrays (fp_parcur_batch_c).  x [B][m][idim]."""
    x = _f64(x)
    B, m = x.shape[0], x.shape[1]
    u = np.zeros((B, m)) if u is None else np.array(u, dtype=np.float64)
    ub = np.zeros(B) if ub is None else np.array(ub, dtype=np.float64)
    ue = np.ones(B) if ue is None else np.array(ue, dtype=np.float64)
    w = None if w is None else _f64(w)
    n_arr = np.zeros(B, dtype=np.int32) if n is None else np.array(n, dtype=np.int32)
    t_arr = np.zeros((B, nest)) if t is None else np.array(t, dtype=np.float64)
    c_arr = np.zeros((B, nest * idim)) if c is None else np.array(c, dtype=np.float64).reshape(B, nest * idim)
    fp = np.zeros(B)
    ier = np.zeros(B, dtype=np.int32)
    fpint = np.zeros((B, nest)) if fpint is None else fpint
    nrdata = np.zeros((B, nest), dtype=np.int32) if nrdata is None else nrdata
    rc = lib().fp_parcur_batch_c(B, int(iopt), int(ipar), int(idim), m, _p(u), _p(x), _p(w), _p(ub), _p(ue),
                                 int(k), float(s), int(nest), _p(n_arr), _p(t_arr), _p(c_arr), _p(fp),
                                 _p(ier), _p(fpint), _p(nrdata), int(ngpus))
    check(rc, "fp_parcur_batch_c")
    return dict(n=n_arr, t=t_arr, c=c_arr, fp=fp, ier=ier, u=u, ub=ub, ue=ue, fpint=fpint, nrdata=nrdata)


# ------------------------------------------------------------------ periodic curves
def percur_lwrk(m, k, nest):
    return m * (k + 1) + nest * (8 + 5 * k)


def percur(iopt, x, y, w, k, s, nest, n=0, t=None, wrk=None, iwrk=None, c=None, fp=0.0):
    """percur (reference core:15784-15863) through fp_percur_c; same conventions as curfit()."""
    x, y = _f64(x), _f64(y)
    m = x.size
    w = np.ones(m) if w is None else _f64(w)
    if wrk is None:
        wrk = np.zeros(max(percur_lwrk(m, max(k, 0), max(nest, 0)), 1))
    t = np.zeros(max(nest, 1)) if t is None else np.array(t, dtype=np.float64)
    c = np.zeros(max(nest, 1)) if c is None else np.array(c, dtype=np.float64)
    iwrk = np.zeros(max(nest, 1), dtype=np.int32) if iwrk is None else iwrk
    n_, fp_, ier_ = C.c_int32(int(n)), C.c_double(float(fp)), C.c_int32(0)
    lib().fp_percur_c(int(iopt), m, _p(x), _p(y), _p(w), int(k), float(s), int(nest), C.addressof(n_), _p(t),
                      _p(c), C.addressof(fp_), _p(wrk), wrk.size, _p(iwrk), C.addressof(ier_))
    if ier_.value <= FPB_ERR_CUDA:
        raise FitpackB200Error("fp_percur_c: %s" % lib().fpb_last_error().decode())
    return dict(n=n_.value, t=t, c=c, fp=fp_.value, ier=ier_.value, wrk=wrk, iwrk=iwrk)


def percur_batch(iopt, x, y, w, k, s, nest, n=None, t=None, fpint=None, nrdata=None, ngpus=1):
    """Many independent periodic fits, host arrays curve-major (fp_percur_batch_c)."""
    x, y = _f64(x), _f64(y)
    B, m = y.shape
    w = None if w is None else _f64(w)
    n_arr = np.zeros(B, dtype=np.int32) if n is None else np.array(n, dtype=np.int32)
    t_arr = np.zeros((B, nest)) if t is None else np.array(t, dtype=np.float64)
    c_arr = np.zeros((B, nest))
    fp = np.zeros(B)
    ier = np.zeros(B, dtype=np.int32)
    rc = lib().fp_percur_batch_c(B, int(iopt), m, _p(x), _p(y), _p(w), int(k), float(s), int(nest), _p(n_arr),
                                 _p(t_arr), _p(c_arr), _p(fp), _p(ier), _p(fpint), _p(nrdata), int(ngpus))
    check(rc, "fp_percur_batch_c")
    return dict(n=n_arr, t=t_arr, c=c_arr, fp=fp, ier=ier)
