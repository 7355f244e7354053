"""ctypes binding of the CPU oracle (oracle/libfitpack_oracle.so).

Test infrastructure only: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(fitpack_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_LIB_PATH = os.path.join(ORACLE_DIR, "libfitpack_oracle.so")
# bench.py's reference arm times the -O3 build of the same source (FITPACK_ORACLE_LIB=<path>)
_ALT = os.environ.get("FITPACK_ORACLE_LIB")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class OrcStats(C.Structure):
    _fields_ = [("lsq_passes", C.c_int32), ("p_iters", C.c_int32), ("rotations", C.c_int64)]


STATS_DTYPE = np.dtype([("lsq_passes", np.int32), ("p_iters", np.int32), ("rotations", np.int64)])


def build_oracle(force=False):
    src = os.path.join(ORACLE_DIR, "fitpack_oracle.c")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", ORACLE_DIR, "-B", "libfitpack_oracle.so"],
                          stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build_oracle()
        _lib = C.CDLL(_ALT if _ALT and os.path.exists(_ALT) else _LIB_PATH)
        _lib.orc_curfit_stats.restype = None
        _lib.orc_splev_idx.restype = None
        _lib.orc_fpchec.restype = C.c_int32
        _lib.orc_equal.restype = C.c_int32
        _lib.orc_equal.argtypes = [C.c_double, C.c_double]
        _lib.orc_knot_interval.restype = C.c_int32
        _lib.orc_message.restype = C.c_char_p
        _lib.orc_max_threads.restype = C.c_int32
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def lwrk_min(m, k, nest):
    return m * (k + 1) + nest * (7 + 3 * k)


def curfit(iopt, x, y, w, xb, xe, k, s, nest, n=0, t=None, wrk=None, iwrk=None, lwrk=None,
           c=None, fp=0.0):
    """orc_curfit with the reference's argument list (core:1240).

    Returns dict(n,t,c,fp,ier,wrk,iwrk,stats); t/c/wrk/iwrk are the (updated)
    arrays so that an iopt=1 continuation can pass them back in.
    """
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    m = x.size
    w = np.ones(m) if w is None else np.ascontiguousarray(w, dtype=np.float64)
    if lwrk is None:
        lwrk = max(lwrk_min(m, k, nest), 1) if wrk is None else wrk.size
    t = np.zeros(max(nest, 1)) if t is None else np.ascontiguousarray(t, dtype=np.float64)
    c = np.zeros(max(nest, 1)) if c is None else np.ascontiguousarray(c, dtype=np.float64)
    wrk = np.zeros(max(lwrk, 1)) if wrk is None else wrk
    iwrk = np.zeros(max(nest, 1), dtype=np.int32) if iwrk is None else iwrk
    n_ = C.c_int32(int(n))
    fp_ = C.c_double(float(fp))
    ier_ = C.c_int32(0)
    st = OrcStats()
    lib().orc_curfit_stats(C.c_int32(iopt), C.c_int32(m), _d(x), _d(y), _d(w), C.c_double(xb),
                           C.c_double(xe), C.c_int32(k), C.c_double(s), C.c_int32(nest),
                           C.byref(n_), _d(t), _d(c), C.byref(fp_), _d(wrk), C.c_int32(lwrk),
                           _i(iwrk), C.byref(ier_), C.byref(st))
    return dict(n=n_.value, t=t, c=c, fp=fp_.value, ier=ier_.value, wrk=wrk, iwrk=iwrk,
                stats=(st.lsq_passes, st.p_iters, st.rotations))


def splev(t, n, c, k, x, e=0, want_idx=False):
    t = np.ascontiguousarray(t, dtype=np.float64)
    c = np.ascontiguousarray(c, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    m = x.size
    y = np.zeros(max(m, 1))
    idx = np.zeros(max(m, 1), dtype=np.int32)
    ier = C.c_int32(0)
    lib().orc_splev_idx(_d(t), C.c_int32(n), _d(c), C.c_int32(k), _d(x), _d(y), _i(idx),
                        C.c_int32(m), C.c_int32(e), C.byref(ier))
    if want_idx:
        return y[:m], ier.value, idx[:m]
    return y[:m], ier.value


def fpchec(x, t, n, k):
    x = np.ascontiguousarray(x, dtype=np.float64)
    t = np.ascontiguousarray(t, dtype=np.float64)
    return lib().orc_fpchec(_d(x), C.c_int32(x.size), _d(t), C.c_int32(n), C.c_int32(k))


def fpbspl(t, n, k, x, l):
    t = np.ascontiguousarray(t, dtype=np.float64)
    h = np.zeros(20)
    lib().orc_fpbspl(_d(t), C.c_int32(n), C.c_int32(k), C.c_double(x), C.c_int32(l), _d(h))
    return h[:k + 1]


def knot_interval(t, x, l_start, l_max):
    t = np.ascontiguousarray(t, dtype=np.float64)
    return lib().orc_knot_interval(_d(t), C.c_double(x), C.c_int32(l_start), C.c_int32(l_max))


def equal(a, b):
    return bool(lib().orc_equal(a, b))


def argsort(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    il = np.zeros(max(x.size, 1), dtype=np.int32)
    lib().orc_argsort(_d(x), C.c_int32(x.size), _i(il))
    return il[:x.size]


def message(ier):
    return lib().orc_message(C.c_int32(ier)).decode()


def max_threads():
    return lib().orc_max_threads()


def curfit_batch(iopt, x, y, w, k, s, nest, xb=None, xe=None, shared_x=False, n=None, t=None,
                 nthreads=0, want_stats=False):
    """OpenMP-over-curves driver.  y is [B][m]; x,w are [B][m] or [m] if shared_x."""
    y = np.ascontiguousarray(y, dtype=np.float64)
    B, m = y.shape
    x = np.ascontiguousarray(x, dtype=np.float64)
    wp = None
    if w is not None:
        w = np.ascontiguousarray(w, dtype=np.float64)
        wp = _d(w)
    n_arr = np.zeros(B, dtype=np.int32) if n is None else np.ascontiguousarray(n, dtype=np.int32)
    t_arr = np.zeros((B, nest)) if t is None else np.ascontiguousarray(t, dtype=np.float64)
    c_arr = np.zeros((B, nest))
    fp = np.zeros(B)
    ier = np.zeros(B, dtype=np.int32)
    stats = np.zeros(B, dtype=STATS_DTYPE) if want_stats else None
    use_range = xb is None
    lib().orc_curfit_batch(C.c_int32(B), C.c_int32(iopt), C.c_int32(m), _d(x), _d(y), wp,
                           C.c_int32(1 if shared_x else 0),
                           C.c_double(0.0 if xb is None else xb),
                           C.c_double(0.0 if xe is None else xe), C.c_int32(1 if use_range else 0),
                           C.c_int32(k), C.c_double(s), C.c_int32(nest), _i(n_arr), _d(t_arr),
                           _d(c_arr), _d(fp), _i(ier),
                           stats.ctypes.data_as(C.c_void_p) if want_stats else None,
                           C.c_int32(nthreads))
    out = dict(n=n_arr, t=t_arr, c=c_arr, fp=fp, ier=ier)
    if want_stats:
        out["stats"] = stats
    return out


def splev_batch(t, n, c, k, x, e=0, nthreads=0):
    t = np.ascontiguousarray(t, dtype=np.float64)
    c = np.ascontiguousarray(c, dtype=np.float64)
    n = np.ascontiguousarray(n, dtype=np.int32)
    x = np.ascontiguousarray(x, dtype=np.float64)
    nspl, ldt = t.shape
    m = x.shape[1]
    y = np.zeros_like(x)
    ier = np.zeros(nspl, dtype=np.int32)
    lib().orc_splev_batch(C.c_int32(nspl), _d(t), _i(n), _d(c), C.c_int32(ldt), C.c_int32(k),
                          _d(x), _d(y), C.c_int32(m), C.c_int32(e), _i(ier), C.c_int32(nthreads))
    return y, ier


# ------------------------------------------------------------------ gridded surfaces (dims = 2)
def regrid_lwrk(mx, my, kx, ky, nxest, nyest):
    """minimum lwrk, kwrk of regrid at dims=2 (core:16789-16797)"""
    maxm, maxnest, maxk1 = max(mx, my), max(nxest, nyest), max(kx, ky) + 1
    maxk2 = maxk1 + 1
    nkx, nky = nxest - kx - 1, nyest - ky - 1
    mm = max(maxnest, maxm, my, nkx, nky)
    mynx = 2 * nkx * my
    lwrk = 4 + maxnest * 2 + maxm * maxk1 * 2 + 2 * maxnest * maxk2 * 2 + mm + mynx
    kwrk = 3 + maxm * 2 + maxnest * 2
    return lwrk, kwrk


def regrid(iopt, x, y, z, xb, xe, yb, ye, kx, ky, s, nxest, nyest, nx=0, ny=0, tx=None, ty=None,
           c=None, wrk=None, iwrk=None, lwrk=None, kwrk=None):
    """orc_regrid with the argument list of fp_regrid_c.  Returns dict(nx,tx,ny,ty,c,fp,ier,wrk,iwrk)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    z = np.ascontiguousarray(z, dtype=np.float64).ravel()
    mx, my = x.size, y.size
    lw, kw = regrid_lwrk(mx, my, kx, ky, nxest, nyest)
    lwrk = max(lw, 8) if lwrk is None else lwrk
    kwrk = max(kw, 8) if kwrk is None else kwrk
    tx = np.zeros(max(nxest, 1)) if tx is None else np.ascontiguousarray(tx, dtype=np.float64).copy()
    ty = np.zeros(max(nyest, 1)) if ty is None else np.ascontiguousarray(ty, dtype=np.float64).copy()
    ncmax = max((nxest - kx - 1) * (nyest - ky - 1), 1)
    c = np.zeros(ncmax) if c is None else c
    wrk = np.zeros(max(lwrk, 8)) if wrk is None else wrk
    iwrk = np.zeros(max(kwrk, 8), dtype=np.int32) if iwrk is None else iwrk
    nx_, ny_ = C.c_int32(int(nx)), C.c_int32(int(ny))
    fp_, ier_ = C.c_double(0.0), C.c_int32(0)
    L = lib()
    L.orc_regrid.restype = None
    L.orc_regrid(C.c_int32(iopt), C.c_int32(mx), _d(x), C.c_int32(my), _d(y), _d(z),
                 C.c_double(xb), C.c_double(xe), C.c_double(yb), C.c_double(ye), C.c_int32(kx),
                 C.c_int32(ky), C.c_double(s), C.c_int32(nxest), C.c_int32(nyest), C.byref(nx_),
                 _d(tx), C.byref(ny_), _d(ty), _d(c), C.byref(fp_), _d(wrk), C.c_int32(lwrk),
                 _i(iwrk), C.c_int32(kwrk), C.byref(ier_))
    return dict(nx=nx_.value, tx=tx, ny=ny_.value, ty=ty, c=c, fp=fp_.value, ier=ier_.value,
                wrk=wrk, iwrk=iwrk)


def bispev(tx, nx, ty, ny, c, kx, ky, x, y):
    tx = np.ascontiguousarray(tx, dtype=np.float64)
    ty = np.ascontiguousarray(ty, dtype=np.float64)
    c = np.ascontiguousarray(c, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    mx, my = x.size, y.size
    z = np.zeros(max(mx * my, 1))
    lwrk = (kx + 1) * mx + (ky + 1) * my
    wrk = np.zeros(max(lwrk, 1))
    iwrk = np.zeros(max(mx + my, 1), dtype=np.int32)
    L = lib()
    L.orc_bispev.restype = C.c_int32
    ier = L.orc_bispev(_d(tx), C.c_int32(nx), _d(ty), C.c_int32(ny), _d(c), C.c_int32(kx),
                       C.c_int32(ky), _d(x), C.c_int32(mx), _d(y), C.c_int32(my), _d(z), _d(wrk),
                       C.c_int32(lwrk), _i(iwrk), C.c_int32(mx + my))
    return z[:mx * my].reshape(mx, my), int(ier)


def set_netlib_compat(on):
    lib().orc_set_netlib_compat(C.c_int32(1 if on else 0))


# ------------------------------------------------------------------ spline calculus
def splder(t, n, c, k, nu, x, e=0):
    t, c, x = (np.ascontiguousarray(a, dtype=np.float64) for a in (t, c, x))
    m = x.size
    y = np.zeros(max(m, 1))
    wrk = np.zeros(max(n, 1))
    ier = C.c_int32(0)
    L = lib()
    L.orc_splder.restype = None
    L.orc_splder(_d(t), C.c_int32(n), _d(c), C.c_int32(k), C.c_int32(nu), _d(x), _d(y), C.c_int32(m),
                 C.c_int32(e), _d(wrk), C.byref(ier))
    return y[:m], ier.value, wrk


def spalde(t, n, c, k1, x):
    t, c = (np.ascontiguousarray(a, dtype=np.float64) for a in (t, c))
    d = np.zeros(k1)
    ier = C.c_int32(0)
    L = lib()
    L.orc_spalde.restype = None
    L.orc_spalde(_d(t), C.c_int32(n), _d(c), C.c_int32(k1), C.c_double(x), _d(d), C.byref(ier))
    return d, ier.value


def splint(t, n, c, k, a, b):
    t, c = (np.ascontiguousarray(v, dtype=np.float64) for v in (t, c))
    wrk = np.zeros(max(n, 1))
    L = lib()
    L.orc_splint.restype = C.c_double
    v = L.orc_splint(_d(t), C.c_int32(n), _d(c), C.c_int32(k), C.c_double(a), C.c_double(b), _d(wrk))
    return float(v), wrk


def cualde(idim, t, n, c, k1, u):
    """orc_cualde (reference core:1062-1101); returns (d [k1][idim], ier)"""
    t, c = (np.ascontiguousarray(v, dtype=np.float64) for v in (t, c))
    d = np.zeros(k1 * idim)
    ier = C.c_int32(0)
    L = lib()
    L.orc_cualde.restype = None
    L.orc_cualde(C.c_int32(idim), _d(t), C.c_int32(n), _d(c), C.c_int32(c.size), C.c_int32(k1), C.c_double(u), _d(d),
                 C.c_int32(d.size), C.byref(ier))
    return d.reshape(k1, idim), ier.value


def insert(iopt, t, n, c, k, x, nest):
    """orc_insert (reference insert_inplace core:15138-15159); returns (t, n, c, ier) -- copies, length nest"""
    tt, cc = np.zeros(nest), np.zeros(nest)
    tt[:min(nest, len(t))] = np.asarray(t, dtype=np.float64)[:nest]
    cc[:min(nest, len(c))] = np.asarray(c, dtype=np.float64)[:nest]
    n_, ier = C.c_int32(int(n)), C.c_int32(0)
    L = lib()
    L.orc_insert.restype = None
    L.orc_insert(C.c_int32(iopt), _d(tt), C.byref(n_), _d(cc), C.c_int32(k), C.c_double(x), C.c_int32(nest), C.byref(ier))
    return tt, n_.value, cc, ier.value


def sproot(t, n, c, mest):
    """orc_sproot (reference core:17962-18109); returns (zeros[:m], m, ier)"""
    t, c = (np.ascontiguousarray(v, dtype=np.float64) for v in (t, c))
    z = np.zeros(max(mest, 1))
    m, ier = C.c_int32(0), C.c_int32(0)
    L = lib()
    L.orc_sproot.restype = None
    L.orc_sproot(_d(t), C.c_int32(n), _d(c), _d(z), C.c_int32(mest), C.byref(m), C.byref(ier))
    return z[:m.value].copy(), m.value, ier.value


def fourco(t, n, c, alfa):
    """orc_fourco (reference core:1474-1520); returns (ress, resc, ier, wrk1, wrk2)"""
    t, c, alfa = (np.ascontiguousarray(v, dtype=np.float64) for v in (t, c, alfa))
    m = alfa.size
    ress, resc, w1, w2 = np.zeros(max(m, 1)), np.zeros(max(m, 1)), np.zeros(max(n, 1)), np.zeros(max(n, 1))
    ier = C.c_int32(0)
    L = lib()
    L.orc_fourco.restype = None
    L.orc_fourco(_d(t), C.c_int32(n), _d(c), _d(alfa), C.c_int32(m), _d(ress), _d(resc), _d(w1), _d(w2), C.byref(ier))
    return ress[:m], resc[:m], ier.value, w1, w2


# ------------------------------------------------------------------ parametric curves
def parcur_lwrk(m, k, nest, idim):
    return m * (k + 1) + nest * (6 + idim + 3 * k)


def parcur(iopt, ipar, idim, x, w, k, s, nest, u=None, ub=0.0, ue=1.0, n=0, t=None, wrk=None,
           iwrk=None, c=None):
    """orc_parcur with the reference's argument list (core:15201).  x is [m][idim] (the memory
    order of Fortran's x(idim,m)).  Returns dict(n,t,c,fp,ier,u,ub,ue,wrk,iwrk)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    m = x.shape[0]
    w = np.ones(m) if w is None else np.ascontiguousarray(w, dtype=np.float64)
    u = np.zeros(m) if u is None else np.array(u, dtype=np.float64)
    lwrk = parcur_lwrk(m, k, nest, idim) if wrk is None else wrk.size
    t = np.zeros(max(nest, 1)) if t is None else np.array(t, dtype=np.float64)
    c = np.zeros(max(nest * idim, 1)) if c is None else np.array(c, dtype=np.float64)
    wrk = np.zeros(max(lwrk, 1)) if wrk is None else wrk
    iwrk = np.zeros(max(nest, 1), dtype=np.int32) if iwrk is None else iwrk
    n_, fp_, ier_ = C.c_int32(int(n)), C.c_double(0.0), C.c_int32(0)
    ub_, ue_ = C.c_double(ub), C.c_double(ue)
    L = lib()
    L.orc_parcur.restype = None
    L.orc_parcur(C.c_int32(iopt), C.c_int32(ipar), C.c_int32(idim), C.c_int32(m), _d(u),
                 C.c_int32(m * idim), _d(x), _d(w), C.byref(ub_), C.byref(ue_), C.c_int32(k),
                 C.c_double(s), C.c_int32(nest), C.byref(n_), _d(t), C.c_int32(nest * idim), _d(c),
                 C.byref(fp_), _d(wrk), C.c_int32(lwrk), _i(iwrk), C.byref(ier_))
    return dict(n=n_.value, t=t, c=c, fp=fp_.value, ier=ier_.value, u=u, ub=ub_.value, ue=ue_.value,
                wrk=wrk, iwrk=iwrk)


def curev(idim, t, n, c, k, u):
    t, c, u = (np.ascontiguousarray(a, dtype=np.float64) for a in (t, c, u))
    m = u.size
    x = np.zeros(max(m * idim, 1))
    ier = C.c_int32(0)
    L = lib()
    L.orc_curev.restype = None
    L.orc_curev(C.c_int32(idim), _d(t), C.c_int32(n), _d(c), C.c_int32(c.size), C.c_int32(k), _d(u),
                C.c_int32(m), _d(x), C.c_int32(m * idim), C.byref(ier))
    return x[:m * idim].reshape(m, idim), ier.value


def parcur_batch(iopt, ipar, idim, x, w, k, s, nest, u=None, ub=None, ue=None, nthreads=0):
    """x [B][m][idim]; returns dict(n,t,c,fp,ier,u,ub,ue) with c [B][nest*idim]."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    B, m = x.shape[0], x.shape[1]
    u = np.zeros((B, m)) if u is None else np.array(u, dtype=np.float64)
    ub = np.zeros(B) if ub is None else np.array(ub, dtype=np.float64)
    ue = np.ones(B) if ue is None else np.array(ue, dtype=np.float64)
    n = np.zeros(B, np.int32)
    t = np.zeros((B, nest))
    c = np.zeros((B, nest * idim))
    fp = np.zeros(B)
    ier = np.zeros(B, np.int32)
    wp = None if w is None else _d(np.ascontiguousarray(w, dtype=np.float64))
    L = lib()
    L.orc_parcur_batch.restype = None
    L.orc_parcur_batch(C.c_int32(B), C.c_int32(iopt), C.c_int32(ipar), C.c_int32(idim), C.c_int32(m),
                       _d(u), _d(x), wp, _d(ub), _d(ue), C.c_int32(k), C.c_double(s), C.c_int32(nest),
                       _i(n), _d(t), _d(c), _d(fp), _i(ier), C.c_int32(nthreads))
    return dict(n=n, t=t, c=c, fp=fp, ier=ier, u=u, ub=ub, ue=ue)


# ------------------------------------------------------------------ periodic curves
def percur_lwrk(m, k, nest):
    return m * (k + 1) + nest * (8 + 5 * k)


def percur(iopt, x, y, w, k, s, nest, n=0, t=None, wrk=None, iwrk=None, c=None, fp=0.0):
    """orc_percur with the reference's argument list (core:15784)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    m = x.size
    w = np.ones(m) if w is None else np.ascontiguousarray(w, dtype=np.float64)
    lwrk = percur_lwrk(m, k, nest) if wrk is None else wrk.size
    t = np.zeros(max(nest, 1)) if t is None else np.array(t, dtype=np.float64)
    c = np.zeros(max(nest, 1)) if c is None else np.array(c, dtype=np.float64)
    wrk = np.zeros(max(lwrk, 1)) if wrk is None else wrk
    iwrk = np.zeros(max(nest, 1), dtype=np.int32) if iwrk is None else iwrk
    n_, fp_, ier_ = C.c_int32(int(n)), C.c_double(float(fp)), C.c_int32(0)
    L = lib()
    L.orc_percur.restype = None
    L.orc_percur(C.c_int32(iopt), C.c_int32(m), _d(x), _d(y), _d(w), C.c_int32(k), C.c_double(s),
                 C.c_int32(nest), C.byref(n_), _d(t), _d(c), C.byref(fp_), _d(wrk), C.c_int32(lwrk),
                 _i(iwrk), C.byref(ier_))
    return dict(n=n_.value, t=t, c=c, fp=fp_.value, ier=ier_.value, wrk=wrk, iwrk=iwrk)


def fpchep(x, t, n, k):
    x, t = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, t))
    L = lib()
    L.orc_fpchep.restype = C.c_int32
    return int(L.orc_fpchep(_d(x), C.c_int32(x.size), _d(t), C.c_int32(n), C.c_int32(k)))


def percur_batch(iopt, x, y, w, k, s, nest, nthreads=0):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    B, m = y.shape
    n = np.zeros(B, np.int32)
    t = np.zeros((B, nest))
    c = np.zeros((B, nest))
    fp = np.zeros(B)
    ier = np.zeros(B, np.int32)
    wp = None if w is None else _d(np.ascontiguousarray(w, dtype=np.float64))
    L = lib()
    L.orc_percur_batch.restype = None
    L.orc_percur_batch(C.c_int32(B), C.c_int32(iopt), C.c_int32(m), _d(x), _d(y), wp, C.c_int32(k),
                       C.c_double(s), C.c_int32(nest), _i(n), _d(t), _d(c), _d(fp), _i(ier),
                       C.c_int32(nthreads))
    return dict(n=n, t=t, c=c, fp=fp, ier=ier)


def curev_batch(idim, t, n, c, k, u, nthreads=0):
    """orc_curev_batch: t [B][ldt], c [B][ldc], u [B][m] -> (x [B][m][idim], ier [B])."""
    t, c, u = (np.ascontiguousarray(a, dtype=np.float64) for a in (t, c, u))
    n = np.ascontiguousarray(n, dtype=np.int32)
    B, m = u.shape
    x = np.zeros((B, m, idim))
    ier = np.zeros(B, np.int32)
    L = lib()
    L.orc_curev_batch.restype = None
    L.orc_curev_batch(C.c_int32(B), C.c_int32(idim), _d(t), _i(n), _d(c), C.c_int32(t.shape[1]),
                      C.c_int64(c.shape[1]), C.c_int32(k), _d(u), _d(x), C.c_int32(m), _i(ier),
                      C.c_int32(nthreads))
    return x, ier
