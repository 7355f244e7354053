"""FitpackCurve: Python mirror of the reference's `fitpack_curve` type (src/fitpack_curves.f90:
54-131) and of the C++ `fpCurve` wrapper (include/fpCurve.hpp:32-169), bound to the same
opaque-handle C-ABI (`fitpack_curve_c_*`, src/fitpack_curves_c.f90).  Same method names,
argument meaning and error convention (integer flag, <=0 is success).
"""
import ctypes as C

import numpy as np

from ._lib import lib


class _Handle(C.Structure):
    _fields_ = [("cptr", C.c_void_p)]


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _opt_d(v):
    return None if v is None else C.c_double(float(v))


def _opt_i(v):
    return None if v is None else C.c_int32(int(v))


def _addr(v):
    return None if v is None else C.addressof(v)


class FitpackCurve:
    def __init__(self, x=None, y=None, w=None, smoothing=None):
        self._h = _Handle(None)
        lib().fitpack_curve_c_allocate(C.addressof(self._h))
        self.ierr = 0
        if x is not None:
            self.ierr = self.new_fit(x, y, w, smoothing)

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def destroy(self):
        if self._h.cptr:
            lib().fitpack_curve_c_destroy(C.addressof(self._h))

    def copy(self):
        other = FitpackCurve()
        lib().fitpack_curve_c_copy(C.addressof(other._h), C.addressof(self._h))
        return other

    def new_points(self, x, y, w=None):
        x, y = _f64(x), _f64(y)
        w = None if w is None else _f64(w)
        lib().fitpack_curve_c_new_points(C.addressof(self._h), x.size, x.ctypes.data, y.ctypes.data,
                                         None if w is None else w.ctypes.data)

    def new_fit(self, x, y, w=None, smoothing=None, order=None):
        """ierr = curve%new_fit(x,y,w,smoothing,order)  (fitpack_curves.f90:168-179)"""
        if order is not None:
            self.new_points(x, y, w)
            return self.fit(smoothing, order)
        x, y = _f64(x), _f64(y)
        w = None if w is None else _f64(w)
        s = _opt_d(smoothing)
        return lib().fitpack_curve_c_new_fit(C.addressof(self._h), x.size, x.ctypes.data,
                                             y.ctypes.data, None if w is None else w.ctypes.data,
                                             _addr(s))

    def fit(self, smoothing=None, order=None):
        s, o = _opt_d(smoothing), _opt_i(order)
        return lib().fitpack_curve_c_fit(C.addressof(self._h), _addr(s), _addr(o))

    def interpolate(self, order=None):
        o = _opt_i(order)
        return lib().fitpack_curve_c_interpolating(C.addressof(self._h), _addr(o))

    def least_squares(self, smoothing=None, reset_knots=False):
        s = _opt_d(smoothing)
        r = C.c_bool(bool(reset_knots))
        return lib().fitpack_curve_c_least_squares(C.addressof(self._h), _addr(s), C.addressof(r))

    def eval(self, x):
        """y, ierr = curve%eval(x)  (scalar or array; fitpack_curves.f90:273-347)"""
        ier = C.c_int32(0)
        if np.isscalar(x):
            y = lib().fitpack_curve_c_eval_one(C.addressof(self._h), float(x), C.addressof(ier))
            return y, ier.value
        x = _f64(x)
        y = np.zeros(max(x.size, 1))
        lib().fitpack_curve_c_eval_many(C.addressof(self._h), x.size, x.ctypes.data, y.ctypes.data,
                                        C.addressof(ier))
        return y[:x.size], ier.value

    def dfdx(self, x, order=1):
        """ddx, ierr = curve%dfdx(x, order)  (scalar x; fitpack_curves.f90:606-618)"""
        ier = C.c_int32(0)
        v = lib().fitpack_curve_c_derivative(C.addressof(self._h), float(x), int(order), C.addressof(ier))
        return v, ier.value

    def dfdx_all(self, x):
        """ddx(1:order+1), ierr = curve%dfdx_all(x)  (fitpack_curves.f90:543-565)"""
        d = np.zeros(self.degree() + 1)
        ier = lib().fitpack_curve_c_all_derivatives(C.addressof(self._h), float(x), d.ctypes.data)
        return d, ier

    def integral(self, a, b):
        """curve%integral(from, to)  (splint)"""
        return lib().fitpack_curve_c_integral(C.addressof(self._h), float(a), float(b))

    def fourier(self, alpha):
        """curve%fourier_coefficients(alpha, A, B)  (fourco; cubic splines): returns (A, B, ierr)"""
        alpha = np.ascontiguousarray(alpha, dtype=np.float64)
        a, b = np.zeros(max(alpha.size, 1)), np.zeros(max(alpha.size, 1))
        ierr = C.c_int32(0)
        lib().fitpack_curve_c_fourier(C.addressof(self._h), alpha.size, alpha.ctypes.data_as(C.c_void_p),
                                      a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), C.addressof(ierr))
        return a[:alpha.size], b[:alpha.size], ierr.value

    # parallel communication (curve%comm_size / comm_pack / comm_expand, fitpack_curves.f90:821-895)
    def comm_size(self):
        return lib().fitpack_curve_c_comm_size(C.addressof(self._h))

    def comm_pack(self):
        buf = np.zeros(self.comm_size())
        lib().fitpack_curve_c_comm_pack(C.addressof(self._h), buf.ctypes.data)
        return buf

    def comm_expand(self, buffer):
        buffer = _f64(buffer)
        lib().fitpack_curve_c_comm_expand(C.addressof(self._h), buffer.ctypes.data)

    # accessors (public components of the Fortran type)
    def smoothing(self):
        return lib().fitpack_curve_c_smoothing(C.addressof(self._h))

    def mse(self):
        return lib().fitpack_curve_c_mse(C.addressof(self._h))

    def degree(self):
        return lib().fitpack_curve_c_degree(C.addressof(self._h))

    def set_bc(self, bc):
        lib().fitpack_curve_c_set_bc(C.addressof(self._h), int(bc))

    def get_bc(self):
        return lib().fitpack_curve_c_get_bc(C.addressof(self._h))

    @property
    def knots(self):
        return lib().fitpack_curve_c_knots(C.addressof(self._h))

    @property
    def iopt(self):
        return lib().fitpack_curve_c_iopt(C.addressof(self._h))

    @property
    def t(self):
        n = self.knots
        t = np.zeros(max(n, 1))
        lib().fitpack_curve_c_get_t(C.addressof(self._h), n, t.ctypes.data)
        return t[:n]

    @property
    def c(self):
        n = max(self.knots - self.degree() - 1, 0)
        c = np.zeros(max(n, 1))
        lib().fitpack_curve_c_get_c(C.addressof(self._h), n, c.ctypes.data)
        return c[:n]

    def set_knots(self, t):
        t = _f64(t)
        return lib().fitpack_curve_c_set_knots(C.addressof(self._h), t.size, t.ctypes.data)
