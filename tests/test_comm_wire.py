"""CPU: the comm wire format of a curve handle (fitpack_curve_c_comm_size/pack/expand) against an
independent restatement of the reference's layout (src/fitpack_fitters.f90:105-147,
src/fitpack_curves.f90:821-895, array helpers src/fitpack_core.F90:18654-18956), and the
test_curve_comm_roundtrip gate of the reference (test/fitpack_curve_tests.f90:1342-1432) as far as it
needs no fit: pack -> expand into a fresh object -> identical state.  (The fitted-curve round trip
runs in the GPU suite.)"""
import struct

import numpy as np

from fitpack_b200.curve import FitpackCurve

NOT_ALLOC = -99999999


def _hdr(lo, hi):
    return np.frombuffer(struct.pack("<ii", lo, hi), dtype=np.float64)


def _arr(a):
    if a is None:
        return [_hdr(NOT_ALLOC, NOT_ALLOC)]
    a = np.asarray(a)
    if a.dtype == np.int32:
        pad = np.zeros((a.size + 1) // 2 * 2, np.int32)
        pad[:a.size] = a
        return [_hdr(1, a.size), pad.view(np.float64)]
    return [_hdr(1, a.size), a.astype(np.float64)]


def expected_buffer(iopt, smoothing, fp, lwrk, liwrk, c, wrk, iwrk, m, order, knots, bc, nest, xl, xr,
                    x, y, sp, w, t, fou_rows):
    parts = [np.array([iopt, smoothing, fp, lwrk, liwrk], float)] + _arr(c) + _arr(wrk) + _arr(iwrk)
    parts += [np.array([m, order, knots, bc, nest, xl, xr], float)]
    for a in (x, y, sp, w, t):
        parts += _arr(a)
    if fou_rows is None:
        parts += [_hdr(NOT_ALLOC, NOT_ALLOC), _hdr(NOT_ALLOC, NOT_ALLOC)]
    else:
        parts += [_hdr(1, fou_rows), _hdr(1, 2), np.zeros(2 * fou_rows)]
    return np.concatenate(parts)


def test_unallocated_curve_wire_format():
    cv = FitpackCurve()
    buf = cv.comm_pack()
    want = expected_buffer(0, 1000.0, 0.0, 0, 0, None, None, None, 0, 3, 0, 3, 0, 0.0, 0.0,
                           None, None, None, None, None, None)
    assert buf.size == cv.comm_size() == want.size == 5 + 3 + 7 + 5 + 2
    assert buf.tobytes() == want.tobytes()


def test_new_points_wire_format_and_roundtrip():
    rng = np.random.default_rng(5)
    m = 37
    x = rng.uniform(0, 3, m)
    y = np.sin(x)
    w = rng.uniform(0.5, 2, m)
    cv = FitpackCurve()
    cv.new_points(x, y, w)
    order = np.argsort(x, kind="stable")
    nest = max(102, int(np.ceil(np.float32(1.4) * np.float32(m))))
    lwrk = m * 6 + nest * 33
    want = expected_buffer(0, 1000.0, 0.0, lwrk, nest, np.zeros(nest), np.zeros(lwrk), np.zeros(nest, np.int32),
                           m, 3, 0, 3, nest, x.min(), x.max(), x[order], y[order], np.zeros(m), w[order],
                           np.zeros(nest), nest)
    buf = cv.comm_pack()
    assert buf.size == cv.comm_size() == want.size
    assert buf.tobytes() == want.tobytes()
    # odd-length int32 payload is padded to a whole double
    assert (1 + (nest + 1) // 2) == len(np.concatenate(_arr(np.zeros(nest, np.int32))))
    other = FitpackCurve()
    other.comm_expand(buf)
    assert other.comm_size() == buf.size
    assert other.comm_pack().tobytes() == buf.tobytes()
    assert other.degree() == 3 and other.get_bc() == 3 and other.smoothing() == 1000.0
    # expanding over an object that already holds data replaces it completely
    third = FitpackCurve()
    third.new_points(np.arange(5.0), np.arange(5.0))
    third.comm_expand(buf)
    assert third.comm_pack().tobytes() == buf.tobytes()
