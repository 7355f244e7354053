"""ctypes loader for libfitpack_b200.so (the C-ABI of include/fitpack_b200.h).

The library is the product: there is no Python/CPU fallback.  If it is missing it is
built in-tree with nvcc (cross-compiles without a GPU); if that fails, importing the
compute API raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FITPACK_B200_LIB") or os.path.join(_HERE, "libfitpack_b200.so")

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int32)

# every symbol include/fitpack_b200.h declares: name -> (restype, argtypes)
_I32, _I64, _F64, _VP = C.c_int32, C.c_int64, C.c_double, C.c_void_p
SYMBOLS = {
    "FITPACK_SUCCESS_c": (C.c_bool, [_I32]),
    "fitpack_message_c": (None, [_I32, C.c_char_p]),
    "fp_curfit_c": (None, [_I32, _I32, _VP, _VP, _VP, _F64, _F64, _I32, _F64, _I32, _VP, _VP, _VP,
                           _VP, _VP, _I32, _VP, _VP]),
    "fp_splev_c": (None, [_VP, _I32, _VP, _I32, _VP, _VP, _I32, _I32, _VP]),
    "fitpack_curve_c_allocate": (None, [_VP]),
    "fitpack_curve_c_destroy": (None, [_VP]),
    "fitpack_curve_c_copy": (None, [_VP, _VP]),
    "fitpack_curve_c_move_alloc": (None, [_VP, _VP]),
    "fitpack_curve_c_new_points": (None, [_VP, _I32, _VP, _VP, _VP]),
    "fitpack_curve_c_new_fit": (_I32, [_VP, _I32, _VP, _VP, _VP, _VP]),
    "fitpack_curve_c_fit": (_I32, [_VP, _VP, _VP]),
    "fitpack_curve_c_interpolating": (_I32, [_VP, _VP]),
    "fitpack_curve_c_least_squares": (_I32, [_VP, _VP, _VP]),
    "fitpack_curve_c_eval_one": (_F64, [_VP, _F64, _VP]),
    "fitpack_curve_c_eval_many": (None, [_VP, _I32, _VP, _VP, _VP]),
    "fitpack_curve_c_smoothing": (_F64, [_VP]),
    "fitpack_curve_c_mse": (_F64, [_VP]),
    "fitpack_curve_c_degree": (_I32, [_VP]),
    "fitpack_curve_c_set_bc": (None, [_VP, _I32]),
    "fitpack_curve_c_get_bc": (_I32, [_VP]),
    "fitpack_curve_c_knots": (_I32, [_VP]),
    "fitpack_curve_c_get_t": (_I32, [_VP, _I32, _VP]),
    "fitpack_curve_c_get_c": (_I32, [_VP, _I32, _VP]),
    "fitpack_curve_c_set_knots": (_I32, [_VP, _I32, _VP]),
    "fitpack_curve_c_iopt": (_I32, [_VP]),
    "fitpack_curve_c_comm_size": (_I32, [_VP]),
    "fitpack_curve_c_comm_pack": (None, [_VP, _VP]),
    "fitpack_curve_c_comm_expand": (None, [_VP, _VP]),
    "fpb_engine_create": (_I32, [_I32, C.POINTER(_VP)]),
    "fpb_engine_destroy": (None, [_VP]),
    "fpb_engine_set_stream": (_I32, [_VP, _VP]),
    "fpb_engine_reset_stream": (_I32, [_VP]),
    "fpb_engine_synchronize": (_I32, [_VP]),
    "fpb_engine_launch_count": (_I64, [_VP]),
    "fpb_default_engine_launch_count": (_I64, [_I32]),
    "fpb_engine_set_shared_mode": (_I32, [_VP, _I32]),
    "fpb_engine_set_curev_mode": (_I32, [_VP, _I32]),
    "fpb_engine_set_bispev_mode": (_I32, [_VP, _I32]),
    "fpb_set_single_curve_kernel": (_I32, [_I32]),
    "fpb_host_sharers": (_I32, []),
    "fpb_host_form": (C.c_char_p, [_I64]),
    "fpb_engine_last_kernel_ms": (_F64, [_VP]),
    "fpb_probe_fp64_gops": (_F64, [_VP, _I32]),
    "fpb_selftest_fastmath": (_I32, [_VP, _I64, _I64, _VP]),
    "fpb_device_count": (_I32, []),
    "fpb_last_error": (C.c_char_p, []),
    "fpb_version": (C.c_char_p, []),
    "fp_curfit_batch_dev_c": (_I32, [_VP, _I32, _I32, _I32, _VP, _I64, _I64, _VP, _I64, _I64, _VP,
                                     _I64, _I64, _VP, _VP, _I32, _I32, _VP, _I32, _I32, _VP, _VP,
                                     _VP, _VP, _VP, _VP, _VP, _VP]),
    "fp_curfit_shared_dev_c": (_I32, [_VP, _I32, _I32, _VP, _VP, _VP, _I64, _I64, _F64, _F64, _I32,
                                      _I32, _VP, _VP, _I64, _VP, _VP]),
    "fp_splev_batch_dev_c": (_I32, [_VP, _I32, _VP, _VP, _VP, _I32, _I32, _VP, _VP, _I32, _I32,
                                    _I32, _I32, _VP, _VP]),
    "fp_curfit_batch_c": (_I32, [_I32, _I32, _I32, _VP, _VP, _VP, _I32, _VP, _VP, _I32, _F64, _I32,
                                 _VP, _VP, _VP, _VP, _VP, _I32]),
    "fp_splev_batch_c": (_I32, [_I32, _VP, _VP, _VP, _I32, _I32, _VP, _VP, _I32, _I32, _I32, _VP,
                                _I32]),
    "fpb_transpose_dev": (_I32, [_VP, _VP, _I64, _VP, _I64, _I64, _I64]),
    "fp_regrid_c": (None, [_I32, _I32, _VP, _I32, _VP, _VP, _F64, _F64, _F64, _F64, _I32, _I32, _F64,
                           _I32, _I32, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _I32, _VP, _I32, _VP]),
    "fp_bispev_c": (_I32, [_VP, _I32, _VP, _I32, _VP, _I32, _I32, _VP, _I32, _VP, _I32, _VP, _VP,
                           _I32, _VP, _I32]),
    "fp_regrid_dev_c": (_I32, [_VP, _I32, _I32, _VP, _I32, _VP, _VP, _F64, _F64, _F64, _F64, _I32,
                               _I32, _F64, _I32, _I32, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _I32, _VP,
                               _I32, _VP]),
    "fp_bispev_dev_c": (_I32, [_VP, _VP, _I32, _VP, _I32, _VP, _I32, _I32, _VP, _I32, _VP, _I32,
                               _VP, _VP]),
    "fpb_engine_grid_fpgrre_calls": (_I64, [_VP]),
    "fp_splder_c": (None, [_VP, _I32, _VP, _I32, _I32, _VP, _VP, _I32, _I32, _VP, _VP]),
    "fp_spalde_c": (None, [_VP, _I32, _VP, _I32, _F64, _VP, _VP]),
    "fp_splint_c": (_F64, [_VP, _I32, _VP, _I32, _F64, _F64, _VP]),
    "fp_fourco_c": (None, [_VP, _I32, _VP, _VP, _I32, _VP, _VP, _VP, _VP, _VP]),
    "fp_cualde_c": (None, [_I32, _VP, _I32, _VP, _I32, _I32, _F64, _VP, _I32, _VP]),
    "fp_insert_c": (None, [_I32, _VP, _VP, _VP, _I32, _F64, _I32, _VP]),
    "fp_sproot_c": (None, [_VP, _I32, _VP, _VP, _I32, _VP, _VP]),
    "fitpack_curve_c_fourier": (None, [_VP, _I32, _VP, _VP, _VP, _VP]),
    "fp_parcur_c": (None, [_I32, _I32, _I32, _I32, _VP, _I32, _VP, _VP, _VP, _VP, _I32, _VP, _I32, _VP, _VP,
                           _I32, _VP, _VP, _VP, _I32, _VP, _VP]),
    "fp_curev_c": (None, [_I32, _VP, _I32, _VP, _I32, _I32, _VP, _I32, _VP, _I32, _VP]),
    "fp_parcur_batch_c": (_I32, [_I32, _I32, _I32, _I32, _I32, _VP, _VP, _VP, _VP, _VP, _I32, _F64, _I32,
                                 _VP, _VP, _VP, _VP, _VP, _VP, _VP, _I32]),
    "fp_parcur_batch_dev_c": (_I32, [_VP, _I32, _I32, _I32, _I32, _I32, _VP, _I64, _I64, _VP, _I64, _I64,
                                     _I64, _VP, _I64, _I64, _VP, _VP, _I32, _I32, _VP, _I32, _I32, _VP,
                                     _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "fp_curev_batch_dev_c": (_I32, [_VP, _I32, _I32, _VP, _VP, _VP, _I32, _I64, _I32, _VP, _VP, _I32, _VP]),
    "fp_percur_c": (None, [_I32, _I32, _VP, _VP, _VP, _I32, _F64, _I32, _VP, _VP, _VP, _VP, _VP, _I32, _VP, _VP]),
    "fp_percur_batch_c": (_I32, [_I32, _I32, _I32, _VP, _VP, _VP, _I32, _F64, _I32, _VP, _VP, _VP, _VP, _VP,
                                 _VP, _VP, _I32]),
    "fp_percur_batch_dev_c": (_I32, [_VP, _I32, _I32, _I32, _VP, _I64, _I64, _VP, _I64, _I64, _VP, _I64, _I64,
                                     _I32, _VP, _I32, _I32, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "fitpack_curve_c_integral": (_F64, [_VP, _F64, _F64]),
    "fitpack_curve_c_derivative": (_F64, [_VP, _F64, _I32, _VP]),
    "fitpack_curve_c_all_derivatives": (_I32, [_VP, _F64, _VP]),
}

_lib = None


def ensure_built(force=False):
    if force or not os.path.exists(LIB_PATH):
        from . import build as _b
        _b.build(force=force)
    return LIB_PATH


def lib():
    """Load the C-ABI library and bind every declared symbol (raises if any is missing)."""
    global _lib
    if _lib is None:
        ensure_built()
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the export is missing
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error():
    return lib().fpb_last_error().decode()


class FitpackB200Error(RuntimeError):
    pass


def check(rc, what):
    if rc != 0:
        raise FitpackB200Error("%s failed (status %d): %s" % (what, rc, last_error()))
