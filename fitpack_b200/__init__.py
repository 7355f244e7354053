"""fitpack_b200 -- B200-native batched spline engine for FITPACK's curfit/splev hot path.

Python surface (thin; the product is the C-ABI library, include/fitpack_b200.h):
  core    : curfit / splev with the reference's flat argument conventions (host arrays)
  curve   : FitpackCurve, mirror of the reference's fitpack_curve type / fpCurve class
  batch   : Engine -- batched device-pointer calls on torch CUDA tensors
"""
from ._lib import FitpackB200Error, lib, ensure_built  # noqa: F401
from .core import (FITPACK_SUCCESS, FITPACK_MESSAGE, curfit, splev, curfit_batch,  # noqa: F401
                   splev_batch, regrid, bispev, regrid_lwrk, splder, spalde, splint, fourco, cualde, insert, sproot, parcur, curev, parcur_batch, parcur_lwrk, percur, percur_batch, percur_lwrk, IOPT_NEW_LEASTSQUARES, IOPT_NEW_SMOOTHING, IOPT_OLD_FIT,
                   OUTSIDE_EXTRAPOLATE, OUTSIDE_ZERO, OUTSIDE_NOT_ALLOWED, OUTSIDE_NEAREST_BND)
from .curve import FitpackCurve  # noqa: F401

__all__ = ["curfit", "splev", "curfit_batch", "splev_batch", "regrid", "bispev", "FitpackCurve", "FITPACK_SUCCESS",
           "FITPACK_MESSAGE", "FitpackB200Error", "lib", "ensure_built"]
