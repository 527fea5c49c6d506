"""cameracalibrations_b200 -- B200-native evaluation of a fitted camera calibration.

The data-parallel hot path of yakir12/CameraCalibrations (pixel<->world maps, full-frame
rectification, reprojection residual + Jacobian) as hand-written sm_100a CUDA kernels
behind a C ABI (include/camcal_b200.h); this package is the host-side mirror of the
reference's Julia interface over that ABI.  No CPU fallback exists.
"""
from ._lib import CamcalError, Context, context, device_count, LIB_PATH, EXPORTS
from .calibration import (Calibration, rectification, get_ratio, get_axes, image_transformations,
                          warp, rectify_map, reproj_jtj, calculate_errors, save, load, views_tensor)
from .fit import fit, detect_fit, fit_model
from .lm import lm_fit, lm_fit_host, initial_guess
from .shard import shard_range, shard_frames

RowCol = "SVector{2}: (row, col) -- arrays of shape (..., 2)"
XYZ = "SVector{3}: (x, y, z) -- arrays of shape (..., 3)"

__all__ = ["Calibration", "rectification", "fit", "detect_fit", "get_ratio", "get_axes",
           "image_transformations", "warp", "rectify_map", "reproj_jtj", "calculate_errors", "save",
           "load", "views_tensor", "shard_range", "shard_frames", "CamcalError", "Context", "context",
           "device_count", "RowCol", "XYZ", "lm_fit", "lm_fit_host", "initial_guess", "fit_model"]
