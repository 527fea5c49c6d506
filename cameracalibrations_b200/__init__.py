"""cameracalibrations_b200 -- B200-native evaluation of a fitted camera calibration.

The data-parallel hot path of yakir12/CameraCalibrations (pixel<->world maps, full-frame
rectification, reprojection residual + Jacobian) as hand-written sm_100a CUDA kernels
behind a C ABI (include/camcal_b200.h); this package is the host-side mirror of the
reference's Julia interface over that ABI.  No CPU fallback exists.
"""
import importlib

from .shard import shard_range, shard_frames      # pure host logic: importable without the library

# Everything else binds libcamcal_b200.so.  It is resolved on first use (PEP 562) so that the
# sharding helpers stay importable on a box without the built library; touching any compute name
# without it raises ImportError -- there is no fallback.
_LAZY = {
    "_lib": ("CamcalError", "Context", "context", "device_count", "LIB_PATH", "EXPORTS"),
    "calibration": ("Calibration", "rectification", "get_ratio", "get_axes", "image_transformations", "warp",
                    "warp_views", "rectify_map", "load_jpegs", "jpeg_info", "reproj_jtj", "calculate_errors", "allreduce_shared", "save", "load", "views_tensor"),
    "fit": ("fit", "detect_fit", "fit_model"),
    "plotting": ("plot",),
    "lm": ("lm_fit", "lm_fit_host", "lm_fit_device", "initial_guess", "initial_guess_device"),
}
_WHERE = {name: mod for mod, names in _LAZY.items() for name in names}


def __getattr__(name):
    if name in _WHERE or name in _LAZY:
        mod = importlib.import_module("." + _WHERE.get(name, name), __name__)
        val = mod if name in _LAZY else getattr(mod, name)
        globals()[name] = val
        return val
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")

RowCol = "SVector{2}: (row, col) -- arrays of shape (..., 2)"
XYZ = "SVector{3}: (x, y, z) -- arrays of shape (..., 3)"

__all__ = ["Calibration", "rectification", "fit", "detect_fit", "get_ratio", "get_axes",
           "image_transformations", "warp", "warp_views", "rectify_map", "load_jpegs", "jpeg_info", "reproj_jtj", "calculate_errors", "save",
           "load", "views_tensor", "shard_range", "shard_frames", "CamcalError", "Context", "context",
           "device_count", "RowCol", "XYZ", "lm_fit", "lm_fit_host", "lm_fit_device", "initial_guess",
           "initial_guess_device", "allreduce_shared", "fit_model", "plot"]
