"""ctypes binding of libcamcal_b200.so (the C ABI in include/camcal_b200.h).

There is no CPU fallback anywhere in this package: if the shared library is missing
(`python __graft_entry__.py build` / `make -C cameracalibrations_b200/csrc` builds it)
importing this module raises, and every compute entry point fails with CC_ERR_NO_DEVICE
when no B200 is visible.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# CAMCAL_B200_LIB: load another build of the same ABI (tuning variants under profiles/variants/)
LIB_PATH = os.environ.get("CAMCAL_B200_LIB") or os.path.join(_HERE, "libcamcal_b200.so")

CC_OK = 0
CC_ERR_INVALID_ARG = -1
CC_ERR_NO_DEVICE = -2
CC_ERR_CUDA = -3
CC_ERR_UNSUPPORTED = -4
CC_ERR_NOMEM = -5

COORD_F64 = 0
COORD_F32 = 1
GATHER_AUTO = 0
GATHER_DIRECT = 16
GATHER_TMA = 32

PER_VIEW = 66
SHARED = 21
LM_YZ = 30
LM_SCHUR = 21
LM_DELTA = 8


class CamcalError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"libcamcal_b200 status {status}: {message}")
        self.status = status


class Intr(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("frow", "fcol", "crow", "ccol", "k", "checker_size")]


class View(C.Structure):
    _fields_ = [("rvec", C.c_double * 3), ("tvec", C.c_double * 3)]


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(needs nvcc). cameracalibrations_b200 has no CPU fallback.")

lib = C.CDLL(LIB_PATH)

_vp, _sz, _i, _u, _d, _f = C.c_void_p, C.c_size_t, C.c_int, C.c_uint, C.c_double, C.c_float
_pI, _pV = C.POINTER(Intr), C.POINTER(View)
_i64p = C.POINTER(C.c_int64)
_u8p = C.POINTER(C.c_uint8)

_SIGS = {
    "cc_abi_version": (_i, []),
    "cc_last_error_string": (C.c_char_p, []),
    "cc_device_count": (_i, [C.POINTER(_i)]),
    "cc_ctx_create": (_i, [_i, C.POINTER(_vp)]),
    "cc_ctx_destroy": (_i, [_vp]),
    "cc_ctx_device": (_i, [_vp, C.POINTER(_i)]),
    "cc_ctx_synchronize": (_i, [_vp]),
    "cc_ctx_launch_count": (_i, [_vp, C.POINTER(C.c_uint64)]),
    "cc_host_alloc": (_i, [C.POINTER(_vp), _sz]),
    "cc_host_free": (_i, [_vp]),
    "cc_host_register": (_i, [_vp, _sz]),
    "cc_host_unregister": (_i, [_vp]),
    "cc_img2world_f64": (_i, [_vp, _pI, _pV, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "cc_img2world_f32": (_i, [_vp, _pI, _pV, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "cc_img2world_f64_host": (_i, [_vp, _pI, _pV, _vp, _vp, _vp, _vp, _vp, _sz]),
    "cc_img2world_f32_host": (_i, [_vp, _pI, _pV, _vp, _vp, _vp, _vp, _vp, _sz]),
    "cc_world2img_f64": (_i, [_vp, _pI, _pV, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "cc_world2img_f32": (_i, [_vp, _pI, _pV, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "cc_world2img_f64_host": (_i, [_vp, _pI, _pV, _vp, _vp, _vp, _vp, _vp, _sz]),
    "cc_world2img_f32_host": (_i, [_vp, _pI, _pV, _vp, _vp, _vp, _vp, _vp, _sz]),
    "cc_rectify_f32c1": (_i, [_vp, _pI, _pV, _d, _i64p, _vp, _vp, _i, _i, _sz, _sz, _i, _f, _u, _vp]),
    "cc_rectify_u8c3": (_i, [_vp, _pI, _pV, _d, _i64p, _vp, _vp, _i, _i, _sz, _sz, _i, _u8p, _u, _vp]),
    "cc_rectify_f32c1_host": (_i, [_vp, _pI, _pV, _d, _i64p, _vp, _vp, _i, _i, _sz, _sz, _i, _f, _u]),
    "cc_rectify_u8c3_host": (_i, [_vp, _pI, _pV, _d, _i64p, _vp, _vp, _i, _i, _sz, _sz, _i, _u8p, _u]),
    "cc_rectify_map_f64": (_i, [_vp, _pI, _pV, _d, _i64p, _vp, _vp, _i, _i, _sz, _vp]),
    "cc_get_ratio": (_i, [_vp, _vp, _i, _i, _d, C.POINTER(_d)]),
    "cc_get_axes": (_i, [_d, _d, _i, _i, _i, _i, _i64p]),
    "cc_reproj_jtj_f64": (_i, [_vp, _pI, _d, _vp, _i, _vp, _vp, _i, _vp, _vp, _vp]),
    "cc_reproj_jtj_f64_host": (_i, [_vp, _pI, _d, _vp, _i, _vp, _vp, _i, _vp, _vp]),
    "cc_calculate_errors_f64": (_i, [_vp, _pI, _vp, _i, _vp, _vp, _i, _i, _vp, _vp, _i, _vp, _vp]),
    "cc_lm_schur_f64": (_i, [_vp, _vp, _i, _d, _vp, _vp, _vp]),
    "cc_lm_update_f64": (_i, [_vp, _vp, _vp, _d, _u, _vp, _vp, _i, _vp, _vp, _vp]),
    "cc_lm_fit_f64_host": (_i, [_vp, _pI, _d, _u, _vp, _i, _vp, _vp, _i, _i, _d, C.POINTER(_d), C.POINTER(_i)]),
}

EXPORTS = tuple(_SIGS)

for _name, (_res, _args) in _SIGS.items():
    _fn = getattr(lib, _name)  # AttributeError here = header/library mismatch
    _fn.restype = _res
    _fn.argtypes = _args


def last_error() -> str:
    return lib.cc_last_error_string().decode("utf-8", "replace")


def check(status: int) -> None:
    if status != CC_OK:
        raise CamcalError(status, last_error())


def make_intr(frow, fcol, crow, ccol, k, checker_size) -> Intr:
    return Intr(float(frow), float(fcol), float(crow), float(ccol), float(k), float(checker_size))


def make_view(rvec, tvec) -> View:
    v = View()
    v.rvec[:] = [float(x) for x in rvec]
    v.tvec[:] = [float(x) for x in tvec]
    return v


class Context:
    """cc_ctx for one device (per-device scratch + host pipeline)."""

    def __init__(self, device: int = 0):
        h = _vp()
        check(lib.cc_ctx_create(int(device), C.byref(h)))
        self._h = h
        self.device = int(device)

    @property
    def handle(self):
        return self._h

    def synchronize(self):
        check(lib.cc_ctx_synchronize(self._h))

    def launch_count(self) -> int:
        n = C.c_uint64()
        check(lib.cc_ctx_launch_count(self._h, C.byref(n)))
        return int(n.value)

    def close(self):
        if getattr(self, "_h", None):
            lib.cc_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_contexts: dict[int, Context] = {}


def context(device: int = 0) -> Context:
    ctx = _contexts.get(device)
    if ctx is None:
        ctx = _contexts[device] = Context(device)
    return ctx


def device_count() -> int:
    n = _i(0)
    check(lib.cc_device_count(C.byref(n)))
    return int(n.value)
