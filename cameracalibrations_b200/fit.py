"""fit(files, n_corners, checker_size) -> (Calibration, errors): src/buildcalibrations.jl:12-26.

Corner detection and the camera-model fit are OpenCV calls in the reference
(src/detect_fit.jl:5-72, third-party C++; out of the hot-path scope)
and stay OpenCV calls here (cv2), with the same flags and the same transposed image.  What
changes is everything downstream: the calibration object evaluates on the GPU and
calculate_errors is one fused kernel.
"""
from __future__ import annotations

import numpy as np

from .calibration import Calibration, calculate_errors


def _cv2():
    try:
        import cv2
    except Exception as e:  # pragma: no cover
        raise ImportError("fit()/detect_fit() need OpenCV (cv2), like the reference needs OpenCV.jl") from e
    return cv2


def _detect_corners(file, n_corners, sz):
    """src/detect_fit.jl:5-21"""
    cv2 = _cv2()
    img = cv2.imread(file, cv2.IMREAD_GRAYSCALE)
    if img is None or tuple(img.shape) != tuple(sz):
        return None
    gry = np.ascontiguousarray(img.T)            # OpenCV x == first RowCol component
    flags = (cv2.CALIB_CB_ADAPTIVE_THRESH + cv2.CALIB_CB_FAST_CHECK + cv2.CALIB_CB_EXHAUSTIVE
             + cv2.CALIB_CB_ACCURACY)
    ok, corners = cv2.findChessboardCorners(gry, tuple(n_corners), flags=flags)
    if not ok:
        return None
    crit = (cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_MAX_ITER, 30, 0.001)   # CRITERIA
    corners = cv2.cornerSubPix(gry, corners, (11, 11), (-1, -1), crit)
    return file, corners.reshape(-1, 2).astype(np.float64)


def fit_model(sz, objpoints, imgpointss, n_corners, with_distortion=True, aspect=1, solver="opencv"):
    """fit_model, src/detect_fit.jl:27-61.  objpoints (nc, 3) board units; imgpointss (nv, nc, 2).
    solver="opencv": OpenCV.calibrateCamera with the reference's flags and CRITERIA;
    solver="b200": the same model fitted by the device LM (lm.py / csrc/lm.cu).
    Returns dict(k, Rs, ts, frow, fcol, crow, ccol, rms)."""
    objpoints = np.asarray(objpoints, dtype=np.float64)
    imgpointss = np.asarray(imgpointss, dtype=np.float64)
    if solver == "b200":
        from . import lm
        intr0, views0 = lm.initial_guess_device(objpoints, imgpointss, sz, float(aspect))
        r = lm.lm_fit_device(intr0, views0, objpoints, imgpointss, aspect=float(aspect), with_distortion=with_distortion)
        return dict(k=float(r["intr"][4]), Rs=[v[:3] for v in r["views"]], ts=[v[3:] for v in r["views"]],
                    frow=r["intr"][0], fcol=r["intr"][1], crow=r["intr"][2], ccol=r["intr"][3], rms=r["rms"])
    assert solver == "opencv", solver
    cv2 = _cv2()
    K0 = np.eye(3)
    K0[:, 0] = aspect
    flags = (cv2.CALIB_ZERO_TANGENT_DIST + cv2.CALIB_FIX_K3 + cv2.CALIB_FIX_K2
             + (0 if with_distortion else cv2.CALIB_FIX_K1) + cv2.CALIB_FIX_ASPECT_RATIO)
    crit = (cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_MAX_ITER, 30, 0.001)
    rms, K, dist, rvecs, tvecs = cv2.calibrateCamera(
        [objpoints.astype(np.float32)] * len(imgpointss),
        [c.astype(np.float32).reshape(-1, 1, 2) for c in imgpointss], (sz[0], sz[1]), K0, np.zeros(5),
        flags=flags, criteria=crit)
    return dict(k=float(dist.ravel()[0]), Rs=[r.ravel() for r in rvecs], ts=[t.ravel() for t in tvecs],
                frow=K[0, 0], fcol=K[1, 1], crow=K[0, 2], ccol=K[1, 2], rms=float(rms))


def detect_fit(files, n_corners, with_distortion=True, aspect=1, solver="opencv"):
    """src/detect_fit.jl:63-72"""
    cv2 = _cv2()
    first = cv2.imread(files[0], cv2.IMREAD_GRAYSCALE)
    sz = tuple(first.shape)
    found = [r for r in (_detect_corners(f, n_corners, sz) for f in files) if r is not None]
    assert found, "No checkers were detected in any of the images, perhaps try a different `n_corners`."
    files = [f for f, _ in found]
    imgpointss = np.stack([c for _, c in found])
    n1, n2 = n_corners
    objpoints = np.array([[a, b, 0.0] for b in range(n2) for a in range(n1)])
    m = fit_model(sz, objpoints, imgpointss, n_corners, with_distortion, aspect, solver=solver)
    return dict(files=files, objpoints=objpoints, imgpointss=imgpointss, sz=sz, **m)


def fit(files, n_corners, checker_size, aspect=1, with_distortion=True, inverse_samples=100,
        with_plot=False, rng=None, solver="opencv"):
    """Returns the tuple (c, eps) like the reference."""
    files = list(dict.fromkeys(files))           # unique(files)
    d = detect_fit(files, n_corners, with_distortion, aspect, solver=solver)
    objpoints = d["objpoints"] * checker_size
    c = Calibration.from_fit(d["Rs"], d["ts"], d["frow"], d["fcol"], d["crow"], d["ccol"],
                             checker_size, d["k"], d["files"])
    eps = calculate_errors(c, d["imgpointss"], objpoints, checker_size, d["sz"], d["files"], n_corners,
                           inverse_samples, rng=rng)
    if with_plot:                                # src/buildcalibrations.jl:22: plot(c, imgpointss, n_corners, checker_size, sz)
        from .plotting import plot
        plot(c, d["imgpointss"], n_corners, checker_size, d["sz"])
    return c, eps
