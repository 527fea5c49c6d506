"""Generate tests/golden/*.json from the reference's own example images with OpenCV.

Run HERE (the build container), where /root/reference is mounted:
    python tests/golden/make_golden.py
The outputs are committed; nothing at test/bench time reads /root/reference.

Why cv2: the reference (Julia, not installable here) delegates corner detection and the
fit to OpenCV (src/detect_fit.jl:10-18,47).  The same OpenCV calls with the same flags,
fed the same (transposed -- src/detect_fit.jl:7) image, give the parameter block the
reference's `fit` would build its Calibration from, and `cv2.projectPoints` gives
known-answer pixels and Jacobians for the forward chain at those parameters.

What is recorded
  example_fit.json     corners (cv2 sub-pixel), the max |corner - corners.json| (the
                       reference's golden fixture, 1 px criterion of test/runtests.jl:63),
                       fitted intrinsics/extrinsics, cv2's own RMS
  project_points.json  cv2.projectPoints pixels + 2x10 Jacobians (fx folded into fy with
                       aspectRatio, as CALIB_FIX_ASPECT_RATIO does) for every view
  cubic_roots.json     known-answer largest real roots of x^3 - x^2 - c from numpy.roots
                       (= the companion-eigenvalue method of src/meta.jl:53-55)
"""
import glob
import json
import os

import cv2
import numpy as np

REF = "/root/reference/test/example"
OUT = os.path.dirname(os.path.abspath(__file__))
N_CORNERS = (5, 8)
CHECKER = 1.0


def detect(path):
    img = cv2.imread(path, cv2.IMREAD_GRAYSCALE)
    gry = np.ascontiguousarray(img.T)  # OpenCV x == Julia row  (src/detect_fit.jl:7)
    flags = (cv2.CALIB_CB_ADAPTIVE_THRESH + cv2.CALIB_CB_FAST_CHECK + cv2.CALIB_CB_EXHAUSTIVE
             + cv2.CALIB_CB_ACCURACY)
    ok, corners = cv2.findChessboardCorners(gry, N_CORNERS, flags=flags)
    assert ok, path
    crit = (cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_MAX_ITER, 30, 0.001)
    corners = cv2.cornerSubPix(gry, corners, (11, 11), (-1, -1), crit)
    return img.shape, corners.reshape(-1, 2).astype(np.float64)  # (row, col), a fastest


def main():
    files = sorted(glob.glob(os.path.join(REF, "*.png")))
    ref_corners = json.load(open(os.path.join(REF, "corners.json")))
    sz = None
    corners = []
    worst = 0.0
    for f in files:
        s, c = detect(f)
        sz = s
        corners.append(c)
        tgt = np.asarray(ref_corners[os.path.basename(f)], dtype=np.float64).reshape(-1, 2)
        worst = max(worst, float(np.max(np.linalg.norm(tgt - c, axis=1))))
    assert worst < 1.0, worst  # test/runtests.jl:63

    n1, n2 = N_CORNERS
    obj = np.array([[a, b, 0.0] for b in range(n2) for a in range(n1)])  # src/detect_fit.jl:69
    objp = [obj.astype(np.float32)] * len(files)
    imgp = [c.astype(np.float32).reshape(-1, 1, 2) for c in corners]
    aspect = 1.0
    K0 = np.eye(3, dtype=np.float32)
    K0[:, 0] = aspect  # cammat[1,:] .= aspect, read transposed (src/detect_fit.jl:34-36)
    flags = (cv2.CALIB_ZERO_TANGENT_DIST + cv2.CALIB_FIX_K3 + cv2.CALIB_FIX_K2
             + cv2.CALIB_FIX_ASPECT_RATIO)
    crit = (cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_MAX_ITER, 30, 0.001)
    rms, K, dist, rvecs, tvecs = cv2.calibrateCamera(
        objp, imgp, (sz[0], sz[1]), K0.astype(np.float64), np.zeros(5), flags=flags, criteria=crit)
    intr = dict(frow=K[0, 0], fcol=K[1, 1], crow=K[0, 2], ccol=K[1, 2],
                k=float(dist.ravel()[0]), checker_size=CHECKER)
    views = [dict(rvec=[float(x) for x in r.ravel()], tvec=[float(x) for x in t.ravel()])
             for r, t in zip(rvecs, tvecs)]
    json.dump(dict(
        source="cv2 %s on /root/reference/test/example/*.png (transposed), flags of "
               "src/detect_fit.jl:10-18,40" % cv2.__version__,
        files=[os.path.basename(f) for f in files], sz=list(sz), n_corners=list(N_CORNERS),
        checker_size=CHECKER, aspect=aspect, corners_json_max_dist=worst, cv2_rms=float(rms),
        intr=intr, views=views, corners=[c.tolist() for c in corners], obj=obj.tolist()),
        open(os.path.join(OUT, "example_fit.json"), "w"), indent=1)

    # known-answer forward projection + Jacobian at the fitted parameters
    pp = []
    d5 = np.array([intr["k"], 0, 0, 0, 0.0])
    for r, t in zip(rvecs, tvecs):
        pix, jac = cv2.projectPoints(obj, r, t, K, d5, aspectRatio=aspect)
        # columns: rvec 0-2, tvec 3-5, fx 6 (==0 with aspectRatio), fy 7, cx 8, cy 9, k1 10
        J10 = np.concatenate([jac[:, 0:6], jac[:, 7:8], jac[:, 8:10], jac[:, 10:11]], axis=1)
        assert np.all(jac[:, 6] == 0)
        pp.append(dict(pix=pix.reshape(-1, 2).tolist(), jac=J10.reshape(-1, 2, 10).tolist()))
    # a second, strongly distorted synthetic camera so k matters
    syn_intr = dict(frow=1400.0 * 1.1, fcol=1400.0, crow=540.0, ccol=960.0, k=-0.12,
                    checker_size=2.5)
    syn_view = dict(rvec=[0.15, -0.1, 0.02], tvec=[-8.0, -12.0, 30.0])
    Ks = np.array([[syn_intr["frow"], 0, syn_intr["crow"]], [0, syn_intr["fcol"], syn_intr["ccol"]],
                   [0, 0, 1.0]])
    rng = np.random.default_rng(5)
    syn_obj = np.concatenate([rng.uniform(-20, 40, (64, 2)), rng.uniform(-1, 1, (64, 1))], axis=1)
    pix, jac = cv2.projectPoints(syn_obj / syn_intr["checker_size"], np.array(syn_view["rvec"]),
                                 np.array(syn_view["tvec"]), Ks,
                                 np.array([syn_intr["k"], 0, 0, 0, 0.0]), aspectRatio=1.1)
    J10 = np.concatenate([jac[:, 0:6], jac[:, 7:8], jac[:, 8:10], jac[:, 10:11]], axis=1)
    json.dump(dict(source="cv2.projectPoints %s" % cv2.__version__, example=pp,
                   synthetic=dict(intr=syn_intr, view=syn_view, aspect=1.1, obj=syn_obj.tolist(),
                                  pix=pix.reshape(-1, 2).tolist(),
                                  jac=J10.reshape(-1, 2, 10).tolist())),
              open(os.path.join(OUT, "project_points.json"), "w"))

    cs = [0.02, 5.0, -0.05, -0.148, -0.2, 1e-8, -1e-8, 0.3, 1.0, 40.0, -0.1, -0.14, 1e3, -3.0]
    roots = []
    for c in cs:
        rs = np.roots([-1.0, 1.0, 0.0, c])
        roots.append(float(np.max(rs[np.abs(rs.imag) < 1e-10].real)))
    json.dump(dict(source="numpy.roots %s" % np.__version__, c=cs, root=roots),
              open(os.path.join(OUT, "cubic_roots.json"), "w"), indent=1)
    print("rms", rms, "worst corner dist", worst, intr)


if __name__ == "__main__":
    main()
