"""A third restatement of the world->pixel chain, in the operation order the reference's Julia libraries use.
TEST INFRASTRUCTURE ONLY (see camcal_oracle.h).

The C oracle (camcal_oracle.c) expands the rotation vector into a matrix once and evaluates the chain with
explicit fma()s -- that order is what the CUDA kernels reproduce bit for bit.  The reference itself
(src/meta.jl:27-33) composes closures over third-party types and applies them one after the other:

  * scale      LinearMap(SDiagonal(1/cs...))          -> elementwise product
  * extrinsic  AffineMap(RotationVec, tvec)           -> Rotations.jl applies a RotationVec to a vector
                                                         through its angle-axis form (Rodrigues' formula ON
                                                         THE VECTOR: ct*v + st*(w x v) + (w.v)(1 - ct)*w, no
                                                         matrix), then adds the translation
  * PerspectiveMap                                    -> scale = 1/v[3]; (v[1]*scale, v[2]*scale)
  * lens_distortion (src/meta.jl:39-44)               -> (1 + k*|v|^2) * v
  * intrinsic  AffineMap(SDiagonal(frow, fcol), c)    -> elementwise product, then the sum

with no fused multiply-adds (Julia does not contract a*b + c).  Written from the published behaviour of
Rotations.jl / CoordinateTransformations.jl (neither is vendored in /root/reference, Julia is absent): like the
C oracle it is NOT pinned to output of the real package.  Its purpose is to MEASURE how far a different but
equally valid operation order moves the results (tests/test_oracle.py::test_operation_order_sensitivity):
coordinates by ~1e-13 px, bilinear tap indices not at all on the test frames.
"""
from __future__ import annotations

import numpy as np


def rotvec_apply(rv, v):
    """RotationVec(sx, sy, sz) * v for v of shape (n, 3): angle-axis application, vector form."""
    sx, sy, sz = (float(x) for x in rv)
    theta = np.sqrt(sx * sx + sy * sy + sz * sz)
    if not theta > np.finfo(np.float64).eps:
        # first-order expansion for tiny angles: v + rv x v
        return np.stack([v[:, 0] + sy * v[:, 2] - sz * v[:, 1],
                         v[:, 1] + sz * v[:, 0] - sx * v[:, 2],
                         v[:, 2] + sx * v[:, 1] - sy * v[:, 0]], axis=1)
    w = np.array([sx / theta, sy / theta, sz / theta])
    w = w / np.sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2])       # AngleAxis normalises its axis
    ct, st = np.cos(theta), np.sin(theta)
    cx = w[1] * v[:, 2] - w[2] * v[:, 1]
    cy = w[2] * v[:, 0] - w[0] * v[:, 2]
    cz = w[0] * v[:, 1] - w[1] * v[:, 0]
    m = (v[:, 0] * w[0] + v[:, 1] * w[1] + v[:, 2] * w[2]) * (1.0 - ct)
    return np.stack([ct * v[:, 0] + st * cx + m * w[0],
                     ct * v[:, 1] + st * cy + m * w[1],
                     ct * v[:, 2] + st * cz + m * w[2]], axis=1)


def world2img(intr, rvec, tvec, xyz):
    """src/meta.jl:29: intrinsic o distort o PerspectiveMap o extrinsic o scale, one closure after the other."""
    frow, fcol, crow, ccol, k, cs = (float(x) for x in intr)
    p = np.atleast_2d(np.asarray(xyz, dtype=np.float64))
    q = p * (1.0 / cs)
    P = rotvec_apply(rvec, q) + np.asarray(tvec, dtype=np.float64)
    s = 1.0 / P[:, 2]
    u, v = P[:, 0] * s, P[:, 1] * s
    if k != 0:
        radial = 1.0 + k * (u * u + v * v)
        u, v = radial * u, radial * v
    return np.stack([frow * u + crow, fcol * v + ccol], axis=1)


def rectify_map(intr, rvec, tvec, inv_ratio, axs_min, sz):
    """tform of src/plot_calibration.jl:17-18 over the output axes: (sz1, sz2, 2) source coordinates."""
    sz1, sz2 = sz
    g1, g2 = np.meshgrid((axs_min[0] + np.arange(sz1)).astype(np.float64) * inv_ratio,
                         (axs_min[1] + np.arange(sz2)).astype(np.float64) * inv_ratio, indexing="ij")
    xyz = np.stack([g1.ravel(), g2.ravel(), np.zeros(g1.size)], axis=1)
    return world2img(intr, rvec, tvec, xyz).reshape(sz1, sz2, 2)
