"""numpy twin of the C oracle.  TEST INFRASTRUCTURE ONLY (see camcal_oracle.h).

An independent second restatement of the reference's hot path, used to cross-check
``camcal_oracle.c``.  It differs from the C twin on purpose where the reference leaves
room:

* the cubic root is found the way ``src/meta.jl:53-55`` finds it: eigenvalues of the
  companion matrix (``numpy.roots`` == ``Polynomials.roots``), imaginary filter 1e-10,
  maximum real part -- the C twin uses monotone Newton;
* the rotation is ``cv2.Rodrigues``-free textbook Rodrigues written with plain numpy
  matrix algebra (no explicit fma ordering).

The two must agree to <= 1e-12 (tests/test_oracle.py); neither is bit-exact with the
other, only the C twin defines the normative operation order.

All citations are file:line in /root/reference (yakir12/CameraCalibrations v0.7.3).
"""
from __future__ import annotations

import numpy as np


def rodrigues(rvec):
    r = np.asarray(rvec, dtype=np.float64)
    th = np.linalg.norm(r)
    if th < np.finfo(np.float64).eps:
        return np.eye(3)
    n = r / th
    K = np.array([[0, -n[2], n[1]], [n[2], 0, -n[0]], [-n[1], n[0], 0]])
    return np.cos(th) * np.eye(3) + (1 - np.cos(th)) * np.outer(n, n) + np.sin(th) * K


def cubic_root(c: float) -> float:
    """src/meta.jl:52-55: max real root of c + 0x + x^2 - x^3."""
    if c == 0:
        return 1.0
    rs = np.roots([-1.0, 1.0, 0.0, c])
    rrs = rs[np.abs(rs.imag) < 1e-10]
    return float(np.max(rrs.real))


def lens_distortion(v, k):
    """src/meta.jl:39-44"""
    if k == 0:
        return v
    r2 = np.sum(v * v, axis=-1, keepdims=True)
    return (1 + k * r2) * v


def inv_lens_distortion(v2, k):
    """src/meta.jl:50-57 (scalar loop: one eigen-solve per point, like the reference)"""
    if k == 0:
        return v2
    v2 = np.atleast_2d(v2)
    out = np.empty_like(v2)
    for i, p in enumerate(v2):
        c = k * float(p @ p)
        out[i] = p / cubic_root(c)
    return out


class Chain:
    """Calibration(...) for one view: src/meta.jl:27-33, 71-76."""

    def __init__(self, intr, rvec, tvec):
        self.frow, self.fcol, self.crow, self.ccol, self.k, self.cs = [float(x) for x in intr]
        self.R = rodrigues(rvec)
        self.t = np.asarray(tvec, dtype=np.float64)
        self.Rinv = self.R.T
        self.tinv = self.Rinv @ (-self.t)

    def world2img(self, xyz):
        """src/meta.jl:88 / :29"""
        p = np.atleast_2d(np.asarray(xyz, dtype=np.float64))
        q = p * (1.0 / self.cs)
        P = q @ self.R.T + self.t
        uv = P[:, :2] * (1.0 / P[:, 2:3])
        uv = lens_distortion(uv, self.k)
        return uv * np.array([self.frow, self.fcol]) + np.array([self.crow, self.ccol])

    def img2world(self, rc):
        """src/meta.jl:82 / :31"""
        rc = np.atleast_2d(np.asarray(rc, dtype=np.float64))
        a = 1.0 / np.array([self.frow, self.fcol])
        uv = rc * a + a * (-np.array([self.crow, self.ccol]))
        uv = inv_lens_distortion(uv, self.k)
        rc1 = np.concatenate([uv, np.ones((len(uv), 1))], axis=1)
        l = self.Rinv[2, :]
        d = -self.tinv[2] / (rc1 @ l)
        w = d[:, None] * rc1
        p = w @ self.Rinv.T + self.tinv
        return p * (1.0 / (1.0 / self.cs))


def get_ratio(imgpoints, checker_size):
    """src/plot_calibration.jl:8-13; imgpoints: (n1, n2, 2)"""
    ip = np.asarray(imgpoints, dtype=np.float64)
    l = 0.5 * (np.mean(np.linalg.norm(np.diff(ip, axis=0), axis=-1))
               + np.mean(np.linalg.norm(np.diff(ip, axis=1), axis=-1)))
    return l / checker_size


def get_axes(ratio, checker_size, n_corners, sz):
    """src/plot_calibration.jl:1-6 (np.rint == Julia round(Int, .): half-even)"""
    w = np.rint(ratio * checker_size * (np.asarray(n_corners) - 1))
    mn = np.rint((w - np.asarray(sz)) / 2).astype(np.int64)
    return int(mn[0]), int(mn[1])


def rectify_gray(chain, inv_ratio, axs_min, img, fill):
    """warp(img, tform, axs), src/plot_calibration.jl:17-18,40.  ``img`` is indexed
    [r-1, c-1] (numpy array of shape (sz1, sz2), any strides).  Small cases only."""
    sz1, sz2 = img.shape
    I1 = axs_min[0] + np.arange(sz1)
    I2 = axs_min[1] + np.arange(sz2)
    g1, g2 = np.meshgrid(I1.astype(np.float64) * inv_ratio, I2.astype(np.float64) * inv_ratio,
                         indexing="ij")
    xyz = np.stack([g1.ravel(), g2.ravel(), np.zeros(g1.size)], axis=1)
    rc = chain.world2img(xyz)
    out = np.full(sz1 * sz2, fill, dtype=np.float64)
    row, col = rc[:, 0], rc[:, 1]
    ok = (row >= 1) & (row <= sz1) & (col >= 1) & (col <= sz2)
    f1 = np.floor(row[ok]); f1 = np.where(f1 > sz1 - 1, f1 - 1, f1)
    f2 = np.floor(col[ok]); f2 = np.where(f2 > sz2 - 1, f2 - 1, f2)
    d1, d2 = row[ok] - f1, col[ok] - f2
    i1, i2 = f1.astype(np.int64) - 1, f2.astype(np.int64) - 1
    im = img.astype(np.float64)
    a00, a10 = im[i1, i2], im[i1 + 1, i2]
    a01, a11 = im[i1, i2 + 1], im[i1 + 1, i2 + 1]
    out[ok] = (1 - d1) * ((1 - d2) * a00 + d2 * a01) + d1 * ((1 - d2) * a10 + d2 * a11)
    return out.reshape(sz1, sz2), rc.reshape(sz1, sz2, 2)


def calculate_errors(intr, views, obj, imgs, n_corners, inv_samples):
    """src/buildcalibrations.jl:37-67.  obj: (n1*n2, 3), imgs: (nviews, n1*n2, 2),
    inv_samples: (nviews, S, 2) pre-drawn rc in [1, sz]."""
    n1, n2 = n_corners
    cs = float(intr[5])
    rep = pro = dis = inv = 0.0
    for i, (rv, tv) in enumerate(views):
        ch = Chain(intr, rv, tv)
        rep += np.sum((ch.world2img(obj) - imgs[i]) ** 2)
        projected = ch.img2world(imgs[i])
        pro += np.sum((projected - obj) ** 2)
        grid = projected.reshape(n2, n1, 3)  # a fastest -> axis 1
        for ax in (1, 0):
            dis += np.sum((np.linalg.norm(np.diff(grid, axis=ax), axis=-1) - cs) ** 2)
        rc = inv_samples[i]
        inv += np.sum((rc - ch.world2img(ch.img2world(rc))) ** 2)
    nv = len(views)
    n = n1 * n2 * nv
    return (np.sqrt(rep / n), np.sqrt(pro / n),
            np.sqrt(dis / ((n1 - 1) * (n2 - 1)) / nv),
            np.sqrt(inv / inv_samples.shape[1] / nv))
