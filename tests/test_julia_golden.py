"""Parity against outputs of the REAL CameraCalibrations.jl (tests/golden/julia_*.json, written by
julia/make_golden.jl on a box that has Julia and the package).  This image has no Julia, so the
files do not exist yet: every test here then reports an EXPECTED FAILURE with the reason "parity
unpinned" -- visible in every test run -- and turns into a real check the moment the files are
committed.  CPU tests check the oracle, the `gpu` tests check the CUDA path through the C ABI.

Tolerances: FP64 point maps 1e-9 (north_star); calculate_errors 1e-9 relative (inverse: the
reference draws its own random samples, so only its size is checked); warp outputs bit-exact
(float32 values and N0f8 bytes), fill decisions included.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden
from oracle import oracle_c as oc

FILES = ("julia_points.json", "julia_errors.json", "julia_warp.json", "julia_calibration.json")
HAVE = all(os.path.exists(os.path.join(GOLDEN, f)) for f in FILES)
unpinned = pytest.mark.xfail(not HAVE, run=HAVE, strict=False,
                             reason="parity unpinned: tests/golden/julia_*.json absent -- run julia/make_golden.jl "
                                    "with the real CameraCalibrations.jl and commit its output")


def _fit(example_fit):
    return example_fit["intr_tuple"], example_fit["view_list"]


def _small_camera(example_fit, scale_px):
    i = example_fit["intr_tuple"]
    return (i[0] * scale_px, i[1] * scale_px, i[2] * scale_px, i[3] * scale_px, i[4], i[5])


def _oracle_errors(intr, views, example_fit, inverse_samples):
    """calculate_errors of the oracle (it applies the reference's normalisations,
    src/buildcalibrations.jl:60-65, itself) on its own uniform samples rc in [1, sz]."""
    nv = len(views)
    sz = np.asarray(example_fit["sz"], dtype=np.float64)
    samples = np.random.default_rng(1).random((nv, inverse_samples, 2)) * (sz - 1) + 1
    e = oc.calculate_errors(intr, views, example_fit["obj_np"], example_fit["corners_np"],
                            tuple(example_fit["n_corners"]), samples)
    return dict(n=nv, reprojection=e[0], projection=e[1], distance=e[2], inverse=e[3])


def _check_points(world_fn, back_fn, rect_fn, d, nviews):
    pix = np.asarray(d["pix"], dtype=np.float64)
    for v in d["views"]:
        i = v["view"] - 1
        xyz = np.asarray(v["xyz"])
        got = world_fn(i, pix)
        assert np.max(np.abs(got - xyz)) <= 1e-9
        assert np.max(np.abs(back_fn(i, xyz) - np.asarray(v["back"]))) <= 1e-9
        assert np.max(np.abs(rect_fn(i, pix) - np.asarray(v["rect"]))) <= 1e-9
    assert len(d["views"]) == nviews


@unpinned
def test_oracle_points_match_julia(example_fit):
    intr, views = _fit(example_fit)
    chains = [oc.chain(intr, *v) for v in views]
    w = lambda i, p: np.stack(oc.img2world_soa(chains[i], p[:, 0], p[:, 1]), -1)
    b = lambda i, q: np.stack(oc.world2img_soa(chains[i], q[:, 0], q[:, 1], q[:, 2]), -1)
    r = lambda i, p: np.stack(oc.img2world_soa(chains[i], p[:, 0], p[:, 1], want_z=False)[:2], -1)
    _check_points(w, b, r, load_golden("julia_points.json"), len(views))


@unpinned
def test_oracle_errors_match_julia(example_fit):
    intr, views = _fit(example_fit)
    d = load_golden("julia_errors.json")
    e = _oracle_errors(intr, views, example_fit, d["inverse_samples"])
    for k in ("reprojection", "projection", "distance"):
        assert abs(e[k] - d[k]) <= 1e-9 * max(1.0, abs(d[k])), k
    assert e["n"] == d["n"] and d["inverse"] < 1e-9 and e["inverse"] < 1e-9


def _warp_cases(example_fit):
    """(chain inputs, ratio, axs_min, frames f32, frames u8, expected f32, expected u8) per dumped case."""
    d = load_golden("julia_warp.json")
    sz = tuple(d["sz"])
    intr_s = _small_camera(example_fit, d["scale_px"])
    img32 = np.asarray(d["img32"], dtype=np.float32).reshape(sz[1], sz[0])          # Julia (sz1, sz2) memory
    img8 = np.asarray(d["img8"], dtype=np.uint8).reshape(sz[1], sz[0], 3)
    for v in d["views"]:
        rv, tv = example_fit["view_list"][v["view"] - 1]
        yield (intr_s, rv, tv, v["ratio"], tuple(v["axs_min"]), img32, img8,
               np.asarray(v["out32"], dtype=np.float64).astype(np.float32).reshape(sz[1], sz[0]),
               np.asarray(v["out8"], dtype=np.uint8).reshape(sz[1], sz[0], 3))
    psz = tuple(d["probe_sz"])
    p32 = np.asarray(d["probe32"], dtype=np.float32).reshape(psz[1], psz[0])
    p8 = np.asarray(d["probe8"], dtype=np.uint8).reshape(psz[1], psz[0], 3)
    ident = (1.0, 1.0, 0.0, 0.0, 0.0, 1.0)
    for v in d["probes"]:
        yield (ident, (0.0, 0.0, 0.0), (0.0, 0.0, 1.0), v["ratio"], tuple(v["axs_min"]), p32, p8,
               np.asarray(v["out32"], dtype=np.float64).astype(np.float32).reshape(psz[1], psz[0]),
               np.asarray(v["out8"], dtype=np.uint8).reshape(psz[1], psz[0], 3))


def _same_f32(a, b):
    return np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(b)], b[~np.isnan(b)])


@unpinned
def test_oracle_warp_matches_julia(example_fit):
    """The bilinear index/weight rule, the fill rule and the N0f8 re-quantisation of
    ImageTransformations.warp (src/plot_calibration.jl:40): bit for bit."""
    for intr, rv, tv, ratio, axs, f32, u8, e32, e8 in _warp_cases(example_fit):
        ch = oc.chain(intr, rv, tv)
        assert _same_f32(oc.rectify_f32c1(ch, 1.0 / ratio, axs, f32[None], fill=np.nan)[0], e32), (ratio, axs)
        assert np.array_equal(oc.rectify_u8c3(ch, 1.0 / ratio, axs, u8[None])[0], e8), (ratio, axs)


@unpinned
def test_ratio_and_axes_match_julia(example_fit):
    d = load_golden("julia_warp.json")
    n1, n2 = example_fit["n_corners"]
    for v in d["views"]:
        ip = example_fit["corners_np"][v["view"] - 1].reshape(n2, n1, 2).transpose(1, 0, 2) * d["scale_px"]
        ratio = oc.get_ratio(ip, 1.0)
        assert abs(ratio - v["ratio"]) <= 1e-12 * v["ratio"]
        assert tuple(oc.get_axes(v["ratio"], 1.0, (n1, n2), tuple(d["sz"]))) == tuple(v["axs_min"])


@unpinned
def test_load_reads_the_reference_json(example_fit):
    """src/io.jl:28-32 as JSON3 really writes it (RotationVec / AffineMap / SDiagonal encodings)."""
    pytest.importorskip("torch")
    import cameracalibrations_b200 as cc
    c = cc.load(os.path.join(GOLDEN, "julia_calibration.json"))
    intr, views = _fit(example_fit)
    assert np.allclose(c.intrinsic, intr[:4], rtol=0, atol=0) and c.k == intr[4]
    for (r, t), (rv, tv) in zip(c.extrinsics, views):
        assert np.array_equal(r, rv) and np.array_equal(t, tv)


# ------------------------------------------------------------------ the CUDA path
@pytest.mark.gpu
@unpinned
def test_gpu_points_match_julia(example_fit):
    torch = pytest.importorskip("torch")
    import cameracalibrations_b200 as cc
    intr, views = _fit(example_fit)
    c = cc.Calibration(intr[:4], views, 1.0 / intr[5], intr[4], example_fit["files"])
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    w = lambda i, p: torch.stack(c.img2world(dev(p[:, 0]), dev(p[:, 1]), i), -1).cpu().numpy()
    b = lambda i, q: torch.stack(c.world2img(dev(q[:, 0]), dev(q[:, 1]), dev(q[:, 2]), i), -1).cpu().numpy()
    r = lambda i, p: cc.rectification(c, i)(dev(p)).cpu().numpy()
    _check_points(w, b, r, load_golden("julia_points.json"), len(views))


@pytest.mark.gpu
@unpinned
def test_gpu_warp_matches_julia(example_fit):
    torch = pytest.importorskip("torch")
    import cameracalibrations_b200 as cc
    for intr, rv, tv, ratio, axs, f32, u8, e32, e8 in _warp_cases(example_fit):
        c = cc.Calibration(intr[:4], [(rv, tv)], 1.0 / intr[5], intr[4], ["extrinsic.png"])
        for gather in ("auto", "direct"):
            got = cc.warp(c, 0, torch.from_numpy(f32[None].copy()).cuda(), ratio, axs, gather=gather).cpu().numpy()[0]
            assert _same_f32(got, e32), (ratio, axs, gather)
            got8 = cc.warp(c, 0, torch.from_numpy(u8[None].copy()).cuda(), ratio, axs, gather=gather).cpu().numpy()[0]
            assert np.array_equal(got8, e8), (ratio, axs, gather)
