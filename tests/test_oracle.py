"""CPU tests that pin the oracle (no GPU, no product code).

Pins, in order of strength:
  * cv2.projectPoints pixels + Jacobians (the OpenCV the reference calls,
    src/detect_fit.jl:47) at the parameters cv2 fits on the reference's example images;
  * the reference's own assertions: test/runtests.jl:76 (all four errors < 1) and
    test/runtests.jl:84 (rectification == first two of c(i, 1));
  * numpy.roots known answers for the cubic of src/meta.jl:53-55;
  * the independent numpy twin (oracle/oracle_np.py).
"""
import numpy as np
import pytest

from conftest import load_golden, C2_INTR, SYN_VIEW
from oracle import oracle_c as oc
from oracle import oracle_np as on
from oracle import julia_order as oj


def test_cubic_root_known_answers():
    g = load_golden("cubic_roots.json")
    for c, r in zip(g["c"], g["root"]):
        got = oc.cubic_root(c)
        tol = 1e-7 if abs(c + 4 / 27) < 1e-3 else 1e-12  # LAPACK itself degrades at the double root
        assert abs(got - r) <= tol * max(1.0, abs(r)), (c, got, r)


def test_cubic_root_vs_companion_eigenvalues_random():
    rng = np.random.default_rng(0)
    cs = np.concatenate([rng.uniform(-0.14, 0.5, 400), rng.uniform(0.5, 50, 50),
                         -rng.uniform(0.16, 5, 50), [0.0]])
    for c in cs:
        a, b = oc.cubic_root(c), on.cubic_root(c)
        assert abs(a - b) <= 1e-12 * max(1.0, abs(b)), (c, a, b)
        assert abs(a ** 3 - a ** 2 - c) <= 1e-13 * max(1.0, abs(c))


def test_cubic_root_special_cases():
    assert oc.cubic_root(0.0) == 1.0          # roots {0,0,1}
    assert oc.cubic_root(-0.2) < 0            # beyond invertibility: the only real root
    assert np.isnan(oc.cubic_root(np.nan))


def test_world2img_matches_cv2_projectpoints(example_fit):
    g = load_golden("project_points.json")
    obj = example_fit["obj_np"]
    for (rv, tv), gg in zip(example_fit["view_list"], g["example"]):
        ch = oc.chain(example_fit["intr_tuple"], rv, tv)
        row, col = oc.world2img(ch, obj)
        pix = np.asarray(gg["pix"])
        assert np.max(np.abs(row - pix[:, 0])) < 1e-9
        assert np.max(np.abs(col - pix[:, 1])) < 1e-9
    s = g["synthetic"]
    i = s["intr"]
    ch = oc.chain((i["frow"], i["fcol"], i["crow"], i["ccol"], i["k"], i["checker_size"]),
                  s["view"]["rvec"], s["view"]["tvec"])
    row, col = oc.world2img(ch, np.asarray(s["obj"]))
    pix = np.asarray(s["pix"])
    assert np.max(np.abs(row - pix[:, 0])) < 1e-9 and np.max(np.abs(col - pix[:, 1])) < 1e-9


def test_jacobian_matches_cv2(example_fit):
    g = load_golden("project_points.json")
    obj, img = example_fit["obj_np"], example_fit["corners_np"]
    pv, sh, jac = oc.reproj_jtj(example_fit["intr_tuple"], 1.0, example_fit["view_list"], obj, img,
                                want_jac=True)
    for vi, gg in enumerate(g["example"]):
        J = np.asarray(gg["jac"])
        assert np.max(np.abs(jac[vi] - J) / (1.0 + np.abs(J))) < 1e-9
    # blocks == J'J of the cv2 Jacobian
    res_all = 0.0
    shJ = np.zeros((4, 4)); shr = np.zeros(4)
    for vi, gg in enumerate(g["example"]):
        J = np.asarray(gg["jac"]).reshape(-1, 10)
        r = (np.asarray(gg["pix"]) - img[vi]).reshape(-1)
        JtJ = J.T @ J
        np.testing.assert_allclose(pv[vi, :36].reshape(6, 6), JtJ[:6, :6], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(pv[vi, 36:60].reshape(6, 4), JtJ[:6, 6:], rtol=1e-9, atol=1e-7)
        np.testing.assert_allclose(pv[vi, 60:66], J[:, :6].T @ r, rtol=1e-9, atol=1e-7)
        shJ += JtJ[6:, 6:]; shr += J[:, 6:].T @ r; res_all += r @ r
    np.testing.assert_allclose(sh[:16].reshape(4, 4), shJ, rtol=1e-9, atol=1e-7)
    np.testing.assert_allclose(sh[16:20], shr, rtol=1e-9, atol=1e-7)
    assert abs(sh[20] - res_all) < 1e-9
    # sse / n == cv2's own RMS^2
    n = img.shape[0] * img.shape[1]
    assert abs(np.sqrt(sh[20] / n) - example_fit["cv2_rms"]) < 1e-6  # cv2 fits on Float32 copies


def test_jacobian_synthetic_aspect_and_scale():
    s = load_golden("project_points.json")["synthetic"]
    i = s["intr"]
    intr = (i["frow"], i["fcol"], i["crow"], i["ccol"], i["k"], i["checker_size"])
    obj = np.asarray(s["obj"])
    img = np.asarray(s["pix"])[None]
    pv, sh, jac = oc.reproj_jtj(intr, s["aspect"], [(s["view"]["rvec"], s["view"]["tvec"])], obj, img,
                                want_jac=True)
    J = np.asarray(s["jac"])
    assert np.max(np.abs(jac[0] - J) / (1.0 + np.abs(J))) < 1e-9
    assert sh[20] < 1e-18  # img == projection


def test_reference_accuracy_bounds(example_fit):
    """test/runtests.jl:73-77: all(<(1), (reprojection, projection, distance, inverse))"""
    rng = np.random.default_rng(1)
    sz = np.asarray(example_fit["sz"], dtype=np.float64)
    nv = len(example_fit["view_list"])
    samples = rng.random((nv, 100, 2)) * (sz - 1) + 1   # src/buildcalibrations.jl:54
    eps = oc.calculate_errors(example_fit["intr_tuple"], example_fit["view_list"],
                              example_fit["obj_np"], example_fit["corners_np"],
                              example_fit["n_corners"], samples)
    assert all(e < 1 for e in eps)
    assert abs(eps[0] - example_fit["cv2_rms"]) < 1e-6
    assert eps[3] < 1e-10            # inverse o forward round trip, px RMS
    eps_np = on.calculate_errors(example_fit["intr_tuple"], example_fit["view_list"],
                                 example_fit["obj_np"], example_fit["corners_np"],
                                 example_fit["n_corners"], samples)
    np.testing.assert_allclose(eps[:3], eps_np[:3], rtol=1e-10)


def test_reference_rectification_identity(example_fit):
    """test/runtests.jl:79-85: rectification(c,1)(RowCol(1,2)) == c(RowCol(1,2),1)[[1,2]]"""
    rv, tv = example_fit["view_list"][0]
    ch = oc.chain(example_fit["intr_tuple"], rv, tv)
    x, y, z = oc.img2world_soa(ch, [1.0], [2.0])
    x2, y2, z2 = oc.img2world_soa(ch, [1.0], [2.0], want_z=False)
    assert z2 is None and x2[0] == x[0] and y2[0] == y[0]
    # SURVEY Appendix A known answer (numpy restatement at these parameters)
    assert abs(x[0] - (-0.709323329459491)) < 1e-12 and abs(y[0] - (-2.9327923607178543)) < 1e-12
    assert abs(z[0]) < 1e-12


def test_c_oracle_vs_numpy_twin_every_pixel(example_fit):
    """BASELINE config 1: round trip over every pixel of one 375x500 frame."""
    rv, tv = example_fit["view_list"][0]
    intr = example_fit["intr_tuple"]
    ch = oc.chain(intr, rv, tv)
    sz1, sz2 = example_fit["sz"]
    r, c = np.meshgrid(np.arange(1, sz1 + 1, dtype=np.float64),
                       np.arange(1, sz2 + 1, dtype=np.float64), indexing="ij")
    r, c = r.ravel(), c.ravel()
    x, y, z = oc.img2world_soa(ch, r, c)
    r2, c2 = oc.world2img_soa(ch, x, y, z)
    assert max(np.max(np.abs(r2 - r)), np.max(np.abs(c2 - c))) < 1e-9
    # numpy twin (eigen-solve per point) on a stride-37 subset
    sel = slice(None, None, 37)
    tw = on.Chain(intr, rv, tv)
    p = tw.img2world(np.stack([r[sel], c[sel]], axis=1))
    assert np.max(np.abs(p - np.stack([x[sel], y[sel], z[sel]], axis=1))) < 1e-11
    rc = tw.world2img(p)
    assert np.max(np.abs(rc - np.stack([r[sel], c[sel]], axis=1))) < 1e-9


def test_ratio_and_axes(example_fit):
    n1, n2 = example_fit["n_corners"]
    ip = example_fit["corners_np"][0].reshape(n2, n1, 2).transpose(1, 0, 2)  # [a, b]
    ratio = oc.get_ratio(ip, 1.0)
    assert abs(ratio - on.get_ratio(ip, 1.0)) < 1e-12
    assert abs(ratio - 37.49350720323292) < 1e-9
    assert oc.get_axes(ratio, 1.0, (n1, n2), example_fit["sz"]) == (-112, -119)
    assert on.get_axes(ratio, 1.0, (n1, n2), example_fit["sz"]) == (-112, -119)


def test_rectify_c_vs_numpy_twin(example_fit):
    rv, tv = example_fit["view_list"][0]
    intr = (55.0, 55.0, 30.0, 24.0, 0.06, 1.0)                # small camera for a 61x47 frame
    ch = oc.chain(intr, rv, tv)
    ratio = 48.0
    sz = (61, 47)
    axs = (-5, -9)
    rng = np.random.default_rng(3)
    img = rng.random((sz[1], sz[0])).astype(np.float32)       # memory order (c, r)
    out = oc.rectify_f32c1(ch, 1.0 / ratio * 6, axs, img[None], fill=np.nan)[0]
    ref, rc = on.rectify_gray(on.Chain(intr, rv, tv), 1.0 / ratio * 6, axs, img.T, np.nan)
    mr, mc = oc.rectify_map(ch, 1.0 / ratio * 6, axs, sz)
    assert np.max(np.abs(mr.T - rc[:, :, 0])) < 1e-10 and np.max(np.abs(mc.T - rc[:, :, 1])) < 1e-10
    got = out.T.astype(np.float64)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    assert 0.01 < np.mean(np.isnan(ref)) < 0.95               # both fill and samples exercised
    ok = ~np.isnan(ref)
    assert np.max(np.abs(got[ok] - ref[ok])) < 1e-6           # float32 store
    # u8 RGB path: same weights, round-half-even store
    img8 = rng.integers(0, 256, (1, sz[1], sz[0], 3), dtype=np.uint8)
    out8 = oc.rectify_u8c3(ch, 1.0 / ratio * 6, axs, img8, fill=(7, 8, 9))[0]
    for ch3 in range(3):
        ref8, _ = on.rectify_gray(on.Chain(intr, rv, tv), 1.0 / ratio * 6, axs,
                                  img8[0, :, :, ch3].T.astype(np.float64), -1.0)
        fillmask = ref8 < 0
        assert np.all(out8[:, :, ch3].T[fillmask] == (7, 8, 9)[ch3])
        d = np.abs(out8[:, :, ch3].T[~fillmask].astype(np.float64) - ref8[~fillmask])
        assert np.max(d) <= 0.5 + 1e-9


def test_rectify_edge_rule():
    """x == n is in bounds with (i, delta) = (n-1, 1); x just outside -> fill."""
    # identity-like camera: f=1, c=0, k=0, R=I, t=(0,0,1): row = I1*inv_ratio
    ch = oc.chain((1.0, 1.0, 0.0, 0.0, 0.0, 1.0), (0, 0, 0), (0, 0, 1.0))
    img = np.arange(12, dtype=np.float32).reshape(1, 3, 4)    # sz1=4 (contiguous), sz2=3
    out = oc.rectify_f32c1(ch, 1.0, (1, 1), img, fill=-1.0)[0]
    assert np.array_equal(out, img[0])                         # exact grid hits incl. x == n
    out = oc.rectify_f32c1(ch, 1.0, (0, 1), img, fill=-1.0)[0]
    assert np.all(out[:, 0] == -1.0) and np.array_equal(out[:, 1:], img[0][:, :3])
    out = oc.rectify_f32c1(ch, 0.5, (2, 2), img, fill=-1.0)[0]  # half-pixel steps
    assert out[0, 0] == img[0, 0, 0] and out[0, 1] == 0.5 * (img[0, 0, 0] + img[0, 0, 1])


# ---- independent stand-ins for the bilinear rule of ImageTransformations.warp -----------------------
# No reference test calls warp (with_plot=false everywhere) and Julia is absent, so the oracle's
# index / weight / fill rule is checked against two implementations nobody here wrote:
#   scipy.ndimage.map_coordinates(order=1, mode="constant"): floor index + double weights, no
#     interpolation beyond the edges -- the same OnGrid + fill rule as Interpolations' BSpline(Linear())
#     with extrapolation = fill (src/plot_calibration.jl:40);
#   cv2.remap(INTER_LINEAR, BORDER_CONSTANT): same taps, weights quantised to 1/32 px (INTER_BITS = 5),
#     so it pins the INDEX selection (an off-by-one tap would be an O(1) error), not the last digit.
def _warp_case():
    rv, tv = (0.21, -0.17, 0.35), (-3.4, -1.2, 7.0)
    intr = (55.0, 55.0, 30.0, 24.0, 0.06, 1.0)
    sz, axs, inv_ratio = (61, 47), (-5, -9), 1.0 / 8.0
    return rv, tv, intr, sz, axs, inv_ratio


def test_warp_rule_vs_scipy_map_coordinates():
    ndi = pytest.importorskip("scipy.ndimage")
    rv, tv, intr, sz, axs, inv_ratio = _warp_case()
    ch = oc.chain(intr, rv, tv)
    rng = np.random.default_rng(11)
    img = rng.random((sz[1], sz[0])).astype(np.float32)            # memory (c, r): img[c, r] = pixel (r+1, c+1)
    mr, mc = oc.rectify_map(ch, inv_ratio, axs, sz)                 # 1-based source coordinates, memory (c, r)
    out = oc.rectify_f32c1(ch, inv_ratio, axs, img[None], fill=np.nan)[0].astype(np.float64)
    ref = ndi.map_coordinates(img.astype(np.float64), [mc - 1.0, mr - 1.0], order=1, mode="constant", cval=np.nan)
    inb = (mr >= 1) & (mr <= sz[0]) & (mc >= 1) & (mc <= sz[1])
    assert 0.05 < inb.mean() < 0.95
    assert np.array_equal(np.isnan(out), ~inb) and np.array_equal(np.isnan(ref), ~inb)
    assert np.max(np.abs(out[inb] - ref[inb])) < 1e-7               # float32 store of the same double blend
    # u8 RGB: rint (half-even) of the same blend; scipy's double result rounded the same way
    img8 = rng.integers(0, 256, (1, sz[1], sz[0], 3), dtype=np.uint8)
    out8 = oc.rectify_u8c3(ch, inv_ratio, axs, img8, fill=(1, 2, 3))[0]
    for c3 in range(3):
        r8 = ndi.map_coordinates(img8[0, :, :, c3].astype(np.float64), [mc - 1.0, mr - 1.0], order=1,
                                 mode="constant", cval=-1.0)
        assert np.all(out8[:, :, c3][~inb] == (1, 2, 3)[c3])
        frac = np.abs(r8[inb] - np.floor(r8[inb]) - 0.5)
        sure = frac > 1e-9                                          # away from exact ties the rounding is unambiguous
        assert np.array_equal(out8[:, :, c3][inb][sure], np.rint(r8[inb][sure]).astype(np.uint8))
        assert sure.mean() > 0.99


def test_warp_rule_edges_vs_scipy():
    """x == 1, x == n, just outside and exact .5 positions (SURVEY 8c / VERDICT r1 #2a)."""
    ndi = pytest.importorskip("scipy.ndimage")
    ch = oc.chain((1.0, 1.0, 0.0, 0.0, 0.0, 1.0), (0, 0, 0), (0, 0, 1.0))   # row = I1 * inv_ratio exactly
    rng = np.random.default_rng(5)
    img = rng.random((9, 12)).astype(np.float32)                    # sz1 = 12, sz2 = 9
    sz = (12, 9)
    for inv_ratio, axs in ((1.0, (1, 1)), (1.0, (0, 0)), (0.5, (1, 1)), (0.5, (2, 2)), (0.25, (3, 5)), (1.0, (-3, 4))):
        mr, mc = oc.rectify_map(ch, inv_ratio, axs, sz)
        out = oc.rectify_f32c1(ch, inv_ratio, axs, img[None], fill=np.nan)[0].astype(np.float64)
        ref = ndi.map_coordinates(img.astype(np.float64), [mc - 1.0, mr - 1.0], order=1, mode="constant", cval=np.nan)
        assert np.array_equal(np.isnan(out), np.isnan(ref)), (inv_ratio, axs)
        ok = ~np.isnan(ref)
        assert ok.any() and np.max(np.abs(out[ok] - ref[ok])) < 1e-7, (inv_ratio, axs)
    # coordinates one ulp outside [1, n] are fill; exactly n is in bounds with the value of the last texel
    mr, mc = oc.rectify_map(ch, 1.0, (1, 1), sz)
    assert mr.max() == sz[0] and mc.max() == sz[1]
    out = oc.rectify_f32c1(ch, 1.0, (1, 1), img[None], fill=-7.0)[0]
    assert np.array_equal(out, img)


def test_warp_index_selection_vs_cv2_remap():
    cv2 = pytest.importorskip("cv2")
    rv, tv, intr, sz, axs, inv_ratio = _warp_case()
    ch = oc.chain(intr, rv, tv)
    rng = np.random.default_rng(12)
    img = rng.random((sz[1], sz[0])).astype(np.float32)
    mr, mc = oc.rectify_map(ch, inv_ratio, axs, sz)
    out = oc.rectify_f32c1(ch, inv_ratio, axs, img[None], fill=np.nan)[0]
    # cv2: dst(y, x) = src(map_y, map_x) with x the contiguous axis = our first RowCol axis
    ref = cv2.remap(img, (mr - 1.0).astype(np.float32), (mc - 1.0).astype(np.float32), cv2.INTER_LINEAR,
                    borderMode=cv2.BORDER_CONSTANT, borderValue=float("nan"))
    inner = (mr >= 2) & (mr <= sz[0] - 1) & (mc >= 2) & (mc <= sz[1] - 1)
    assert inner.mean() > 0.05
    err = np.abs(out[inner].astype(np.float64) - ref[inner])
    # weights quantised to 1/32 px on an image with |gradient| <= 1 per texel: <= 2/32 per axis
    assert err.max() < 4.0 / 32 + 1e-6
    assert err.mean() < 0.02                                        # an off-by-one tap would be ~0.3 on white noise


def test_operation_order_sensitivity(example_fit):
    """How much does the ORDER of the floating-point operations matter?  The CUDA kernels are bit-exact with the C
    oracle's order (matrix form, explicit fma).  The Julia libraries evaluate the same chain in another order
    (oracle/julia_order.py: Rodrigues on the vector, no fma).  On the reference's example views and on the bench view
    of BASELINE configs[1]: coordinates agree to 1e-10 px, NO bilinear tap index differs, no coordinate lies close
    enough to an integer for ulp-level differences to move a tap, and the fp32 result of the blend differs in at
    most a few pixels per million, by one ulp."""
    cases = [(example_fit["intr_tuple"], rv, tv, example_fit["sz"]) for rv, tv in example_fit["view_list"]]
    cases.append((C2_INTR, (0.05, -0.04, 0.02), (-9.3, -6.4, 30.0), (1080, 1920)))
    n1, n2 = example_fit["n_corners"]
    rng = np.random.default_rng(5)
    tot = diff_px = 0
    for ci, (intr, rv, tv, sz) in enumerate(cases):
        ch = oc.chain(intr, rv, tv)
        if ci < len(example_fit["view_list"]):
            ip = example_fit["corners_np"][ci].reshape(n2, n1, 2).transpose(1, 0, 2)
        else:
            a, b = np.meshgrid(np.arange(20, dtype=np.float64), np.arange(14, dtype=np.float64), indexing="ij")
            row, col = oc.world2img_soa(ch, a.ravel() * intr[5], b.ravel() * intr[5])
            ip = np.stack([row, col], axis=-1).reshape(20, 14, 2)
        ratio = oc.get_ratio(ip, intr[5])
        axs = oc.get_axes(ratio, intr[5], ip.shape[:2], sz)
        mr, mc = oc.rectify_map(ch, 1.0 / ratio, axs, sz)          # (sz2, sz1): first axis contiguous
        mj = oj.rectify_map(intr, rv, tv, 1.0 / ratio, axs, sz)    # (sz1, sz2, 2)
        jr, jc = mj[..., 0].T, mj[..., 1].T
        inb = (mr >= 1) & (mr <= sz[0]) & (mc >= 1) & (mc <= sz[1])
        assert inb.mean() > 0.2
        assert np.array_equal(inb, (jr >= 1) & (jr <= sz[0]) & (jc >= 1) & (jc <= sz[1])), ci
        assert max(np.max(np.abs(mr - jr)[inb]), np.max(np.abs(mc - jc)[inb])) < 1e-10, ci
        assert np.array_equal(np.floor(mr[inb]), np.floor(jr[inb])) and np.array_equal(np.floor(mc[inb]), np.floor(jc[inb])), ci
        # the only pixels an ulp-level difference could move: coordinates within 1e-9 of an integer
        near = (np.abs(mr - np.rint(mr)) < 1e-9) | (np.abs(mc - np.rint(mc)) < 1e-9)
        assert not np.any(near & inb), ci
        # fp32 blend with both sets of weights (same taps): a rounding boundary is crossed a few times per million
        img = rng.random((sz[1] + 1, sz[0] + 1)).astype(np.float32).astype(np.float64)
        def blend(r, c):
            f1, f2 = np.floor(r[inb]), np.floor(c[inb])
            f1, f2 = np.where(f1 > sz[0] - 1, f1 - 1, f1), np.where(f2 > sz[1] - 1, f2 - 1, f2)
            d1, d2 = r[inb] - f1, c[inb] - f2
            i1, i2 = f1.astype(np.int64) - 1, f2.astype(np.int64) - 1
            lo = d2 * img[i2 + 1, i1] + (1 - d2) * img[i2, i1]
            hi = d2 * img[i2 + 1, i1 + 1] + (1 - d2) * img[i2, i1 + 1]
            return (d1 * hi + (1 - d1) * lo).astype(np.float32)
        va, vb = blend(mr, mc), blend(jr, jc)
        if ci == 0:
            # RGB{N0f8}: the oracle blends the raw 0..255 values and rounds half-even; the reference blends
            # value/255 and re-quantises on store (FixedPointNumbers: round(255 x)).  Same real number; on a
            # random image no channel value lands close enough to a .5 tie for the two to differ
            b8 = rng.integers(0, 256, (sz[1] + 1, sz[0] + 1)).astype(np.float64)
            f1, f2 = np.floor(jr[inb]), np.floor(jc[inb])
            f1, f2 = np.where(f1 > sz[0] - 1, f1 - 1, f1), np.where(f2 > sz[1] - 1, f2 - 1, f2)
            d1, d2 = jr[inb] - f1, jc[inb] - f2
            i1, i2 = f1.astype(np.int64) - 1, f2.astype(np.int64) - 1
            def bl(im):
                return (1 - d1) * ((1 - d2) * im[i2, i1] + d2 * im[i2 + 1, i1]) + d1 * ((1 - d2) * im[i2, i1 + 1] + d2 * im[i2 + 1, i1 + 1])
            raw = np.rint(bl(b8))
            n0f8 = np.rint(255.0 * bl(b8 / 255.0))
            assert np.array_equal(raw, n0f8)
        ne = va != vb
        tot += va.size; diff_px += int(ne.sum())
        if ne.any():
            assert np.max(np.abs(va[ne].view(np.int32) - vb[ne].view(np.int32))) <= 1, ci
    assert diff_px / tot < 2e-5
