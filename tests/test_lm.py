"""Levenberg-Marquardt fit (cameracalibrations_b200/lm.py + csrc/lm.cu): the solve the reference
delegates to OpenCV.calibrateCamera (src/detect_fit.jl:47).  Checked against
  * a dense numpy solve of the same damped normal equations built from the ORACLE's blocks,
  * the cv2.calibrateCamera result stored in tests/golden/example_fit.json (same flags as
    src/detect_fit.jl:40; cv2 stops at the reference's CRITERIA, the LM here may go further,
    so its RMS must be <= cv2's).
"""
import numpy as np
import pytest

from oracle import oracle_c as oc


def _dense_step(pv, sh, lam, views, mask=0b1111):
    """(A + lam diag A) delta = -g on the full arrowhead matrix, numpy"""
    nv = len(views)
    n = 6 * nv + 4
    H, g = np.zeros((n, n)), np.zeros(n)
    for v in range(nv):
        s = slice(6 * v, 6 * v + 6)
        H[s, s] = pv[v, :36].reshape(6, 6)
        H[s, 6 * nv:] = pv[v, 36:60].reshape(6, 4)
        H[6 * nv:, s] = pv[v, 36:60].reshape(6, 4).T
        g[s] = pv[v, 60:66]
    H[6 * nv:, 6 * nv:] = sh[:16].reshape(4, 4)
    g[6 * nv:] = sh[16:20]
    H[np.diag_indices(n)] *= 1.0 + lam
    for a in range(4):
        if not (mask >> a) & 1:
            j = 6 * nv + a
            H[j, :] = 0; H[:, j] = 0; H[j, j] = 1; g[j] = 0
    d = np.linalg.solve(H, -g)
    return d[6 * nv:], np.asarray(views) + d[:6 * nv].reshape(nv, 6)


# ------------------------------------------------------------------ host logic (CPU)
def test_initial_guess_close_to_cv2(example_fit):
    from cameracalibrations_b200 import lm
    intr0, views0 = lm.initial_guess(example_fit["obj_np"], example_fit["corners_np"], example_fit["sz"], 1.0)
    cv = example_fit["intr_tuple"]
    assert abs(intr0[0] - cv[0]) / cv[0] < 0.02 and intr0[0] == intr0[1]
    assert intr0[2:4] == ((example_fit["sz"][0] - 1) / 2.0, (example_fit["sz"][1] - 1) / 2.0) and intr0[4] == 0.0
    for v0, (rv, tv) in zip(views0, example_fit["view_list"]):
        assert np.max(np.abs(v0[:3] - rv)) < 0.05 and np.max(np.abs(v0[3:] - tv)) < 0.25
    # reprojection error of the starting point, through the oracle: a usable start (a few px)
    _, sh, _ = oc.reproj_jtj(tuple(intr0) + (1.0,), 1.0, [(v[:3], v[3:]) for v in views0],
                             example_fit["obj_np"], example_fit["corners_np"])
    assert np.sqrt(sh[20] / example_fit["corners_np"][..., 0].size) < 3.0


def test_rodrigues_inverse_round_trip():
    from cameracalibrations_b200 import lm
    from oracle import oracle_np as on
    rng = np.random.default_rng(5)
    for rv in list(rng.normal(0, 1.0, (50, 3))) + [np.zeros(3), np.array([np.pi - 1e-9, 0, 0]), np.array([1e-14, 0, 0])]:
        back = lm._rodrigues_inv(on.rodrigues(rv))
        assert np.allclose(on.rodrigues(back), on.rodrigues(rv), atol=1e-9)


# ------------------------------------------------------------------ device (GPU)
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def cc():
    import cameracalibrations_b200 as m
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    m.context(0)
    return m


def _c5(nviews, seed=7):
    rng = np.random.default_rng(seed)
    n1, n2 = 20, 14
    intr = (2800.0, 2800.0, 1080.0, 1920.0, -0.12, 1.0)
    obj = np.array([[a, b, 0.0] for b in range(n2) for a in range(n1)], dtype=np.float64)
    rv = rng.normal(0, 0.3, (nviews, 3))
    tv = np.array([-10.0, -7.0, 40.0]) + rng.normal(0, 2.0, (nviews, 3))
    img = np.empty((nviews, n1 * n2, 2))
    for i in range(nviews):
        img[i, :, 0], img[i, :, 1] = oc.world2img(oc.chain(intr, rv[i], tv[i]), obj)
    return intr, np.concatenate([rv, tv], 1), obj, img, rng


@pytest.mark.gpu
@pytest.mark.parametrize("lam,mask", [(1e-3, 0b1111), (10.0, 0b1111), (0.0, 0b0111), (1e-6, 0b0111)])
def test_lm_step_matches_dense_solve(cc, lam, mask):
    from cameracalibrations_b200 import lm
    intr, views, obj, img, rng = _c5(37)
    img = img + rng.normal(0, 0.25, img.shape)
    start = views + rng.normal(0, 1e-3, views.shape)                # off the optimum: non-zero gradient
    pv_o, sh_o, _ = oc.reproj_jtj(intr, 1.0, [(v[:3], v[3:]) for v in start], obj, img)
    di_ref, cand_ref = _dense_step(pv_o, sh_o, lam, start, mask)
    dv = torch.device("cuda")
    pv, sh = cc.reproj_jtj(intr, 1.0, torch.from_numpy(start).to(dv), torch.from_numpy(obj).to(dv), torch.from_numpy(img).to(dv))
    yz, schur = lm.lm_schur(pv, lam)
    cand, delta = lm.lm_update(sh, schur, lam, mask, yz, torch.from_numpy(start).to(dv))
    d = delta.cpu().numpy()
    assert schur[20].item() == 0.0 and d[6] == 1.0
    scale = np.abs(di_ref).max()
    assert np.max(np.abs(d[:4] - di_ref)) <= 1e-7 * scale + 1e-12
    if not mask & 0b1000:
        assert d[3] == 0.0
    assert np.max(np.abs(cand.cpu().numpy() - cand_ref)) <= 1e-7 * np.abs(cand_ref - start).max() + 1e-12
    assert np.isclose(d[4], np.sum((cand_ref - start) ** 2), rtol=1e-6)
    assert np.isclose(d[5], np.sum(start ** 2), rtol=1e-12)


@pytest.mark.gpu
def test_lm_fit_example_reaches_cv2_optimum(cc, example_fit):
    """From the homography start to (at least) cv2.calibrateCamera's optimum on test/example."""
    from cameracalibrations_b200 import lm
    obj, imgs = example_fit["obj_np"], example_fit["corners_np"]
    intr0, views0 = lm.initial_guess(obj, imgs, example_fit["sz"], 1.0)
    hist = []
    r = lm.lm_fit(intr0, views0, obj, imgs, history=hist)            # CRITERIA defaults (30, 1e-3)
    assert r["iterations"] <= 30 and r["accepted"] >= 2
    assert r["rms"] <= example_fit["cv2_rms"] + 1e-6
    tight = lm.lm_fit(intr0, views0, obj, imgs, max_iter=60, eps=1e-10)
    assert tight["rms"] <= example_fit["cv2_rms"] and tight["rms"] <= r["rms"] + 1e-12
    cv = example_fit["intr_tuple"]
    assert abs(tight["intr"][0] - cv[0]) / cv[0] < 1e-4 and tight["intr"][0] == tight["intr"][1]
    assert abs(tight["intr"][2] - cv[2]) < 0.05 and abs(tight["intr"][3] - cv[3]) < 0.05
    assert abs(tight["intr"][4] - cv[4]) < 1e-3
    for v, (rv, tv) in zip(tight["views"], example_fit["view_list"]):
        assert np.max(np.abs(v[:3] - rv)) < 1e-3 and np.max(np.abs(v[3:] - tv)) < 5e-3
    # at the optimum the gradient vanishes (oracle's blocks at the fitted parameters)
    pv, sh, _ = oc.reproj_jtj(tuple(tight["intr"]) + (1.0,), 1.0, [(v[:3], v[3:]) for v in tight["views"]], obj, imgs)
    pv0, sh0, _ = oc.reproj_jtj(tuple(intr0) + (1.0,), 1.0, [(v[:3], v[3:]) for v in views0], obj, imgs)
    assert np.abs(sh[16:20]).max() <= 1e-6 * np.abs(sh0[16:20]).max()
    assert np.abs(pv[:, 60:]).max() <= 1e-6 * np.abs(pv0[:, 60:]).max()
    # the fitted object is the reference's calibration object: errors below the reference's bounds
    c = cc.Calibration.from_fit([v[:3] for v in tight["views"]], [v[3:] for v in tight["views"]], *tight["intr"][:4],
                                1.0, tight["intr"][4], example_fit["files"])
    eps = cc.calculate_errors(c, imgs, obj, 1.0, example_fit["sz"], example_fit["files"], example_fit["n_corners"],
                              rng=np.random.default_rng(1))
    assert all(eps[k] < 1.0 for k in ("reprojection", "projection", "distance", "inverse"))   # test/runtests.jl:76


@pytest.mark.gpu
def test_lm_fit_host_entry_point_is_the_same_fit(cc, example_fit):
    """cc_lm_fit_f64_host (one C-ABI call, host arrays) runs the same loop as lm.lm_fit."""
    from cameracalibrations_b200 import lm
    obj, imgs = example_fit["obj_np"], example_fit["corners_np"]
    intr0, views0 = lm.initial_guess(obj, imgs, example_fit["sz"], 1.0)
    a = lm.lm_fit(intr0, views0, obj, imgs)
    b = lm.lm_fit_host(intr0, views0, obj, imgs)
    assert a["iterations"] == b["iterations"]
    assert np.array_equal(np.asarray(a["intr"]), np.asarray(b["intr"])) and np.array_equal(a["views"], b["views"])
    assert a["rms"] == b["rms"] and b["rms"] <= example_fit["cv2_rms"] + 1e-6
    # bad arguments are status codes
    with pytest.raises(cc.CamcalError):
        lm.lm_fit_host(intr0, views0, obj, imgs, max_iter=-1)


@pytest.mark.gpu
def test_fit_model_device_solver_vs_opencv(cc, example_fit):
    """fit_model (src/detect_fit.jl:27-61) with solver="b200" against solver="opencv" (cv2 here,
    OpenCV.jl in the reference) on the golden corners of test/example: same model, same optimum."""
    pytest.importorskip("cv2")
    obj, imgs, sz = example_fit["obj_np"], example_fit["corners_np"], tuple(example_fit["sz"])
    a = cc.fit_model(sz, obj, imgs, example_fit["n_corners"], solver="b200")
    b = cc.fit_model(sz, obj, imgs, example_fit["n_corners"], solver="opencv")
    assert a["rms"] <= b["rms"] + 1e-6                     # cv2 works on float32 copies of the corners
    assert abs(a["frow"] - b["frow"]) / b["frow"] < 1e-3 and a["frow"] == a["fcol"]
    assert abs(a["crow"] - b["crow"]) < 0.1 and abs(a["ccol"] - b["ccol"]) < 0.1 and abs(a["k"] - b["k"]) < 2e-3
    for ra, rb, ta, tb in zip(a["Rs"], b["Rs"], a["ts"], b["ts"]):
        assert np.max(np.abs(ra - rb)) < 2e-3 and np.max(np.abs(ta - tb)) < 1e-2
    # without distortion too (CALIB_FIX_K1)
    a0 = cc.fit_model(sz, obj, imgs, example_fit["n_corners"], with_distortion=False, solver="b200")
    b0 = cc.fit_model(sz, obj, imgs, example_fit["n_corners"], with_distortion=False, solver="opencv")
    assert a0["k"] == 0.0 and b0["k"] == 0.0 and a0["rms"] <= b0["rms"] + 1e-6


@pytest.mark.gpu
def test_lm_fit_without_distortion_and_many_views(cc):
    """CALIB_FIX_K1 (with_distortion == false) keeps k at 0; 500 noisy synthetic views converge to the
    generating camera."""
    from cameracalibrations_b200 import lm
    intr, views, obj, img, rng = _c5(500, seed=11)
    intr = intr[:4] + (0.0, 1.0)
    for i in range(len(views)):
        img[i, :, 0], img[i, :, 1] = oc.world2img(oc.chain(intr, views[i, :3], views[i, 3:]), obj)
    noisy = img + rng.normal(0, 0.1, img.shape)
    intr0, views0 = lm.initial_guess(obj, noisy, (2160, 3840), 1.0)
    r = lm.lm_fit(intr0, views0, obj, noisy, with_distortion=False, max_iter=40, eps=1e-9)
    assert r["intr"][4] == 0.0
    assert abs(r["rms"] - 0.1 * np.sqrt(2)) < 0.005          # noise floor: sqrt(2) sigma per point
    assert abs(r["intr"][0] - 2800.0) < 1.0 and abs(r["intr"][2] - 1080.0) < 1.0 and abs(r["intr"][3] - 1920.0) < 1.0
    assert np.max(np.abs(r["views"] - views)) < 0.05


@pytest.mark.gpu
def test_lm_fit_device_loop_is_the_stepwise_loop(cc, example_fit):
    """cc_lm_fit_f64 (device-resident state: damping, accept/reject and stopping rule decided by a
    kernel, no host synchronisation per iteration) takes exactly the decisions of the stepwise
    loop lm.lm_fit drives from the host: same iterations, same parameters bit for bit."""
    from cameracalibrations_b200 import lm
    obj, imgs = example_fit["obj_np"], example_fit["corners_np"]
    intr0, views0 = lm.initial_guess(obj, imgs, example_fit["sz"], 1.0)
    for kw in (dict(), dict(max_iter=60, eps=1e-10), dict(with_distortion=False), dict(max_iter=3), dict(max_iter=0)):
        a = lm.lm_fit(intr0, views0, obj, imgs, **kw)
        b = lm.lm_fit_device(intr0, views0, obj, imgs, **kw)
        assert a["iterations"] == b["iterations"], kw
        assert np.array_equal(np.asarray(a["intr"]), np.asarray(b["intr"])), kw
        assert np.array_equal(a["views"], b["views"]) and a["rms"] == b["rms"], kw
    # device tensors in, and a second call reuses the context's workspace
    vt = torch.from_numpy(np.asarray(views0)).cuda()
    c = lm.lm_fit_device(intr0, vt, torch.from_numpy(obj).cuda(), torch.from_numpy(imgs).cuda())
    assert c["rms"] <= example_fit["cv2_rms"] + 1e-6 and c["views_device"].is_cuda
    # a world of one: the collective of the ABI is a no-op and costs nothing
    ctx = cc.context(0)
    assert ctx.comm_size() == (1, 0)
    t = torch.arange(21, dtype=torch.float64, device="cuda")
    n0 = ctx.collective_count()
    assert torch.equal(ctx.allreduce(t.clone()), t) and ctx.collective_count() == n0


@pytest.mark.gpu
def test_initial_guess_on_the_device(cc, example_fit):
    """cc_lm_initial_guess_f64 (batched DLT + pose kernels) against the numpy starting values, and as
    the start of the fit: same optimum."""
    from cameracalibrations_b200 import lm
    obj, imgs = example_fit["obj_np"], example_fit["corners_np"]
    ih, vh = lm.initial_guess(obj, imgs, example_fit["sz"], 1.0)
    idv, vd = lm.initial_guess_device(obj, imgs, example_fit["sz"], 1.0)
    assert np.allclose(idv, ih, rtol=1e-7, atol=1e-12)
    assert np.allclose(vd.cpu().numpy(), vh, rtol=1e-6, atol=1e-7)
    a = lm.lm_fit_device(ih, vh, obj, imgs, max_iter=60, eps=1e-10)
    b = lm.lm_fit_device(idv, vd, obj, imgs, max_iter=60, eps=1e-10)
    assert abs(a["rms"] - b["rms"]) < 1e-9 and np.allclose(a["intr"], b["intr"], rtol=1e-7)
    # many synthetic views, steeper tilts, a rectangular aspect
    intr, views, obj5, img5, rng = _c5(300, seed=5)
    noisy = img5 + rng.normal(0, 0.1, img5.shape)
    ih, vh = lm.initial_guess(obj5, noisy, (2160, 3840), 1.0)
    idv, vd = lm.initial_guess_device(obj5, noisy, (2160, 3840), 1.0)
    assert np.allclose(idv, ih, rtol=1e-6)
    assert np.allclose(vd.cpu().numpy(), vh, rtol=1e-5, atol=1e-6)
    with pytest.raises(cc.CamcalError):
        lm.initial_guess_device(obj5[:3], noisy[:, :3], (2160, 3840), 1.0)     # a homography needs 4 points


def _nccl_worker(rank, world, port, out):
    import os
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from cameracalibrations_b200 import lm
    from cameracalibrations_b200.shard import shard_range
    intr, views, obj, img, rng = _c5(64, seed=3)
    noisy = img + rng.normal(0, 0.2, img.shape)
    intr0, views0 = lm.initial_guess(obj, noisy, (2160, 3840), 1.0)
    lo, hi = shard_range(len(views0), rank, world)
    r = lm.lm_fit(intr0, views0[lo:hi], obj, noisy[lo:hi], max_iter=40, eps=1e-9, device=rank)
    # the device-resident loop on raw NCCL (cc_comm_init_rank + cc_lm_fit_f64), start values from
    # the device too (the focal-length equations are all-reduced over the ranks)
    import cameracalibrations_b200 as cc
    ctx = cc.context(rank)
    n_coll = ctx.collective_count()
    id0, vd0 = lm.initial_guess_device(obj, noisy[lo:hi], (2160, 3840), 1.0, device=rank)
    rd = lm.lm_fit_device(intr0, views0[lo:hi], obj, noisy[lo:hi], max_iter=40, eps=1e-9, device=rank)
    rd.pop("views_device")
    rd["collectives"] = ctx.collective_count() - n_coll
    rd["comm"] = ctx.comm_size()
    rd["init_intr"] = id0
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, r, rd))
    if rank == 0:
        out.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_lm_fit_views_sharded_over_two_gpus(cc):
    """Views sharded over 2 ranks (NCCL all-reduce of the shared blocks): same fit as one GPU."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import socket
    import torch.multiprocessing as mp
    from cameracalibrations_b200 import lm
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    gathered = out.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    intr, views, obj, img, rng = _c5(64, seed=3)
    noisy = img + rng.normal(0, 0.2, img.shape)
    intr0, views0 = lm.initial_guess(obj, noisy, (2160, 3840), 1.0)
    one = lm.lm_fit(intr0, views0, obj, noisy, max_iter=40, eps=1e-9)
    r0, r1 = gathered[0][2], gathered[1][2]
    # both ranks take the same decisions; the summation order differs from the one-GPU run, so
    # near the 1e-9 floor the number of accepted / rejected steps may differ by a few
    assert r0["intr"] == r1["intr"] and r0["iterations"] == r1["iterations"]
    np.testing.assert_allclose(r0["intr"], one["intr"], rtol=1e-9)
    np.testing.assert_allclose(np.concatenate([r0["views"], r1["views"]]), one["views"], rtol=1e-7, atol=1e-9)
    assert abs(r0["rms"] - one["rms"]) < 1e-9
    # cc_lm_fit_f64 over the communicator of the C ABI: the same decisions as the stepwise loop
    d0, d1 = gathered[0][3], gathered[1][3]
    assert d0["comm"] == (2, 0) and d1["comm"] == (2, 1)
    assert d0["intr"] == d1["intr"] == r0["intr"] and d0["iterations"] == d1["iterations"] == r0["iterations"]
    assert np.array_equal(d0["views"], r0["views"]) and np.array_equal(d1["views"], r1["views"])
    assert d0["rms"] == r0["rms"]
    assert d0["collectives"] == 2 + 2 * d0["iterations"] or d0["collectives"] >= 2 + 2 * d0["iterations"]   # init + two per iteration
    assert np.allclose(d0["init_intr"], intr0, rtol=1e-6) and d0["init_intr"] == d1["init_intr"]


def _group_worker(out):
    """One PROCESS, two GPUs: cc_ctx_create_group + cc_reproj_jtj_f64 per device + one grouped all-reduce."""
    import ctypes as C
    import numpy as np
    import torch
    from cameracalibrations_b200 import _lib
    lib = _lib.lib
    devs = (C.c_int * 2)(0, 1)
    ctxs = (C.c_void_p * 2)()
    _lib.check(lib.cc_ctx_create_group(2, devs, ctxs))
    nr, rk = C.c_int(), C.c_int()
    sizes = []
    for i in range(2):
        _lib.check(lib.cc_comm_size(ctxs[i], C.byref(nr), C.byref(rk)))
        sizes.append((nr.value, rk.value))
    rng = np.random.default_rng(5)
    nv, nc = 40, 280
    views = np.concatenate([rng.normal(0, 0.3, (nv, 3)), np.array([-10.0, -7.0, 40.0]) + rng.normal(0, 2.0, (nv, 3))], 1)
    obj = np.array([[a, b, 0.0] for b in range(14) for a in range(20)], dtype=np.float64)
    img = rng.normal(1000.0, 300.0, (nv, nc, 2))
    intr = _lib.make_intr(2800.0, 2800.0, 1080.0, 1920.0, -0.12, 1.0)
    def run(ctx, dev, v, im):
        with torch.cuda.device(dev):
            tv = torch.from_numpy(v).to(f"cuda:{dev}"); to = torch.from_numpy(obj).to(f"cuda:{dev}"); ti = torch.from_numpy(im).to(f"cuda:{dev}")
            pv = torch.empty((len(v), 66), dtype=torch.float64, device=f"cuda:{dev}")
            sh = torch.empty(21, dtype=torch.float64, device=f"cuda:{dev}")
            _lib.check(lib.cc_reproj_jtj_f64(ctx, C.byref(intr), 1.0, C.c_void_p(tv.data_ptr()), len(v), C.c_void_p(to.data_ptr()),
                                             C.c_void_p(ti.data_ptr()), nc, C.c_void_p(pv.data_ptr()), C.c_void_p(sh.data_ptr()),
                                             C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
            return sh, (tv, to, ti, pv)
    whole, keep0 = run(ctxs[0], 0, views, img)
    torch.cuda.synchronize(0)
    whole = whole.cpu().numpy().copy()
    a, keep1 = run(ctxs[0], 0, views[:25], img[:25])
    b, keep2 = run(ctxs[1], 1, views[25:], img[25:])
    bufs = (C.c_void_p * 2)(a.data_ptr(), b.data_ptr())
    streams = (C.c_void_p * 2)(torch.cuda.current_stream(0).cuda_stream, torch.cuda.current_stream(1).cuda_stream)
    _lib.check(lib.cc_allreduce_shared_group(ctxs, 2, bufs, 21, streams))
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    out.put((sizes, whole, a.cpu().numpy(), b.cpu().numpy()))
    _lib.check(lib.cc_ctx_destroy_group(2, ctxs))


@pytest.mark.gpu
def test_single_process_group_of_two_gpus(cc):
    """SURVEY 8b's multi-device context: one process, cc_ctx_create_group over two devices, the 21-double block of
    each device's views summed by one grouped NCCL call == the block of all views on one device."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    p = ctx.Process(target=_group_worker, args=(out,))     # own process: NCCL state does not leak into the suite
    p.start()
    sizes, whole, a, b = out.get(timeout=300)
    p.join(timeout=120)
    assert p.exitcode == 0
    assert sizes == [(2, 0), (2, 1)]
    assert np.array_equal(a, b)                             # every rank holds the same sum
    np.testing.assert_allclose(a, whole, rtol=1e-12)
