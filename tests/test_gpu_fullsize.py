"""BASELINE.json configurations at their FULL sizes, checked through size-independent properties
(the oracle cannot cover them in seconds): round trips, batch/chunk independence, linearity of the
normal-equation blocks, and bit-equality with the same computation at a size the oracle did cover
(tests/test_gpu_parity.py)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from conftest import BENCH_VIEW, C2_INTR, C3_INTR, SYN_VIEW
from oracle import oracle_c as oc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cc():
    import cameracalibrations_b200 as m
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    m.context(0)
    return m


def _calib(cc, intr, views):
    return cc.Calibration(intr[:4], views, 1.0 / intr[5], intr[4], [f"{i}.png" for i in range(len(views))])


def test_config4_100M_points_round_trip_fp64_and_fp32(cc):
    """C4: 100 M random RowCol -> world -> RowCol.  FP64 within 1e-9 px, FP32 within 1e-3 px
    (through the FP64 inverse), and the first 2^20 points are bit-identical to a separate call on
    just those points (the grid-stride kernels are position independent)."""
    c = _calib(cc, C3_INTR, [SYN_VIEW])
    n = 100_000_000
    g = torch.Generator(device="cuda").manual_seed(99)
    row = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * 2160
    col = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * 3840
    x, y, z = c.img2world(row, col, 0)
    r2, c2 = c.world2img(x, y, z, 0)
    assert torch.max(torch.abs(r2 - row)).item() <= 1e-9 and torch.max(torch.abs(c2 - col)).item() <= 1e-9
    assert torch.max(torch.abs(z)).item() <= 1e-9                     # the board plane
    m = 1 << 20
    xs, ys, zs = c.img2world(row[:m].clone(), col[:m].clone(), 0)
    assert torch.equal(xs, x[:m]) and torch.equal(ys, y[:m]) and torch.equal(zs, z[:m])
    ox, oy, oz = oc.img2world_soa(oc.chain(C3_INTR, *SYN_VIEW), row[:4096].cpu().numpy(), col[:4096].cpu().numpy())
    assert np.max(np.abs(x[:4096].cpu().numpy() - ox)) <= 1e-9 * max(1.0, np.abs(ox).max())
    del x, y, z, r2, c2
    # FP32 fast path: forward in FP32, back in FP64
    row32, col32 = row.to(torch.float32), col.to(torch.float32)
    x32, y32, z32 = c.img2world(row32, col32, 0)
    r3, c3 = c.world2img(x32.double(), y32.double(), z32.double(), 0)
    assert torch.max(torch.abs(r3 - row32.double())).item() <= 1e-3
    assert torch.max(torch.abs(c3 - col32.double())).item() <= 1e-3


def test_config5_10k_views_blocks_are_additive_and_view_local(cc):
    """C5: 10,000 views x 280 corners.  The shared block of the whole set is the sum of the shared
    blocks of its halves (what the NCCL all-reduce relies on); a view's 66-double block does not
    depend on which batch it is evaluated in (bit-equal); sum r^2 equals the per-view residuals."""
    rng = np.random.default_rng(7)
    nv, n1, n2 = 10_000, 20, 14
    obj = np.array([[a, b, 0.0] for b in range(n2) for a in range(n1)], dtype=np.float64)
    views = np.concatenate([rng.normal(0, 0.3, (nv, 3)), np.array([-10.0, -7.0, 40.0]) + rng.normal(0, 2.0, (nv, 3))], 1)
    tv, to = torch.from_numpy(views).cuda(), torch.from_numpy(obj).cuda()
    c = _calib(cc, C3_INTR, [(v[:3], v[3:]) for v in views[:4]])
    # image points = forward projection (device) + noise
    img = torch.empty((nv, n1 * n2, 2), dtype=torch.float64, device="cuda")
    pv0, sh0 = cc.reproj_jtj(C3_INTR, 1.0, tv, to, torch.zeros_like(img))      # residual = projection - 0
    for i in range(4):                                                         # spot-check the projection
        r, q = c.world2img(to[:, 0].contiguous(), to[:, 1].contiguous(), to[:, 2].contiguous(), i)
        orow, ocol = oc.world2img(oc.chain(C3_INTR, views[i, :3], views[i, 3:]), obj)
        assert np.max(np.abs(r.cpu().numpy() - orow)) <= 1e-9 and np.max(np.abs(q.cpu().numpy() - ocol)) <= 1e-9
        img[i, :, 0], img[i, :, 1] = r, q
    g = torch.Generator(device="cuda").manual_seed(5)
    img = torch.rand((nv, n1 * n2, 2), dtype=torch.float64, device="cuda", generator=g) * 2000
    pv, sh = cc.reproj_jtj(C3_INTR, 1.0, tv, to, img)
    h = nv // 2
    pv_a, sh_a = cc.reproj_jtj(C3_INTR, 1.0, tv[:h].contiguous(), to, img[:h].contiguous())
    pv_b, sh_b = cc.reproj_jtj(C3_INTR, 1.0, tv[h:].contiguous(), to, img[h:].contiguous())
    assert torch.equal(pv[:h], pv_a) and torch.equal(pv[h:], pv_b)
    assert torch.allclose(sh, sh_a + sh_b, rtol=1e-12, atol=0.0)
    # oracle on a handful of the 10k views
    idx = [0, 1, 4999, 5000, 9999]
    pv_o, _, _ = oc.reproj_jtj(C3_INTR, 1.0, [(views[i, :3], views[i, 3:]) for i in idx], obj, img[idx].cpu().numpy())
    assert np.max(np.abs(pv[idx].cpu().numpy() - pv_o) / (1.0 + np.abs(pv_o))) <= 1e-9


def test_config2_full_batch_every_frame_like_the_first(cc):
    """C2: 64 x 1080x1920 fp32 frames in one call: equal inputs give bit-equal outputs in every
    frame slot, both coordinate pipelines (frame 0 itself is checked against the oracle in
    test_gpu_parity.py::test_rectify_batch_frames_independent)."""
    sz = (1080, 1920)
    import bench
    wl = bench.WORKLOADS["c2"]
    ip = bench.geometry(wl)
    ratio = cc.get_ratio(ip, 1.0)
    axs = cc.get_axes(ratio, 1.0, bench.N_CORNERS, sz)
    c = _calib(cc, C2_INTR, [BENCH_VIEW])
    g = torch.Generator(device="cuda").manual_seed(1234)
    one = torch.rand((1, sz[1], sz[0]), device="cuda", generator=g)
    batch = one.repeat(64, 1, 1).contiguous()
    for coord in ("f64", "f32"):
        out = torch.nan_to_num(cc.warp(c, 0, batch, ratio, axs, coord=coord), nan=-1.0)
        first = out[0:1].expand_as(out)
        assert torch.equal(out, first), coord
        single = torch.nan_to_num(cc.warp(c, 0, one, ratio, axs, coord=coord), nan=-1.0)
        assert torch.equal(single[0], out[17]), coord


def test_config3_stream_ring_groups(cc):
    """C3: the 4096-frame 4K u8 RGB stream does not fit in HBM; it is processed in ring groups
    (shard_frames).  40 frames through groups of 16: every group is rectified by the same plan and
    a frame's output does not depend on its group (bit-equal to the frame rectified alone)."""
    import bench
    sz = (2160, 3840)
    wl = bench.WORKLOADS["c3"]
    ratio = cc.get_ratio(bench.geometry(wl), 1.0)
    axs = cc.get_axes(ratio, 1.0, bench.N_CORNERS, sz)
    c = _calib(cc, C3_INTR, [BENCH_VIEW])
    g = torch.Generator(device="cuda").manual_seed(4321)
    base = torch.randint(0, 256, (3, sz[1], sz[0], 3), dtype=torch.uint8, device="cuda", generator=g)
    alone = [cc.warp(c, 0, base[i:i + 1].contiguous(), ratio, axs) for i in range(3)]
    groups = cc.shard_frames(40, 0, 1, ring=16)
    assert groups == [(0, 16), (16, 32), (32, 40)]
    ring = torch.empty((16, sz[1], sz[0], 3), dtype=torch.uint8, device="cuda")
    out = torch.empty_like(ring)
    for lo, hi in groups:
        for f in range(lo, hi):
            ring[f - lo] = base[f % 3]
        cc.warp(c, 0, ring[:hi - lo], ratio, axs, out=out[:hi - lo])
        for f in range(lo, hi):
            assert torch.equal(out[f - lo], alone[f % 3][0]), f
