"""Launch each hot kernel a few times on its BASELINE-size input (for ncu / launch lists).
    python profiles/prof_kernels.py [rectify|points|jtj|all] [--coord f64|f32|both] [--gather auto|direct|tma]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
import numpy as np
import torch
import cameracalibrations_b200 as cc
import bench

ap = argparse.ArgumentParser()
ap.add_argument("what", nargs="?", default="all")
ap.add_argument("--coord", default="both")
ap.add_argument("--gather", default="auto")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda", 0)
coords = ("f64", "f32") if a.coord == "both" else (a.coord,)
if a.what in ("rectify", "all", "c2", "c3"):
    for wname in ("c2", "c3"):
        if a.what in ("c2", "c3") and a.what != wname:
            continue
        wl = bench.WORKLOADS[wname]
        sz, nfr = wl["sz"], wl["frames"]
        cal = cc.Calibration(wl["intr"][:4], [bench.BENCH_VIEW], 1.0, wl["intr"][4], ["extrinsic.png"])
        ratio = cc.get_ratio(bench.geometry(wl), 1.0)
        axs = cc.get_axes(ratio, 1.0, bench.N_CORNERS, sz)
        if wl["u8"]:
            src = torch.randint(0, 256, (nfr, sz[1], sz[0], 3), dtype=torch.uint8, device=dev)
        else:
            src = torch.rand((nfr, sz[1], sz[0]), dtype=torch.float32, device=dev)
        dst = torch.empty_like(src)
        for coord in coords:
            for _ in range(a.reps):
                cc.warp(cal, 0, src, ratio, axs, coord=coord, gather=a.gather, out=dst)
        torch.cuda.synchronize()
        del src, dst
if a.what in ("views", "all"):
    # frames with different views in one launch (rectify_*_views_kernel): 64 x 1080p fp32, 16 x 4K u8 RGB
    BV = bench.BENCH_VIEW
    vlist = [((BV[0][0] + 0.002 * i, BV[0][1], BV[0][2] + 0.001 * i), (BV[1][0] + 0.01 * i, BV[1][1], BV[1][2] + 0.05 * i)) for i in range(64)]
    for wname, nv in (("c2", 64), ("c3", 16)):
        wl = bench.WORKLOADS[wname]
        sz = wl["sz"]
        cal = cc.Calibration(wl["intr"][:4], vlist[:nv], 1.0, wl["intr"][4], [f"{i}.png" for i in range(nv)])
        ratio = cc.get_ratio(bench.geometry(wl), 1.0)
        axs = cc.get_axes(ratio, 1.0, bench.N_CORNERS, sz)
        if wl["u8"]:
            src = torch.randint(0, 256, (nv, sz[1], sz[0], 3), dtype=torch.uint8, device=dev)
        else:
            src = torch.rand((nv, sz[1], sz[0]), dtype=torch.float32, device=dev)
        dst = torch.empty_like(src)
        for coord in coords:
            for _ in range(a.reps):
                cc.warp_views(cal, list(range(nv)), src, [ratio] * nv, [axs] * nv, coord=coord, out=dst)
        torch.cuda.synchronize()
        del src, dst
if a.what in ("points", "all"):
    wl = bench.WORKLOADS["c3"]
    cal = cc.Calibration(wl["intr"][:4], [bench.BENCH_VIEW], 1.0, wl["intr"][4], ["extrinsic.png"])
    n = 100_000_000
    for dt in (torch.float64, torch.float32):
        row = torch.rand(n, dtype=dt, device=dev) * 2160
        col = torch.rand(n, dtype=dt, device=dev) * 3840
        for _ in range(a.reps):
            x, y, z = cal.img2world(row, col, 0)
            r, c = cal.world2img(x, y, z, 0)
        torch.cuda.synchronize()
        del row, col, x, y, z, r, c
if a.what in ("jtj", "all"):
    wl = bench.WORKLOADS["c3"]
    rng = np.random.default_rng(7)
    nv, nc = 10_000, 280
    views = np.concatenate([rng.normal(0, 0.3, (nv, 3)), np.array([-10.0, -7.0, 40.0]) + rng.normal(0, 2.0, (nv, 3))], 1)
    obj = np.array([[i, j, 0.0] for j in range(14) for i in range(20)], dtype=np.float64)
    tv, to = torch.from_numpy(views).to(dev), torch.from_numpy(obj).to(dev)
    ti = torch.rand((nv, nc, 2), dtype=torch.float64, device=dev) * 2000
    for _ in range(a.reps):
        cc.reproj_jtj(wl["intr"], 1.0, tv, to, ti)
    torch.cuda.synchronize()
print("ok")
