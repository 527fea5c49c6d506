"""Cold call of a single view (tile plan built on the host, uploaded, first launch): wall clock per frame size."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bench
import cameracalibrations_b200 as cc

BV = bench.BENCH_VIEW
for wname in ("c2", "c3"):
    wl = bench.WORKLOADS[wname]
    sz = wl["sz"]
    ratio = cc.get_ratio(bench.geometry(wl), 1.0)
    axs = cc.get_axes(ratio, 1.0, bench.N_CORNERS, sz)
    src = (torch.randint(0, 256, (1, sz[1], sz[0], 3), dtype=torch.uint8, device="cuda") if wl["u8"]
           else torch.rand((1, sz[1], sz[0]), dtype=torch.float32, device="cuda"))
    dst = torch.empty_like(src)
    ts = []
    for i in range(6):                               # six views nobody has planned yet
        v = ((BV[0][0] + 0.01 * (i + 1), BV[0][1], BV[0][2]), BV[1])
        cal = cc.Calibration(wl["intr"][:4], [v], 1.0, wl["intr"][4], ["x.png"])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        cc.warp(cal, 0, src, ratio, axs, coord="f32", out=dst)
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    print(f"{wname} {sz}: cold call of a new view, ms: " + " ".join(f"{t:.2f}" for t in ts), flush=True)
