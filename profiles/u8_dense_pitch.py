"""u8 RGB views whose tile footprint needs 145..256 bytes per staged line (no bank-disjoint pitch fits TMA's 256-element
box line): staged with the dense pitch (round 2, late) against the direct kernels they fell back to before.
16 x 4K u8 RGB, the bench view, ratio scaled by 0.75 (a tile of 32 output pixels covers ~57 source texels)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bench
import cameracalibrations_b200 as cc

wl = bench.WORKLOADS["c3"]
sz = wl["sz"]
cal = cc.Calibration(wl["intr"][:4], [bench.BENCH_VIEW], 1.0, wl["intr"][4], ["extrinsic.png"])
peak = bench.measured_peak()[0]
src = torch.randint(0, 256, (16, sz[1], sz[0], 3), dtype=torch.uint8, device="cuda")
dst = torch.empty_like(src)
npx = 16 * sz[0] * sz[1]
for scale in (1.0, 0.75, 0.6):
    ratio = cc.get_ratio(bench.geometry(wl), 1.0) * scale
    axs = cc.get_axes(ratio, 1.0, bench.N_CORNERS, sz)
    for coord in ("f64", "f32"):
        for gather in ("auto", "direct"):
            os.environ["CAMCAL_DEBUG"] = "1" if (gather == "auto" and coord == "f64") else ""
            if not os.environ["CAMCAL_DEBUG"]:
                os.environ.pop("CAMCAL_DEBUG")
            cc.warp(cal, 0, src, ratio, axs, coord=coord, gather=gather, out=dst)
            os.environ.pop("CAMCAL_DEBUG", None)
            ms = bench._time_ms(torch, lambda: cc.warp(cal, 0, src, ratio, axs, coord=coord, gather=gather, out=dst), 20)
            print(f"ratio x {scale}: {coord} {gather:6s} {ms:.4f} ms  frac {6 * npx / (ms * 1e-3) / 1e9 / peak:.3f}", flush=True)
