"""Per-call cost of single-frame rectification (the reference's plot() pattern: one frame per view)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, bench
import cameracalibrations_b200 as cc
wl = bench.WORKLOADS["c2"]; sz = wl["sz"]
ratio = cc.get_ratio(bench.geometry(wl), 1.0); axs = cc.get_axes(ratio, 1.0, bench.N_CORNERS, sz)
c = cc.Calibration(wl["intr"][:4], [bench.BENCH_VIEW], 1.0, wl["intr"][4], ["extrinsic.png"])
fr = torch.rand((1, sz[1], sz[0]), device="cuda"); out = torch.empty_like(fr)
for coord in ("f64", "f32"):
    for _ in range(20): cc.warp(c, 0, fr, ratio, axs, coord=coord, out=out)
    torch.cuda.synchronize(); n = 2000
    t = time.perf_counter()
    for _ in range(n): cc.warp(c, 0, fr, ratio, axs, coord=coord, out=out)
    t_enq = time.perf_counter() - t
    torch.cuda.synchronize(); t_all = time.perf_counter() - t
    print(f"{coord}: enqueue {t_enq / n * 1e6:.1f} us/call, end-to-end {t_all / n * 1e6:.1f} us/frame ({sz[0] * sz[1] / (t_all / n) / 1e9:.1f} Gpix/s)")
# the C entry point alone (arguments prebuilt): what a compiled host (Julia ccall) pays per call
import ctypes as C
from cameracalibrations_b200 import _lib
h = _lib.context(0).handle
axs_c = (C.c_int64 * 2)(*axs)
ci, cv = C.byref(c._intr), C.byref(c._views[0])
args = (h, ci, cv, float(ratio), axs_c, C.c_void_p(fr.data_ptr()), C.c_void_p(out.data_ptr()), sz[0], sz[1],
        C.c_size_t(sz[0]), C.c_size_t(sz[0] * sz[1]), 1, C.c_float(0.0), 1, None)
fn = _lib.lib.cc_rectify_f32c1
for _ in range(20): fn(*args)
torch.cuda.synchronize(); n = 5000
t = time.perf_counter()
for _ in range(n): fn(*args)
t_enq = time.perf_counter() - t
torch.cuda.synchronize(); t_all = time.perf_counter() - t
print(f"C ABI only (f32 coords): enqueue {t_enq / n * 1e6:.1f} us/call, end-to-end {t_all / n * 1e6:.1f} us/frame")
