#!/bin/bash
mkdir -p gpurun_out
timeout 700 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "views or plot or jpeg or bounds or tilted" 2>&1 | tail -5 > gpurun_out/late2_pytest.txt; cat gpurun_out/late2_pytest.txt
timeout 200 python profiles/views_diag.py short 2>&1 | tee gpurun_out/late2_views.txt
timeout 200 python - <<'PY' 2>&1 | tee -a gpurun_out/late2_views.txt
import sys; sys.argv=['x']
import torch, bench, cameracalibrations_b200 as cc
wl3 = bench.WORKLOADS["c3"]; sz3 = wl3["sz"]; BV=bench.BENCH_VIEW
vlist = [((BV[0][0] + 0.002 * i, BV[0][1], BV[0][2] + 0.001 * i), (BV[1][0] + 0.01 * i, BV[1][1], BV[1][2] + 0.05 * i)) for i in range(16)]
cal3 = cc.Calibration(wl3["intr"][:4], vlist, 1.0, wl3["intr"][4], [f"{i}.png" for i in range(16)])
ratio3 = cc.get_ratio(bench.geometry(wl3), 1.0); axs3 = cc.get_axes(ratio3, 1.0, bench.N_CORNERS, sz3)
src3 = torch.randint(0, 256, (16, sz3[1], sz3[0], 3), dtype=torch.uint8, device="cuda"); dst3 = torch.empty_like(src3)
for coord in ("f64", "f32"):
    ms = bench._time_ms(torch, lambda: cc.warp_views(cal3, list(range(16)), src3, [ratio3] * 16, [axs3] * 16, coord=coord, out=dst3), 20)
    print("u8 views 16x4K", coord, round(ms, 4), "ms")
PY
