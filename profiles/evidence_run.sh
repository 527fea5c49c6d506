#!/bin/bash
# one GPU call: full GPU test suite, smoke, plain bench lines (both arms), then the profiler passes (never a bench value)
mkdir -p gpurun_out
t0=$(date +%s); lap() { echo "[lap] $1: $(( $(date +%s) - t0 )) s"; }
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -3 > gpurun_out/r2_pytest_gpu.txt; cat gpurun_out/r2_pytest_gpu.txt; lap pytest
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2; lap smoke
timeout 400 python bench.py --impl reference > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; lap bench_ref
timeout 400 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -c 600 gpurun_out/r2_bench.json; lap bench
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2_ncu_launch.log 2>&1; lap ncu_launches
timeout 600 ncu --set full --import-source on --clock-control none -o gpurun_out/r2_all -f python profiles/prof_kernels.py all --reps 1 > gpurun_out/r2_ncu_full.log 2>&1; lap ncu_full
ls -la gpurun_out/r2_*
