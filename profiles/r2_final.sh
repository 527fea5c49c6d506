#!/bin/bash
# final GPU call of the round: full GPU suite, smoke, plain bench line, ncu --set full of the views kernels as they are now
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -3 > gpurun_out/r2_pytest_gpu.txt; cat gpurun_out/r2_pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -c 300 gpurun_out/r2_bench.err
timeout 300 ncu --set full --import-source on --clock-control none -o gpurun_out/r2_views -f python profiles/prof_kernels.py views --reps 1 > gpurun_out/r2_ncu_views.log 2>&1; tail -2 gpurun_out/r2_ncu_views.log
