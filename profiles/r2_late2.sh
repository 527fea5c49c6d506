#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -m gpu -x -k "rectify or views or plot or jpeg" 2>&1 | tail -4 > gpurun_out/late2_pytest.txt; cat gpurun_out/late2_pytest.txt
timeout 100 python profiles/cold_call.py 2>&1 | grep -v Warn | tee gpurun_out/late_cold_call.txt
