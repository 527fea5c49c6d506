#!/bin/bash
mkdir -p gpurun_out
timeout 700 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "rectify or views or plot or jpeg" 2>&1 | tail -15 > gpurun_out/late2_pytest.txt; cat gpurun_out/late2_pytest.txt
