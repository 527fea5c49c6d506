#!/bin/bash
mkdir -p gpurun_out
timeout 700 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "views or plot or jpeg or bounds" 2>&1 | tail -5 > gpurun_out/late2_pytest.txt; cat gpurun_out/late2_pytest.txt
timeout 200 python profiles/views_diag.py short 2>&1 | tee gpurun_out/late2_views.txt
CAMCAL_VIEWS_SINGLE=0 timeout 200 python profiles/views_diag.py short 2>&1 | tee -a gpurun_out/late2_views.txt
