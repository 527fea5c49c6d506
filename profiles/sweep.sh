#!/bin/bash
export CAMCAL_B200_LIB=$PWD/profiles/variants/lib_u8bytes.so
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu -k "u8c3 or ties or config3 or 1e3" 2>&1 | tail -3
python profiles/ktime.py c3
