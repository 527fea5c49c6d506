#!/bin/bash
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k rectify 2>&1 | tail -5
for so in "" profiles/variants/lib_*.so; do
  CAMCAL_B200_LIB=${so:+$PWD/$so} python profiles/ktime.py c3 2>&1 | grep -v Warning
done
