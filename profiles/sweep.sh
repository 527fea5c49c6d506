#!/bin/bash
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu -k "rectify or config2 or config3" 2>&1 | tail -5
CAMCAL_DEBUG=1 python profiles/ktime.py c2 f32 auto 1 2>&1 | grep camcal | head -1
CAMCAL_DEBUG=1 python profiles/ktime.py c3 f32 auto 1 2>&1 | grep camcal | head -1
for fg in 3 4 5 6 8 10 12 16; do
  CAMCAL_FG=$fg python profiles/ktime.py c2 2>&1 | grep -v Warning | sed "s/^default/fg=$fg/"
done
for fg in 1 2 4 8 16; do
  CAMCAL_FG=$fg python profiles/ktime.py c3 2>&1 | grep -v Warning | sed "s/^default/fg=$fg/"
done
