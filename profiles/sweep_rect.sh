#!/bin/bash
# round 2: c2 (fp32 gray) and c3 (u8 RGB) timing of the default build and every profiles/variants/lib_*.so
mkdir -p gpurun_out
for so in "" ${VARIANTS:-profiles/variants/lib_*.so}; do
  for w in ${WL:-c2 c3}; do
    CAMCAL_B200_LIB=${so:+$PWD/$so} python profiles/ktime.py $w 2>&1 | grep -v Warning
  done
done
