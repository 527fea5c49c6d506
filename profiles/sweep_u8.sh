#!/bin/bash
# round 2: u8 RGB (workload c3) timing of the default build and every profiles/variants/lib_*.so
mkdir -p gpurun_out
for so in "" ${VARIANTS:-profiles/variants/lib_*.so}; do
  CAMCAL_B200_LIB=${so:+$PWD/$so} python profiles/ktime.py c3 2>&1 | grep -v Warning
done
