#!/bin/bash
for so in "" profiles/variants/lib_*.so; do
  CAMCAL_B200_LIB=${so:+$PWD/$so} python profiles/ktime.py c2 f64 2>&1 | grep -v Warning
done
