#!/bin/bash
for st in 2 3; do
for so in "" profiles/variants/lib_*.so; do
  CAMCAL_B200_LIB=${so:+$PWD/$so} CAMCAL_STAGES=$st python profiles/ktime.py c2 2>&1 | grep -v Warning | sed "s/^/st=$st /"
done; done
