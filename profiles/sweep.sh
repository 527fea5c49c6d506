#!/bin/bash
for so in "" profiles/variants/lib_tl64.so profiles/variants/lib_tl16.so; do
for fg in 1 2; do
  CAMCAL_B200_LIB=${so:+$PWD/$so} CAMCAL_FG=$fg python profiles/ktime.py c2 2>&1 | grep -v Warning | sed "s/^/fg=$fg /"
done; done
