#!/bin/bash
for so in "" profiles/variants/lib_sleep0.so profiles/variants/lib_sleep32.so; do
for st in 2 3; do for tps in 2 8 30; do
  CAMCAL_B200_LIB=${so:+$PWD/$so} CAMCAL_STAGES=$st CAMCAL_TPS=$tps python profiles/ktime.py c2 f32 2>&1 | grep -v Warning | sed "s/^/st=$st tps=$tps /"
done; done; done
