#!/bin/bash
# frames per map group (CAMCAL_FG) x library variant
for so in ${LIBS:-profiles/variants/lib_head.so}; do
  for fg in ${FGS:-10 12 15 20 30}; do
    echo -n "FG=$fg "; CAMCAL_FG=$fg CAMCAL_B200_LIB=$PWD/$so python profiles/ktime.py ${WL:-c2} ${CO:-f64} 2>&1 | tail -1
  done
done
