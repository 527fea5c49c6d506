# A/B of library variants (LIBS: paths, "default" = the built library), WL/CO = workload / coords, two rounds
for r in 1 2; do for so in ${LIBS:-profiles/variants/lib_head.so default}; do
  if [ "$so" = default ]; then python profiles/ktime.py ${WL:-c2} ${CO:-f64} 2>&1 | tail -1
  else CAMCAL_B200_LIB=$PWD/$so python profiles/ktime.py ${WL:-c2} ${CO:-f64} 2>&1 | tail -1; fi
done; done
