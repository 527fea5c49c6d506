for so in "" profiles/variants/lib_*.so; do CAMCAL_B200_LIB=${so:+$PWD/$so} python profiles/ktime.py c2 f64 2>&1 | grep -v Warn; done
for fg in 6 8 10 12 16 22 32; do echo -n "FG$fg "; CAMCAL_FG=$fg python profiles/ktime.py c2 f64 2>&1 | grep -v Warn; done
for st in 2 3 4; do echo -n "ST$st "; CAMCAL_STAGES=$st python profiles/ktime.py c2 f64 2>&1 | grep -v Warn; done
for c in 3 4; do echo -n "CTAS$c "; CAMCAL_CTAS_PER_SM=$c python profiles/ktime.py c2 f64 2>&1 | grep -v Warn; done
