#!/bin/bash
# round 2: environment-knob sweep (CAMCAL_STAGES / CAMCAL_FG / CAMCAL_CTAS_PER_SM) of the staged kernels
for st in 4 5 6 8; do echo "STAGES=$st"; CAMCAL_STAGES=$st python profiles/ktime.py c2 f64; done
for fg in 6 8 10 16; do echo "FG=$fg"; CAMCAL_FG=$fg python profiles/ktime.py c2 f64; done
for fg in 4 8 12 16; do echo "FG=$fg"; CAMCAL_FG=$fg python profiles/ktime.py c3; done
for ct in 2 3; do echo "CTAS=$ct"; CAMCAL_CTAS_PER_SM=$ct python profiles/ktime.py c3; done
