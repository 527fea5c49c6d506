#!/bin/bash
# quick GPU iteration: rectify parity tests, then short bench lines (device-resident only)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "rectify" 2>&1 | tail -5
for w in ${WL:-c2}; do for c in f64 f32; do
python bench.py --workload $w --steps 30 --warmup 5 --no-cpu --no-extras --coord $c --gather ${GATHER:-auto} 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w $c', round(d['ms_per_step'],4),'ms frac', round(d['roofline']['frac'],3), round(d['value']/1e3,1),'Gpix/s', 'e2e', round(d['e2e']['value']/1e3,1))"
done; done
