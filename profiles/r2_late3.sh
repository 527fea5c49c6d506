#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/late3_bench.json 2> gpurun_out/late3_bench.err; tail -c 300 gpurun_out/late3_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/late3_bench.json') if l.startswith('{')][0])
print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['link_frac'])
print(d['extras']['fit_100_views'])
for k in d['extras']:
    if 'views' in k and 'rectify' in k: print(k, d['extras'][k]['ms'], d['extras'][k]['hbm_frac'])
PY
