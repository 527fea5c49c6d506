#!/bin/bash
# late round-2 GPU call: full GPU suite on HEAD, the e2e chunk ramp A/B (same box, back to back),
# and where the time of the views call goes (host enqueue vs device, CTAs per SM per launch)
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x 2>&1 | tail -4 > gpurun_out/late_pytest_gpu.txt; cat gpurun_out/late_pytest_gpu.txt
for rep in 1 2; do for r in 0 1; do
  CAMCAL_CHUNK_RAMP=$r python bench.py --steps 20 --warmup 5 --no-cpu --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); e=d['e2e']; print('ramp $r: e2e', round(e['value']/1e3,2), 'Gpix/s link_frac', round(e['link_frac'],3), 'ceiling', round(e['link_ceiling']['duplex_gbs_per_dir_per_gpu'],1), 'value', round(d['value']/1e3,1), 'frac', round(d['roofline']['frac'],3))"
done; done 2>&1 | tee gpurun_out/late_ramp.txt
python profiles/views_diag.py 2>&1 | tee gpurun_out/late_views.txt
