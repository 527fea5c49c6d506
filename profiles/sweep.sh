for tps in 8 15 30 60; do for st in 2 3 4; do
for c in f32; do CAMCAL_TPS=$tps CAMCAL_STAGES=$st python bench.py --workload c2 --steps 30 --warmup 5 --no-cpu --no-extras --coord $c --gather tma 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c2 tps=$tps stages=$st $c', round(d['ms_per_step'],3), round(d['roofline']['frac'],3))"; done; done; done
