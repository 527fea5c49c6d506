#!/bin/bash
for mb in 4 8 16 32 64; do
  CAMCAL_CHUNK_MB=$mb python bench.py --steps 5 --warmup 3 --no-cpu --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunk $mb MB: e2e', round(d['e2e']['value']/1e3,2), 'Gpix/s')"
done
