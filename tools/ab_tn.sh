#!/bin/bash
# A/B: compile-time-size acceleration kernels (default) vs runtime-size ones, same box, same bench command
for tn in 1 0 1 0; do
  B200MJ_TN=$tn python bench.py --steps 30 --warmup 10 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('TN=$tn value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', d['ms_per_step'])"
done
