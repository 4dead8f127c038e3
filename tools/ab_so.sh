#!/bin/bash
# A/B of kernel variants built side by side under build/: tools/ab_so.sh v0 v1 ...  (each twice, interleaved)
for rep in 1 2; do
for v in "$@"; do
  B200MJ_SO=$PWD/build/libb200mj_$v.so python bench.py --steps 30 --warmup 10 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],3))"
done
done
