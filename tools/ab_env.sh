#!/bin/bash
# A/B with environment knobs: each argument is "label:SO_VARIANT:ENV1=V1,ENV2=V2" ; every variant twice, interleaved
for rep in 1 2; do
for spec in "$@"; do
  label=${spec%%:*}; rest=${spec#*:}; so=${rest%%:*}; envs=${rest#*:}
  envs=${envs//,/ }
  env B200MJ_SO=$PWD/build/libb200mj_$so.so $envs python bench.py --steps 30 --warmup 10 --no-cpu --no-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$label value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],3))"
done
done
