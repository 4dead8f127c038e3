#!/bin/bash
# One-GPU knob sweep of the step's scheduling choices (environment groups, row buckets, position-kernel CTA width);
# every line is a full `bench.py --no-cpu --no-configs` run. Usage (under gpurun): tools/knob_sweep.sh > gpurun_out/knobs.txt
run() {
  env "$@" python bench.py --steps 25 --warmup 10 --no-cpu --no-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'kernel_ms', round(d['roofline']['kernel_ms'],3))"
}
run B200MJ_GROUPS=2
run B200MJ_GROUPS=1
run B200MJ_GROUPS=3
run B200MJ_BUCKETS=8,20
run B200MJ_BUCKETS=12,28
run B200MJ_BUCKETS=6,12,24
run B200MJ_BUCKETS=10,16,32
run B200MJ_EPB_POS=4
run B200MJ_EPB_POS=3
run B200MJ_GROUPS=2
