run() { env "$@" python tools/time_cmu_env.py "$*" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['label'], '| ms', round(d['ms_per_step'],3), 'env/s', round(d['env_steps_per_s']), d['nefc_le'], d['warnings'][:3])"; }
run A=0
run B200MJ_BUCKETS=12,32
run B200MJ_BUCKETS=32,72
run B200MJ_BUCKETS=12,72
run B200MJ_BUCKET_ORDER=1
run B200MJ_ACC_SYNC=0
run B200MJ_ACC_SYNC=1
run B200MJ_DUAL_MIN_NV=0 B200MJ_BUCKETS=12,32
