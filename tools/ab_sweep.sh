#!/bin/bash
for rep in 1 2; do
  for cfg in "2 10,24" "2 6,14,28" "2 8,16,32" "2 10,20,40" "2 4,10,24"; do
    set -- $cfg
    echo -n "rep $rep split $1 buckets $2: "
    B200MJ_SPLIT=$1 B200MJ_BUCKETS=$2 python tools/step_timeline.py 2>&1 | tail -1
  done
done
