#!/bin/bash
for rep in 1 2; do
  for g in 1 2 3; do
    echo -n "rep $rep groups $g: "
    B200MJ_GROUPS=$g python tools/step_timeline.py 2>&1 | tail -1
  done
done
