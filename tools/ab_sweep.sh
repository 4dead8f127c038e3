#!/bin/bash
for rep in 1 2; do
  for e in 1 2 4 5 8; do
    echo -n "rep $rep epb_pos $e: "
    B200MJ_EPB_POS=$e python tools/step_timeline.py 2>&1 | tail -1
  done
done
