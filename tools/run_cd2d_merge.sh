#!/bin/bash
# 2D CD on large grids: forked bulk + rim launches vs the merged single launch (SWB_CDF_MERGE_MAX raises the merged form's size limit)
run() { timeout 200 python tools/bench_sim.py "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().splitlines()[-1]); print('   ', d['kind'], d['n'], d['dtype'], 'fwd %.2f us %.0f GB/s  adj %.2f us %.0f GB/s' % (d['fwd']['us'], d['fwd']['GBps'], d['adj']['us'], d['adj']['GBps']))"; }
for mm in default 1000000000; do
  unset SWB_CDF_MERGE_MAX; [ $mm != default ] && export SWB_CDF_MERGE_MAX=$mm
  echo "merge_max=$mm"
  run --kind cd --n 4096 4096 --nt 300 --check-freq 17 --reps 3
  run --kind cd --n 2560 2560 --nt 300 --check-freq 17 --reps 3
  run --kind cd --n 8192 8192 --nt 100 --check-freq 10 --reps 2
done
