#!/bin/bash
# chunk length of the marched x / y strip boxes (SWB_CDF_RIM_ZC): forward step times
run() { timeout 100 python tools/bench_sim.py "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().splitlines()[-1]); print('   ', d['kind'], d['n'], 'fwd %.1f us %.0f GB/s' % (d['fwd']['us'], d['fwd']['GBps']))"; }
for zc in ${ZCS:-8 12 16 24 32}; do
  export SWB_CDF_RIM_ZC=$zc
  echo "rim_zc=$zc"
  run --kind cd --n 768 768 768 --nt 30 --no-grad --reps 2
  run --kind cd --n 512 512 512 --nt 40 --no-grad --reps 2
done
