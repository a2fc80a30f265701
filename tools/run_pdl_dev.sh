#!/bin/bash
# programmatic dependent launch of the step kernels: parity tests, then per-launch times with and without (SWB_NO_PDL=1)
KIND=${KIND:-vd}
timeout 400 python -m pytest -o faulthandler_timeout=150 tests/test_gpu_acoustic.py tests/test_gpu_benchmarked_mode_parity.py -x -q -k "${TESTK:-vd or VD or eager or c2}" 2>&1 | tail -4
run() { timeout 200 python tools/bench_sim.py "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().splitlines()[-1]); print('   ', d['kind'], d['n'], 'fwd %.2f us %.0f GB/s  adj %.2f us %.0f GB/s' % (d['fwd']['us'], d['fwd']['GBps'], d['adj']['us'], d['adj']['GBps']))"; }
for pdl in on off on off; do
  unset SWB_NO_PDL; [ $pdl = off ] && export SWB_NO_PDL=1
  echo "pdl=$pdl"
  for n in ${SIZES:-"4096 4096" "2048 2048"}; do
    run --kind $KIND --n $n --nt 300 --check-freq 17 --reps 3
  done
done
