#!/bin/bash
# development run of the 3D rim (x / y strip boxes marched along z, z-plane boxes per vector): CD parity tests, step timings
timeout 300 python -m pytest -o faulthandler_timeout=120 tests/test_gpu_acoustic.py tests/test_gpu_slab_local.py tests/test_gpu_benchmarked_mode_parity.py -x -q -k "cd or CD or 3d or 3D or slab or eager or c4 or C4" 2>&1 | tail -4
run() { timeout 200 python tools/bench_sim.py "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().splitlines()[-1]); print('   ', d['kind'], d['n'], 'fwd %.1f us %.0f GB/s  adj %.1f us %.0f GB/s' % (d['fwd']['us'], d['fwd']['GBps'], d['adj']['us'], d['adj']['GBps']))"; }
for zc in ${ZCS:-16}; do
  export SWB_CDF_RIM_ZC=$zc
  echo "rim_zc=$zc"
  run --kind cd --n 768 768 768 --nt 30 --check-freq 10 --reps 1
done
