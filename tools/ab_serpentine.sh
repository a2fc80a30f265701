run() { python tools/bench_sim.py "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().splitlines()[-1]); print('   ', d['kind'], d['n'], d['dtype'], 'fwd %.2f us %.0f GB/s  adj %.2f us %.0f GB/s' % (d['fwd']['us'], d['fwd']['GBps'], d['adj']['us'], d['adj']['GBps']))"; }
for mode in off on; do
  if [ $mode = off ]; then export SWB_ELF_NO_SERPENTINE=1 SWB_CDF_NO_SERPENTINE_UNUSED=1; else unset SWB_ELF_NO_SERPENTINE; export SWB_CDF_SERPENTINE=1; fi
  echo "serpentine $mode"
  run --kind ela --n 4096 2048 --nt 200 --check-freq 14 --dtype f32 --nrec 10 --reps 2
  run --kind ela --n 4096 2048 --nt 100 --check-freq 10 --dtype f64 --nrec 10 --reps 2
  run --kind cd --n 4096 4096 --nt 200 --check-freq 14 --reps 2
  run --kind cd --n 512 512 512 --nt 40 --check-freq 10 --reps 2
done
