run() { python tools/bench_sim.py "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().splitlines()[-1]); print('   ', d['kind'], d['n'], d['dtype'], 'fwd %.1f us %.0f GB/s  adj %.1f us %.0f GB/s' % (d['fwd']['us'], d['fwd']['GBps'], d['adj']['us'], d['adj']['GBps']))"; }
A="--kind ela --n 4096 2048 --nt 200 --check-freq 14 --dtype f32 --nrec 10 --reps 2"
echo default; run $A
echo split; SWB_ELF_SPLIT=1 run $A
echo tz16; SWB_ELF_TZ=16 run $A
echo tz16split; SWB_ELF_TZ=16 SWB_ELF_SPLIT=1 run $A
echo allinterior-bound; SWB_ELF_DEBUG_ALL_INTERIOR=1 run $A
