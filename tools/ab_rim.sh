run() { python tools/bench_sim.py "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().splitlines()[-1]); print('   ', d['kind'], d['n'], 'fwd %.1f us %.0f GB/s  adj %.1f us %.0f GB/s' % (d['fwd']['us'], d['fwd']['GBps'], d['adj']['us'], d['adj']['GBps']))"; }
for lib in default minb6; do for pf in on off; do
  unset SWB200_LIB SWB_CDF_NO_RIM_PREFETCH
  [ $lib = minb6 ] && export SWB200_LIB=$PWD/seismicwaves.jl_b200/libswb200_minb6.so
  [ $pf = off ] && export SWB_CDF_NO_RIM_PREFETCH=1
  echo "lib=$lib prefetch=$pf"
  run --kind cd --n 512 512 512 --nt 40 --check-freq 10 --reps 2
  run --kind cd --n 768 768 768 --nt 30 --check-freq 10 --reps 1
done; done
