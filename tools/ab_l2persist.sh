run() { python tools/bench_sim.py "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().splitlines()[-1]); print('   ', d['kind'], d['n'], d['dtype'], 'fwd %.2f us %.0f GB/s  adj %.2f us %.0f GB/s  wall %.1f Gcell/s' % (d['fwd']['us'], d['fwd']['GBps'], d['adj']['us'], d['adj']['GBps'], d['Gcell_per_s_wall_last']))"; }
python -c "
import torch
p=torch.cuda.get_device_properties(0); print('L2', p.L2_cache_size)
import ctypes
rt=ctypes.CDLL('libcudart.so.12'); v=ctypes.c_int()
for name,a in (('MaxPersistingL2CacheSize',108),('MaxAccessPolicyWindowSize',109)):
    rt.cudaDeviceGetAttribute(ctypes.byref(v), a, 0); print(name, v.value)
"
for mode in off on on48; do
  unset SWB_NO_L2_PERSIST SWB_L2_PERSIST_MB
  if [ $mode = off ]; then export SWB_NO_L2_PERSIST=1; fi
  if [ $mode = on48 ]; then export SWB_L2_PERSIST_MB=48; fi
  echo "l2 persist $mode"
  run --kind vd --n 4096 4096 --nt 200 --check-freq 14 --reps 2
  run --kind vd --n 2048 2048 --nt 200 --check-freq 14 --reps 2
done
