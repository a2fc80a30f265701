#!/bin/bash
# development run of the z-marching 3D rim: CD parity tests, marched rim timings, per-kind times of the rim launch
timeout 300 python -m pytest -o faulthandler_timeout=120 tests/test_gpu_acoustic.py tests/test_gpu_slab_local.py -x -q -k "cd or CD or 3d or 3D or slab or eager" 2>&1 | tail -4
run() { timeout 200 python tools/bench_sim.py "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().splitlines()[-1]); print('   ', d['kind'], d['n'], 'fwd %.1f us %.0f GB/s  adj %.1f us %.0f GB/s' % (d['fwd']['us'], d['fwd']['GBps'], d['adj']['us'], d['adj']['GBps']))"; }
B="python tools/bench_sim.py"
run --kind cd --n 512 512 512 --nt 40 --check-freq 10 --reps 2
run --kind cd --n 768 768 768 --nt 30 --check-freq 10 --reps 1
for skip in 0 6 5 3; do
  SWB_CDF_RIM_SKIP=$skip timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:cd_rimz -s 10 -c 1 --csv --log-file gpurun_out/rimz_skip$skip.csv $B --kind cd --n 768 768 768 --nt 16 --no-grad --reps 0 > /dev/null 2>&1
  echo "skip=$skip" $(grep -o 'cd_rimz_kernel.*' gpurun_out/rimz_skip$skip.csv | cut -c150- | tail -3 | tr '\n' ' ')
done
