#!/bin/bash
# z-slab forward run on 2 GPUs: the marching rim (chunks of SWB_CDF_RIM_ZC planes) against the per-vector rim kernel (SWB_CDF_RIM_ZC=0)
for zc in 0 16 32 64; do
  export SWB_CDF_RIM_ZC=$zc
  echo "rim_zc=$zc"
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/bench_slab.py --grid 2048 2048 256 --nt 100 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().splitlines()[-1]); print('   ms_per_step', d.get('ms_per_step'), 'Gcell/s', d.get('value'))"
done
unset SWB_CDF_RIM_ZC
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:cd_ -s 40 -c 4 --csv --log-file gpurun_out/slabgeom_launches.csv python tools/bench_sim.py --kind cd --n 2048 2048 128 --nt 30 --no-grad --reps 0 > /dev/null 2>&1
grep -o 'cd_[a-z_]*kernel.*' gpurun_out/slabgeom_launches.csv | cut -c1-30,150-
