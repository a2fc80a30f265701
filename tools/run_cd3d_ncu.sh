#!/bin/bash
B="python tools/bench_sim.py"
$B --kind cd --n 2048 2048 128 --nt 30 --no-grad 2>&1 | tail -1 | tee gpurun_out/cd3d_slab_timing.log
ncu --set full --clock-control none --import-source on -k regex:cd_bulk -s 10 -c 1 -o gpurun_out/cd3d_bulk_2048 $B --kind cd --n 2048 2048 128 --nt 20 --no-grad --reps 0 > gpurun_out/ncu_cd3d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cd_rim -s 10 -c 1 -o gpurun_out/cd3d_rim_2048 $B --kind cd --n 2048 2048 128 --nt 20 --no-grad --reps 0 >> gpurun_out/ncu_cd3d.log 2>&1
