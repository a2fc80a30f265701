#!/bin/bash
B="python tools/bench_sim.py"
for skip in 5 3; do
SWB_CDF_RIM_SKIP=$skip timeout 300 ncu --set full --clock-control none --import-source on -k regex:cd_rimz -s 10 -c 1 -f -o gpurun_out/rimz_skip$skip $B --kind cd --n 768 768 768 --nt 16 --no-grad --reps 0 > gpurun_out/ncu_rimz_skip$skip.log 2>&1
done
