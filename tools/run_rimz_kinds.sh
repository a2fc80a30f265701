#!/bin/bash
# time of the marching rim launch with box kinds skipped (timing experiment; results are wrong while a kind is skipped)
B="python tools/bench_sim.py"
for skip in 0 1 2 4 6 5 3; do
  SWB_CDF_RIM_SKIP=$skip timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:cd_rimz -s 10 -c 1 --csv --log-file gpurun_out/rimz_skip$skip.csv $B --kind cd --n 768 768 768 --nt 16 --no-grad --reps 0 > /dev/null 2>&1
  echo "skip=$skip"; grep -o 'cd_rimz_kernel.*' gpurun_out/rimz_skip$skip.csv | cut -c150- | tail -3
done
