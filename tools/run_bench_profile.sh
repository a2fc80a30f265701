#!/bin/bash
# headline bench line + the ncu launch list of the same command (short nt) + ncu --set full of the VD forward / adjoint launches
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
cut -c1-300 gpurun_out/bench_final.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_bench_final.csv python bench.py --steps 1 --warmup 1 --nt 100 --no-cpu --no-extras > gpurun_out/ncu_bench_final.log 2>&1
B="python tools/bench_sim.py"
ncu --set full --clock-control none --import-source on -k regex:vd_fused -s 20 -c 1 -o gpurun_out/vd_fwd_v6 $B --kind vd --n 4096 4096 --nt 40 --no-grad --reps 0 > gpurun_out/ncu_vd6.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:vd_fused -s 60 -c 1 -o gpurun_out/vd_adj_v6 $B --kind vd --n 4096 4096 --nt 40 --check-freq 10 --reps 0 >> gpurun_out/ncu_vd6.log 2>&1
