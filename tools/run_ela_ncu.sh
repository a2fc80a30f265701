#!/bin/bash
# ncu --set full of one forward launch and one correlating adjoint launch of the fused elastic kernel at C3 size
B="python tools/bench_sim.py"
ncu --set full --clock-control none --import-source on -k regex:ela_fused -s 20 -c 1 -o gpurun_out/ela_fused_v6_fwd $B --kind ela --n 4096 2048 --nt 40 --no-grad --nrec 10 --reps 0 > gpurun_out/ncu_ela_v6.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ela_fused -s 49 -c 1 -o gpurun_out/ela_fused_v6_adj $B --kind ela --n 4096 2048 --nt 40 --check-freq 10 --nrec 10 --reps 0 >> gpurun_out/ncu_ela_v6.log 2>&1
