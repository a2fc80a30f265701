#!/bin/bash
python -m pytest tests/test_gpu_acoustic.py -x -q 2>&1 | tail -3
B="python tools/bench_sim.py"
$B --kind vd --n 4096 4096 --nt 100 --check-freq 10 2>&1 | tail -1 | tee gpurun_out/vd_timing.log
ncu --set full --clock-control none --import-source on -k regex:vd_fused -s 20 -c 1 -o gpurun_out/vd_fwd_s4 $B --kind vd --n 4096 4096 --nt 40 --no-grad --reps 0 > gpurun_out/ncu_vd.log 2>&1
