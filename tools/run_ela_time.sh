#!/bin/bash
# CUDA-event timings of the fused elastic engine at C3 size (no tests)
for cfg in "f32 1" "f32 0" "f64 0"; do
  set -- $cfg
  python tools/bench_sim.py --kind ela --n 4096 2048 --nt 60 --check-freq 10 --dtype $1 --fast-f32 $2 --nrec 10 2>&1 | tail -1 | tee -a gpurun_out/ela_timings.log
done
