#!/bin/bash
# development run of the fused elastic engine on one B200: parity tests, then CUDA-event timings at C3 size
mkdir -p gpurun_out
python -m pytest tests/test_gpu_elastic.py -x -q 2>&1 | tail -8 > gpurun_out/ela_tests.log
cat gpurun_out/ela_tests.log
SWB_ELF_TZ=16 python -m pytest tests/test_gpu_elastic.py -x -q -k "fused" 2>&1 | tail -4 | tee gpurun_out/ela_tests_tz16.log
B="python tools/bench_sim.py --kind ela --n 4096 2048 --nt 100 --check-freq 10 --nrec 10"
for cfg in "f32 1" "f32 0" "f64 0"; do
  set -- $cfg
  $B --dtype $1 --fast-f32 $2 2>&1 | tail -1 | tee -a gpurun_out/ela_timings.log
done
SWB_ELF_TZ=16 $B --dtype f32 --fast-f32 1 2>&1 | tail -1 | tee -a gpurun_out/ela_timings.log
