#!/bin/bash
python -m pytest tests/test_gpu_acoustic.py -x -q 2>&1 | tail -3
B="python tools/bench_sim.py"
$B --kind cd --n 512 512 512 --nt 40 --check-freq 10 2>&1 | tail -1 | tee gpurun_out/cd_timing.log
$B --kind cd --n 768 768 768 --nt 30 --no-grad 2>&1 | tail -1 | tee -a gpurun_out/cd_timing.log
$B --kind cd --n 4096 4096 --nt 100 --check-freq 10 2>&1 | tail -1 | tee -a gpurun_out/cd_timing.log
