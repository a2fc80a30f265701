#!/bin/bash
python -m pytest tests/test_gpu_acoustic.py -x -q 2>&1 | tail -3
python tools/bench_sim.py --kind vd --n 4096 4096 --nt 100 --check-freq 10 2>&1 | tail -1 | tee gpurun_out/vd_timing.log
