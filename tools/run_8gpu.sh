#!/bin/bash
# 8-GPU check: headline bench under torchrun, C5 z-slab forward run, multi-GPU parity tests
mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 4 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
tail -1 gpurun_out/bench_${N}gpu.json | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/bench_slab.py --grid 2048 2048 1024 --nt 100 > gpurun_out/slab_${N}gpu.log 2>&1
tail -1 gpurun_out/slab_${N}gpu.log | cut -c1-600
python -m pytest tests/test_gpu_slab.py -x -q 2>&1 | tail -3 | tee gpurun_out/slab_tests_${N}gpu.log
