#!/bin/bash
# 2-GPU check of the final tree: multi-GPU parity tests (z slabs across processes, both exchanges), then the bench line under torchrun
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_slab.py tests/test_gpu_slab_local.py -x -q -o faulthandler_timeout=200 2>&1 | tail -3 | tee gpurun_out/r2b_slab_tests_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu > gpurun_out/r2b_bench_2gpu.json 2> gpurun_out/r2b_bench_2gpu.err
tail -c 300 gpurun_out/r2b_bench_2gpu.err
tail -1 gpurun_out/r2b_bench_2gpu.json | cut -c1-300
