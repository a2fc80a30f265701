#!/bin/bash
# round-2 final evidence on one B200: full GPU suite, the bench line (wall time noted), the launch list of the same command (short nt),
# ncu --set full of the marching 3D CD rim launch
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -o faulthandler_timeout=300 2>&1 | grep -v "wavelength\|instead of\|Grid spacing" | tail -6 | tee gpurun_out/r2b_pytest_gpu.log
SECONDS=0
timeout 900 python bench.py > gpurun_out/r2b_bench_1gpu.json 2> gpurun_out/r2b_bench_1gpu.err
echo "bench.py wall seconds: $SECONDS" | tee -a gpurun_out/r2b_bench_1gpu.err
tail -c 300 gpurun_out/r2b_bench_1gpu.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2b_launches_bench.csv python bench.py --steps 1 --warmup 1 --nt 100 --no-cpu --no-extras > gpurun_out/r2b_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cd_rimz -s 10 -c 1 -f -o gpurun_out/r2b_cd3d_rimz_768 python tools/bench_sim.py --kind cd --n 768 768 768 --nt 20 --no-grad --reps 0 > gpurun_out/r2b_ncu_rimz.log 2>&1
tail -2 gpurun_out/r2b_ncu_rimz.log | cut -c1-200
