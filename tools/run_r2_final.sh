#!/bin/bash
# round-2 final evidence on one B200: full GPU suite, the bench line, the launch list of the same command (short nt), ncu --set full of the VD adjoint launch
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | grep -v "wavelength\|instead of\|Grid spacing" | tail -6 | tee gpurun_out/r2_pytest_gpu.log
python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err
tail -c 400 gpurun_out/r2_bench_1gpu.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 1 --warmup 1 --nt 100 --no-cpu --no-extras > gpurun_out/r2_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:vd_fused_kernel<float, float, \(bool\)1' -s 20 -c 1 -o gpurun_out/r2_vd_adj python tools/bench_sim.py --kind vd --n 4096 4096 --nt 40 --check-freq 10 --reps 0 > gpurun_out/r2_ncu_vd_adj.log 2>&1
tail -2 gpurun_out/r2_ncu_vd_adj.log | cut -c1-200
