#!/bin/bash
# round-2 evidence: full GPU suite, ncu --set full of the VD adjoint launch and of the 3D CD bulk / rim launches at 768^3, launch list of bench.py
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | grep -v "wavelength\|instead of\|Grid spacing" | tail -8 | tee gpurun_out/r2_pytest_gpu.log
B="python tools/bench_sim.py"
ncu --set full --clock-control none --import-source on -k regex:vd_fused -s 60 -c 1 -o gpurun_out/r2_vd_adj $B --kind vd --n 4096 4096 --nt 40 --check-freq 10 --reps 0 > gpurun_out/r2_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cd_bulk -s 10 -c 1 -o gpurun_out/r2_cd3d_bulk_768 $B --kind cd --n 768 768 768 --nt 20 --no-grad --reps 0 >> gpurun_out/r2_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cd_rim -s 10 -c 1 -o gpurun_out/r2_cd3d_rim_768 $B --kind cd --n 768 768 768 --nt 20 --no-grad --reps 0 >> gpurun_out/r2_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 1 --warmup 1 --nt 100 --no-cpu --no-extras > gpurun_out/r2_ncu_bench.log 2>&1
tail -3 gpurun_out/r2_ncu.log
