python -m pytest tests/test_gpu_elastic.py -m gpu -q -x -p no:logging 2>&1 | grep -v "points per wavelength\|instead of\|Grid spacing" | tail -30 > gpurun_out/ela_tests3.log
B="python tools/bench_sim.py"
{
echo "ela 4096x2048 grad f32 fast"; $B --kind ela --n 4096 2048 --nt 100 --check-freq 10 --nrec 10 2>&1 | tail -1
echo "ela 4096x2048 grad f32 faithful"; $B --kind ela --n 4096 2048 --nt 100 --check-freq 10 --nrec 10 --fast-f32 0 2>&1 | tail -1
echo "ela 4096x2048 grad f64"; $B --kind ela --n 4096 2048 --nt 100 --check-freq 10 --nrec 10 --dtype f64 2>&1 | tail -1
} > gpurun_out/ela_bench3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ela_fused -s 20 -c 1 -o gpurun_out/ela_fused_v2 $B --kind ela --n 4096 2048 --nt 40 --no-grad --nrec 10 --reps 0 > gpurun_out/ncu_ela_v2.log 2>&1
