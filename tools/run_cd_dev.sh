B="python tools/bench_sim.py --kind cd"
{
echo "3d768 fwd"; $B --n 768 768 768 --nt 30 --no-grad 2>/dev/null | tail -1
echo "2d4096 grad"; $B --n 4096 4096 --nt 200 --check-freq 14 2>/dev/null | tail -1
echo "C4 full: 3d768 grad nt500 cf50"; $B --n 768 768 768 --nt 500 --check-freq 50 --reps 1 --nrec 1024 2>&1 | tail -2
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
} > gpurun_out/cd_bench8.log 2>&1
python -m pytest tests/test_gpu_acoustic.py -m gpu -q -x -p no:logging -k "cd_fused" 2>&1 | tail -3 > gpurun_out/cd_tests8.log
