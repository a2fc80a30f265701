python -m pytest tests/test_gpu_acoustic.py -m gpu -q -x -p no:logging 2>&1 | grep -v "points per wavelength\|instead of\|Grid spacing" | tail -30 > gpurun_out/cd_tests3.log
B="python tools/bench_sim.py --kind cd"
{
for z in 16 32 64; do echo "zc=$z 3d512 fwd"; SWB_CDF_ZC=$z $B --n 512 512 512 --nt 30 --no-grad 2>/dev/null | tail -1; done
echo "3d512 serial rim"; SWB_CDF_SERIAL_RIM=1 $B --n 512 512 512 --nt 30 --no-grad 2>/dev/null | tail -1
echo "3d768 fwd"; $B --n 768 768 768 --nt 30 --no-grad 2>/dev/null | tail -1
echo "3d512 grad"; $B --n 512 512 512 --nt 40 --check-freq 10 2>/dev/null | tail -1
echo "3d512 grad faithful"; $B --n 512 512 512 --nt 40 --check-freq 10 --fast-f32 0 2>/dev/null | tail -1
for z in 8 16 32; do echo "zc=$z 2d4096 fwd"; SWB_CDF_ZC=$z $B --n 4096 4096 --nt 200 --no-grad 2>/dev/null | tail -1; done
echo "2d4096 grad"; $B --n 4096 4096 --nt 200 --check-freq 14 2>/dev/null | tail -1
echo "3d512 unfused fwd"; $B --n 512 512 512 --nt 30 --no-grad --fused 0 2>/dev/null | tail -1
} > gpurun_out/cd_bench3.log 2>&1
