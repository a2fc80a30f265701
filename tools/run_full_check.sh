#!/bin/bash
# full GPU check: parity suite, per-config timings (tools/bench_sim.py), headline bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/full_tests.log
cat gpurun_out/full_tests.log
rm -f gpurun_out/full_timings.log
B="python tools/bench_sim.py"
{
$B --kind ela --n 4096 2048 --nt 100 --check-freq 10 --dtype f32 --fast-f32 1 --nrec 10 2>&1 | tail -1
$B --kind ela --n 4096 2048 --nt 100 --check-freq 10 --dtype f32 --fast-f32 0 --nrec 10 2>&1 | tail -1
$B --kind ela --n 4096 2048 --nt 100 --check-freq 10 --dtype f64 --nrec 10 2>&1 | tail -1
$B --kind cd --n 4096 4096 --nt 100 --check-freq 10 2>&1 | tail -1
$B --kind cd --n 512 512 512 --nt 40 --check-freq 10 2>&1 | tail -1
$B --kind vd --n 4096 4096 --nt 100 --check-freq 10 2>&1 | tail -1
} > gpurun_out/full_timings.log 2>&1
cat gpurun_out/full_timings.log | cut -c1-600
python bench.py > gpurun_out/full_bench.json 2> gpurun_out/full_bench.err
cat gpurun_out/full_bench.json | cut -c1-300
