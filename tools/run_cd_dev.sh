#!/bin/bash
python -m pytest tests/test_gpu_acoustic.py tests/test_gpu_slab.py -x -q -k "cd or CD or c1 or eager or slab" 2>&1 | tail -3
B="python tools/bench_sim.py --kind cd --check-freq 10 --nrec 10"
for n in "300 280" "1024 1024" "2048 2048" "4096 4096" "160 160 160" "512 512 512"; do
  echo "n=$n"; $B --n $n --nt 200 2>&1 | tail -1 | cut -c150-520
done
$B --n 300 280 --dtype f64 --nt 1500 --check-freq 1 2>&1 | tail -1 | cut -c150-520
