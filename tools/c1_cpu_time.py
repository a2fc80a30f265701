#!/usr/bin/env python
"""Wall time of BASELINE config 1 (examples/simple_example_acoustic.jl: 300x280 Float64, nt 1500, 3 shots, check_freq 1) with the
CPU restatement of the reference's Threads backend (oracle, OpenMP) on this machine's cores: forward of 3 shots + gradient of 3 shots."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O

O.build()
O.use_openmp(True)
nt, dt, nx, nz, dh, f0 = 1500, 0.001, 300, 280, 8.0, 12.0
velmod = np.zeros((nx, nz), order="F")
velmod[:, :] = 2000.0 + 12.0 * np.arange(nz)[None, :]
t = np.arange(nt) * dt
ixsrc = np.round(np.linspace(32, nx - 31, 3)).astype(int)
ixrec = np.round(np.linspace(30, nx - 29, 10)).astype(int)
stf = (1000.0 * O.rickerstf(t, 1.20 / f0, f0)).reshape(nt, 1)
posrecs = np.zeros((10, 2))
posrecs[:, 0] = (ixrec - 1) * dh
posrecs[:, 1] = 2 * dh
shots = [O.ScalarShot(src_positions=np.array([[(ixsrc[i] - 1) * dh, (nz - 40) * dh]]), src_tf=np.asfortranarray(stf.copy()), domfreq=f0, rec_positions=posrecs.copy())
         for i in range(3)]
params = O.Params(nt=nt, dt=dt, gridsize=(nx, nz), spacing=(dh, dh), halo=20, rcoef=0.0001, freetop=True)
t0 = time.perf_counter()
O.swforward(O.build_wavesim("acoustic_cd", params), [velmod], shots)
t_fwd = time.perf_counter() - t0
obs = [s.seismograms.copy() for s in shots]
newvel = velmod - 0.2
newvel[29:40, 32:44] *= 0.9
sim = O.build_wavesim("acoustic_cd", params, gradient=True, check_freq=1)
mis = [O.L2Misfit(observed=o) for o in obs]
t0 = time.perf_counter()
O.swgradient(sim, [newvel], shots, mis)
t_grad = time.perf_counter() - t0
print({"cores": os.cpu_count(), "omp_threads": os.environ.get("OMP_NUM_THREADS"), "forward_3shots_s": t_fwd, "gradient_3shots_s": t_grad})
