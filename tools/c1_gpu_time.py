#!/usr/bin/env python
"""Wall time of BASELINE config 1 (examples/simple_example_acoustic.jl verbatim) through the public API on one B200:
swforward of 3 shots, swgradient of 3 shots (check_freq 1, mute 5/2), second call timed (graphs captured, buffers allocated)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import swb200 as S

nt, dt, nx, nz, dh, f0 = 1500, 0.001, 300, 280, 8.0, 12.0
velmod = np.zeros((nx, nz), order="F")
velmod[:, :] = 2000.0 + 12.0 * np.arange(nz)[None, :]
t = np.arange(nt) * dt
ixsrc = np.round(np.linspace(32, nx - 31, 3)).astype(int)
ixrec = np.round(np.linspace(30, nx - 29, 10)).astype(int)
stf = (1000.0 * S.rickerstf(t, 1.20 / f0, f0)).reshape(nt, 1)
posrecs = np.zeros((10, 2))
posrecs[:, 0] = (ixrec - 1) * dh
posrecs[:, 1] = 2 * dh
def mk():
    return [S.ScalarShot(srcs=S.ScalarSources(np.array([[(ixsrc[i] - 1) * dh, (nz - 40) * dh]]), stf.copy(), f0), recs=S.ScalarReceivers(posrecs.copy(), nt)) for i in range(3)]
bc = S.CPMLBoundaryConditionParameters(halo=20, rcoef=0.0001, freeboundtop=True)
params = S.InputParametersAcoustic(nt, dt, (nx, nz), (dh, dh), bc)
rp = S.RunParameters(parall="B200")
gp = S.GradParameters(mute_radius_src=5, mute_radius_rec=2, compute_misfit=True, check_freq=1)
mat = S.VpAcousticCDMaterialProperties(velmod)
ws = S.build_wavesim(params, mat, runparams=rp)
out = {}
for rep in range(2):
    shots = mk()
    t0 = time.perf_counter()
    S.swforward(ws, mat, shots)
    out["forward_3shots_s"] = time.perf_counter() - t0
obs = [s.recs.seismograms.copy() for s in shots]
newvel = velmod - 0.2
newvel[29:40, 32:44] *= 0.9
mat2 = S.VpAcousticCDMaterialProperties(newvel)
wg = S.build_wavesim(params, mat2, runparams=rp, gradparams=gp, gradient=True)
for rep in range(2):
    shots = mk()
    t0 = time.perf_counter()
    g, m = S.swgradient(wg, mat2, shots, [S.L2Misfit(observed=o) for o in obs])
    out["gradient_3shots_s"] = time.perf_counter() - t0
out["misfit"] = float(m)
print(out)
