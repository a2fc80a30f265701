#!/usr/bin/env python
"""cProfile of the public swgradient call on the C2 geometry with few time steps: shows the host-side cost per shot."""
import cProfile, pstats, sys, os, io
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import swb200 as S

prob = bench.c2_problem(n=4096, nt=100)
n, nt, h = prob["n"], prob["nt"], prob["h"]
T = np.float32
bc = S.CPMLBoundaryConditionParameters(halo=prob["halo"], rcoef=T(1e-4), freeboundtop=True)
params = S.InputParametersAcoustic(nt, T(prob["dt"]), (n, n), (T(h), T(h)), bc, dtype=np.dtype(T))
rp_ = S.RunParameters(parall="B200", erroronPPW=False, fast_f32=True)
gp = S.GradParameters(mute_radius_src=3, mute_radius_rec=0, compute_misfit=True, check_freq=10)
matprop = S.VpRhoAcousticVDMaterialProperties(prob["vp"], prob["rho"])
tf = np.asfortranarray((1000.0 * S.gaussderivstf(prob["t"], 2.0 / prob["f0"], prob["f0"])).astype(T).reshape(nt, 1))
rp = np.stack([prob["xr"], np.full_like(prob["xr"], 3 * h)], axis=1).astype(T)
def make_shot(g):
    sp = np.array([[prob["xs"][g], 2 * h]], dtype=T)
    return S.ScalarShot(srcs=S.ScalarSources(sp, tf, T(prob["f0"])), recs=S.ScalarReceivers(rp, nt, dtype=np.dtype(T)))
ws = S.build_wavesim(params, matprop, runparams=rp_, gradparams=gp, gradient=True)
obs = np.zeros((nt, rp.shape[0]), dtype=T, order="F")
shots = [make_shot(k) for k in range(3)]
S.swgradient(ws, matprop, shots, [S.L2Misfit(observed=obs) for _ in shots])
shots = [make_shot(k) for k in range(3, 9)]
pr = cProfile.Profile()
pr.enable()
S.swgradient(ws, matprop, shots, [S.L2Misfit(observed=obs) for _ in shots])
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
print(s.getvalue())
