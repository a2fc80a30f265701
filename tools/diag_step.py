"""Host-side phase timing of one bench step (diagnostic; run on the GPU box): python tools/diag_step.py [fast]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import swb200 as S
import torch

fast = len(sys.argv) > 1 and sys.argv[1] == "1"
prob = bench.c2_problem()
n, nt, h = prob["n"], prob["nt"], prob["h"]
T = np.float32
bc = S.CPMLBoundaryConditionParameters(halo=20, rcoef=T(1e-4), freeboundtop=True)
params = S.InputParametersAcoustic(nt, T(prob["dt"]), (n, n), (T(h), T(h)), bc, dtype=np.dtype(T))
rp_ = S.RunParameters(parall="B200", erroronPPW=False, fast_f32=fast)
gp = S.GradParameters(mute_radius_src=3, compute_misfit=True, check_freq=prob["check_freq"])
matprop = S.VpRhoAcousticVDMaterialProperties(prob["vp"], prob["rho"])
tf = np.asfortranarray((1000.0 * S.gaussderivstf(prob["t"], 2.0 / prob["f0"], prob["f0"])).astype(T).reshape(nt, 1))
rp = np.stack([prob["xr"], np.full_like(prob["xr"], 3 * h)], axis=1).astype(T)
ws = S.build_wavesim(params, matprop, runparams=rp_, gradparams=gp, gradient=True)
lib = S._lib.load()
ws.set_wavesim_matprop(matprop)
mis = S._lib.C.c_double()
def sync():
    torch.cuda.synchronize()
for k in range(6):
    sp = np.array([[prob["xs"][k], 2 * h]], dtype=T)
    shot = S.ScalarShot(srcs=S.ScalarSources(sp, tf, T(prob["f0"])), recs=S.ScalarReceivers(rp, nt, dtype=np.dtype(T)))
    sync(); t0 = time.perf_counter()
    ws.init_shot(shot); sync(); t1 = time.perf_counter()
    ws._bind(shot); sync(); t2 = time.perf_counter()
    S._lib.check(lib.swb_sim_gradient_forward(ws._h, None)); sync(); t3 = time.perf_counter()
    S._lib.check(lib.swb_sim_gradient_l2(ws._h, None, None, S._lib.C.byref(mis))); sync(); t4 = time.perf_counter()
    S._lib.check(lib.swb_sim_accumulate_gradient(ws._h, 1, S.api._vp(np.asfortranarray(sp)), 3, rp.shape[0], S.api._vp(np.asfortranarray(rp)), 0)); sync(); t5 = time.perf_counter()
    print(f"step {k}: init_shot {1e3*(t1-t0):.1f} ms, bind {1e3*(t2-t1):.1f}, fwd-only {1e3*(t3-t2):.1f}, fwd+adj {1e3*(t4-t3):.1f}, accumulate {1e3*(t5-t4):.1f}", flush=True)
