#!/usr/bin/env python
"""tools/bench_sim.py -- per-launch timing of one simulation kind on one GPU (development aid, not the driver's bench.py).

    python tools/bench_sim.py --kind cd --n 768 768 768 --nt 40 --check-freq 10 [--fast-f32 0] [--no-grad] [--fused 0]

Prints one JSON line: average forward-step and adjoint-step launch durations (CUDA events on the engine's stream, taken
around whole replayed sweeps), the achieved algorithmic HBM bandwidth (SURVEY 8d byte model) and the sweep throughput.
"""
import argparse
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    args = parse_args()
    print(json.dumps(measure(args)), flush=True)


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", default="cd", choices=["cd", "vd", "ela"])
    ap.add_argument("--n", type=int, nargs="+", default=[768, 768, 768])
    ap.add_argument("--nt", type=int, default=40)
    ap.add_argument("--check-freq", type=int, default=10)
    ap.add_argument("--halo", type=int, default=20)
    ap.add_argument("--fast-f32", type=int, default=1)
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--fused", type=int, default=1)
    ap.add_argument("--no-grad", action="store_true")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--nrec", type=int, default=64)
    return ap.parse_args(argv)


def measure(args):
    """run the configuration `args` describes (see parse_args) and return the timing record"""
    import swb200 as S

    S._lib.require_device()
    T = np.float32 if args.dtype == "f32" else np.float64
    n = tuple(args.n)
    N = len(n)
    h = 10.0
    rng = np.random.default_rng(1236)
    depth = np.arange(n[-1], dtype=np.float64) / (n[-1] - 1)
    vp = (1500.0 + 3000.0 * depth.reshape((1,) * (N - 1) + (n[-1],)) + np.zeros(n)).astype(T)
    vp += rng.normal(0, 30.0, size=n).astype(T)
    vp = np.asfortranarray(vp)
    cfl = (6.0 / 7.0 if args.kind in ("vd", "ela") else 1.0)
    dt = 0.99 * cfl * h / (float(vp.max()) * math.sqrt(N))
    nt = args.nt
    bc = S.CPMLBoundaryConditionParameters(halo=args.halo, rcoef=T(1e-4), freeboundtop=True)
    PC = S.InputParametersElastic if args.kind == "ela" else S.InputParametersAcoustic
    params = PC(nt, T(dt), n, tuple(T(h) for _ in n), bc, dtype=np.dtype(T))
    rp = S.RunParameters(parall="B200", erroronPPW=False, fast_f32=bool(args.fast_f32), fused=bool(args.fused))
    gp = S.GradParameters(mute_radius_src=2, mute_radius_rec=0, compute_misfit=True, check_freq=args.check_freq)
    if args.kind == "cd":
        matprop = S.VpAcousticCDMaterialProperties(vp)
    elif args.kind == "ela":
        vp64 = vp.astype(np.float64)
        rho = np.full(n, 2100.0)
        mu = (vp64 / np.sqrt(3.0)) ** 2 * rho
        lam = vp64**2 * rho - 2 * mu
        matprop = S.ElasticIsoMaterialProperties(np.asfortranarray(rho.astype(T)), np.asfortranarray(lam.astype(T)), np.asfortranarray(mu.astype(T)))
    else:
        matprop = S.VpRhoAcousticVDMaterialProperties(vp, np.asfortranarray((310.0 * vp.astype(np.float64) ** 0.25).astype(T)))
    f0 = 8.0
    t = np.arange(nt) * dt
    tf = np.asfortranarray((1000.0 * S.rickerstf(t, 1.2 / f0, f0)).astype(T).reshape(nt, 1))
    ext = [(n[d] - 1) * h for d in range(N)]
    sp = np.array([[0.5 * ext[d] for d in range(N - 1)] + [2 * h]], dtype=T)
    rpos = np.zeros((args.nrec, N), dtype=T)
    for d in range(N - 1):
        rpos[:, d] = np.linspace(0.1, 0.9, args.nrec) * ext[d]
    rpos[:, -1] = 3 * h

    def shot():
        if args.kind == "ela":
            spo = sp + T(0.124)  # off-grid: exercises the sinc spreading
            srcs = S.MomentTensorSources(spo, (tf * T(1e-3)).astype(T), [S.MomentTensor2D(T(5e10), T(5e10), T(0.89e10))], T(f0))
            return S.MomentTensorShot(srcs=srcs, recs=S.VectorReceivers((rpos - T(0.324)).astype(T), nt, dtype=np.dtype(T)))
        return S.ScalarShot(srcs=S.ScalarSources(sp, tf, T(f0)), recs=S.ScalarReceivers(rpos, nt, dtype=np.dtype(T)))

    grad = not args.no_grad
    ws = S.build_wavesim(params, matprop, runparams=rp, gradparams=gp if grad else None, gradient=grad)
    ncell = float(np.prod(n))
    out = {"kind": args.kind, "n": list(n), "nt": nt, "check_freq": args.check_freq, "dtype": args.dtype, "fast_f32": args.fast_f32, "fused": args.fused,
           "device_GB": ws.device_bytes() / 1e9}
    es = np.dtype(T).itemsize
    bytes_fwd = {"cd": 4, "vd": 9, "ela": 11}[args.kind] * es
    bytes_adj = {"cd": 9, "vd": 17, "ela": 27}[args.kind] * es
    obs = np.zeros((nt, 2, args.nrec) if args.kind == "ela" else (nt, args.nrec), dtype=T, order="F")
    for rep in range(args.reps + 1):
        if rep == 1:
            ws.kernel_timing(1)
        cu0 = ws.cell_updates()
        t0 = time.perf_counter()
        if grad:
            S.swgradient(ws, matprop, [shot()], [S.L2Misfit(observed=obs)])
        else:
            S.swforward(ws, matprop, [shot()])
        wall = time.perf_counter() - t0
        out["wall_s_last"] = wall
        out["Gcell_per_s_wall_last"] = (ws.cell_updates() - cu0) / wall / 1e9
    ws.kernel_timing(0)
    (f_ms, f_n), (a_ms, a_n), (_, r_n) = ws.kernel_timing_class(0), ws.kernel_timing_class(1), ws.kernel_timing_class(2)
    if f_n:
        d = f_ms / f_n * 1e-3
        out["fwd"] = {"us": d * 1e6, "GBps": bytes_fwd * ncell / d / 1e9, "Gcell_s": ncell / d / 1e9, "launches": f_n}
        if a_n:
            da = (a_ms - r_n * (f_ms / f_n)) / a_n * 1e-3
            out["adj"] = {"us": da * 1e6, "GBps": bytes_adj * ncell / da / 1e9, "Gcell_s": ncell / da / 1e9, "launches": a_n, "refwd_launches": r_n}
    ws.close()
    return out


if __name__ == "__main__":
    main()
