#!/usr/bin/env python
"""Generates tests/golden/c2_fullsize_oracle.npz: the CPU oracle's result for ONE shot of the benchmarked workload itself
(bench.c2_problem(): 2D acoustic VD 4096^2 Float32, nt = 1000, check_freq = 31, shot 32 of 64, 512 receivers, zero observed data,
mute radius 3 around the source) -- misfit, seismograms (every 4th receiver), vp / rho gradients (every 16th cell per axis, plus
their full L2 norms and sums).  bench.py re-computes this shot on the GPU after the timed region and compares; the `-m gpu`
suite compares both arithmetic modes.  The oracle's outputs, not the reference's (Julia is not installed): see DESIGN.md 2.

    python tests/golden/make_c2_fullsize.py          # ~4 min on 8 cores, ~9 GB of host memory
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SHOT, REC_STRIDE, CELL_STRIDE = 32, 4, 16


def main():
    import cases
    import twins
    from oracle import oracle as O

    O.build()
    O.use_openmp(True)
    case = twins.c2_twin(n=4096, nt=1000, nrec=512, shot_index=SHOT, f0=8.0)
    observed = [np.zeros((1000, 512), dtype=np.float32, order="F")]
    t0 = time.time()
    (g, mis), seis, _ = cases.oracle_gradient(case, observed, check_freq=case["check_freq"], mute_src=3)
    secs = time.time() - t0
    out = dict(shot=SHOT, rec_stride=REC_STRIDE, cell_stride=CELL_STRIDE, misfit=np.float64(mis), oracle_seconds=secs, cores=os.cpu_count(),
               seis=np.ascontiguousarray(seis[0][:, ::REC_STRIDE]), seis_norm=np.linalg.norm(seis[0].astype(np.float64)))
    for k in ("vp", "rho"):
        a = g[k].astype(np.float64)
        out[f"grad_{k}"] = np.ascontiguousarray(g[k][::CELL_STRIDE, ::CELL_STRIDE])
        out[f"grad_{k}_norm"] = np.linalg.norm(a)
        out[f"grad_{k}_sum"] = a.sum()
    np.savez_compressed(os.path.join(HERE, "c2_fullsize_oracle.npz"), **out)
    print(f"oracle: {secs:.1f} s, misfit {float(mis):.9e}, |g_vp| {out['grad_vp_norm']:.6e}, |g_rho| {out['grad_rho_norm']:.6e}")


if __name__ == "__main__":
    main()
