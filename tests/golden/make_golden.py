#!/usr/bin/env python
"""Regenerates tests/golden/oracle_v1.npz: seismograms, misfits and gradients of the CPU oracle on small seeded cases.

The reference (SeismicWaves.jl, 100 % Julia) cannot run in this image and ships no golden vectors, so these
fixtures do NOT come from the reference: they freeze the oracle -- the sole arbiter of absolute values on heterogeneous and
free-surface models -- at the state that passed the reference's known-answer tests (tests/test_oracle_known_answers.py,
tests/test_oracle_elastic.py), so that a later edit of oracle/ cannot drift silently.  Run from the repository root:

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import cases as AC  # noqa: E402
import elastic_cases as EC  # noqa: E402

ACOUSTIC = [("acoustic_cd", (48, 40), np.float64, True), ("acoustic_cd", (22, 20, 24), np.float32, False), ("acoustic_vd", (52, 44), np.float64, True),
            ("acoustic_vd", (45, 41), np.float32, False)]
ELASTIC = [("momten", np.float64, True), ("extforce", np.float32, False)]


def compute():
    out = {}
    for k, (kind, n, dtype, freetop) in enumerate(ACOUSTIC):
        case = AC.acoustic_case(kind=kind, n=n, nt=60, halo=6, freetop=freetop, dtype=dtype, seed=100 + k, nshots=1, nsrc=2, nrec=4)
        for sh in case["shots"]:
            sh["src_positions"][:, -1] = 9.0 * case["h"]
        seis, _ = AC.oracle_forward(case)
        obs = AC.make_observed(case, seis)
        (grad, mis), _, _ = AC.oracle_gradient(case, obs, check_freq=7, mute_src=2, mute_rec=1)
        out[f"acou{k}_seis"] = np.asarray(seis[0])
        out[f"acou{k}_misfit"] = np.float64(mis)
        for name, g in grad.items():
            out[f"acou{k}_grad_{name}"] = np.asarray(g)
    for k, (kind, dtype, freetop) in enumerate(ELASTIC):
        case = EC.elastic_case(n=(56, 44), nt=60, halo=6, freetop=freetop, dtype=dtype, kind=kind, nshots=1, nsrc=1, nrec=3, seed=200 + k)
        for sh in case["shots"]:
            sh["src_positions"][:, 1] = 11.3 * case["h"]
        seis, _ = EC.oracle_forward(case)
        obs = EC.make_observed(case, seis)
        (grad, mis), _ = EC.oracle_gradient(case, obs, check_freq=7, mute_src=2, mute_rec=1)
        out[f"ela{k}_seis"] = np.asarray(seis[0])
        out[f"ela{k}_misfit"] = np.float64(mis)
        for name, g in grad.items():
            out[f"ela{k}_grad_{name}"] = np.asarray(g)
    return out


if __name__ == "__main__":
    data = compute()
    np.savez_compressed(os.path.join(HERE, "oracle_v1.npz"), **data)
    print({k: (v.shape, str(v.dtype)) for k, v in data.items()})
