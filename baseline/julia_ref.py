"""Bridge to the real reference (SeismicWaves.jl, `parall = :threads`) for machines that have Julia -- the build image and the GPU box
do not, so nothing here runs in the driver's rounds; bench.py probes `available()` and falls back to the OpenMP oracle.

    python baseline/julia_ref.py pin      # run the seeded parity cases through Julia and compare the CPU oracle with them

write_problem() serialises one acoustic shot (the dict format of tests/cases.py) for baseline/run_reference.jl; run() times or
dumps it.  `pin` is how the oracle's "parity unpinned" status gets lifted: it needs `julia` plus SeismicWaves.jl v0.9.0 either in
the active Julia project or under baseline/_ref/SeismicWaves.jl."""
from __future__ import annotations

import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SCRIPT = os.path.join(HERE, "run_reference.jl")


def available() -> str | None:
    """path of the julia executable if the reference can be run, else None"""
    exe = shutil.which("julia")
    if exe is None:
        return None
    if os.path.isdir(os.path.join(HERE, "_ref", "SeismicWaves.jl")) or os.environ.get("JULIA_PROJECT") or os.environ.get("JULIA_LOAD_PATH"):
        return exe
    return None


def write_problem(path: str, case: dict, shot: int = 0, observed: np.ndarray | None = None, check_freq: int = 1, mute_src: int = 0, mute_rec: int = 0) -> None:
    T = case["dtype"].type
    sh = case["shots"][shot]
    n = case["n"]
    nsrc, nrec = sh["src_positions"].shape[0], sh["rec_positions"].shape[0]
    meta = dict(kind=case["kind"], T="Float32" if case["dtype"] == np.float32 else "Float64", ndim=len(n), h=repr(float(case["h"])), dt=repr(float(T(case["dt"]))),
                nt=case["nt"], halo=case["halo"], rcoef=repr(float(case["rcoef"])), freetop=int(case["freetop"]), check_freq=check_freq, mute_src=mute_src,
                mute_rec=mute_rec, domfreq=repr(float(sh["domfreq"])), nsrc=nsrc, nrec=nrec)
    for d, nd in enumerate(n):
        meta[f"n{d + 1}"] = nd
    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, "meta.txt"), "w") as f:
        for k, v in meta.items():
            f.write(f"{k} {v}\n")

    def dump(name, a):
        np.asfortranarray(np.asarray(a, dtype=T)).ravel(order="F").tofile(os.path.join(path, name))

    dump("vp.bin", case["vp"])
    if case["kind"] == "acoustic_vd":
        dump("rho.bin", case["rho"])
    dump("srcpos.bin", sh["src_positions"])
    dump("recpos.bin", sh["rec_positions"])
    dump("srctf.bin", sh["src_tf"])
    dump("observed.bin", observed if observed is not None else np.zeros((case["nt"], nrec)))


def run(path: str, mode: str = "time", warmup: int = 1, steps: int = 2, threads: str = "auto") -> dict:
    exe = available()
    if exe is None:
        raise RuntimeError("julia / SeismicWaves.jl not found")
    out = subprocess.run([exe, f"--threads={threads}", SCRIPT, path, mode, str(warmup), str(steps)], check=True, capture_output=True, text=True).stdout
    return json.loads(out.strip().splitlines()[-1])


def read_dump(path: str, case: dict, nrec: int) -> dict:
    T = case["dtype"].type
    out = {"seis": np.fromfile(os.path.join(path, "seis.bin"), dtype=T).reshape((case["nt"], nrec), order="F"),
           "misfit": float(open(os.path.join(path, "misfit.txt")).read())}
    for k in ("vp", "rho"):
        p = os.path.join(path, f"grad_{k}.bin")
        if os.path.exists(p):
            out[k] = np.fromfile(p, dtype=T).reshape(case["n"], order="F")
    return out


def pin() -> int:
    """oracle vs the real reference on the seeded parity cases; exit code 0 iff every rel-L2 error is within north_star's tolerance"""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases

    if available() is None:
        print("julia / SeismicWaves.jl not found: the oracle stays 'parity unpinned' (DESIGN.md 2)")
        return 2
    worst = 0.0
    for kind, n, dtype, cf in [("acoustic_cd", (64, 56), np.float64, 7), ("acoustic_cd", (61, 53), np.float32, 1), ("acoustic_cd", (30, 26, 28), np.float64, 6),
                               ("acoustic_vd", (64, 56), np.float64, 7), ("acoustic_vd", (67, 59), np.float32, 9)]:
        case = cases.acoustic_case(kind=kind, n=n, nt=120, halo=6, dtype=dtype, seed=3, nshots=1)
        syn, _ = cases.oracle_forward(case)
        obs = cases.make_observed(case, syn)
        (g, mis), seis, _ = cases.oracle_gradient(case, obs, check_freq=cf, mute_src=3, mute_rec=2)
        with tempfile.TemporaryDirectory() as d:
            write_problem(d, case, observed=obs[0], check_freq=cf, mute_src=3, mute_rec=2)
            run(d, "dump", 0, 1)
            ref = read_dump(d, case, obs[0].shape[1])
        errs = {"seis": cases.rel_l2(seis[0], ref["seis"]), "misfit": abs(float(mis) - ref["misfit"]) / abs(ref["misfit"])}
        errs.update({k: cases.rel_l2(g[k], ref[k]) for k in g})
        worst = max(worst, max(errs.values()) / cases.tol(dtype))
        print(kind, n, np.dtype(dtype).name, "check_freq", cf, {k: f"{v:.2e}" for k, v in errs.items()})
    print("oracle pinned against the reference" if worst <= 1 else "ORACLE DISAGREES WITH THE REFERENCE")
    return 0 if worst <= 1 else 1


if __name__ == "__main__":
    sys.exit(pin() if sys.argv[1:2] == ["pin"] else 2)
