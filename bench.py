#!/usr/bin/env python
"""bench.py -- forward + adjoint-gradient throughput of the finite-difference hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W                      # this repo's B200 engine
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   (N > 1, one rank per GPU)
    python bench.py --impl reference ...                               # CPU restatement of the reference's Threads backend

Workload (config C2 of BASELINE.md / SURVEY.md 8d): 2D acoustic variable-density, 4096 x 4096 cells, Float32,
nt = 1000, C-PML halo 20, free surface, check_freq = 31, one source and 512 receivers per shot, L2 misfit against
zero observed data.  A *step* is the complete gradient of ONE shot on every rank (forward sweep with checkpoints,
adjoint sweep with checkpoint re-forwarding and in-kernel zero-lag correlation, gradient post-processing and
accumulation); shots are independent, so rank r works on its own shots (weak scaling: per-GPU work is fixed) and
the per-rank total gradients are summed by one NCCL all-reduce at the end of the timed region.

metric = cell-updates per second (one cell advanced one time step by one sweep: forward, re-forward or adjoint;
SURVEY 8d), whole job.  `value` is timed with CUDA events on the engine's stream with the material fields resident
in HBM; `e2e` times the public API call swgradient(wavesim, matprop, shots, misfit) on host buffers (material
upload, per-shot binding, seismogram and gradient downloads inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Gcell-updates/s (fwd+adjoint)"
UNIT = "Gcell-updates/s"


# ---------------------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------------------
def c2_problem(n=4096, nt=1000, nshots_total=64, nrec=512, seed=1234):
    h = 10.0
    rng = np.random.default_rng(seed)
    depth = np.arange(n, dtype=np.float64) / (n - 1)
    vp = 2000.0 + 2500.0 * depth[None, :] + rng.normal(0.0, 50.0, size=(n, n))
    vp = np.clip(vp, 1500.0, 4700.0)
    rho = 310.0 * vp**0.25
    vp32 = np.asfortranarray(vp.astype(np.float32))
    rho32 = np.asfortranarray(rho.astype(np.float32))
    dt = 0.99 * (6.0 / 7.0) * h / (float(vp32.max()) * math.sqrt(2.0))
    f0 = 8.0
    t = np.arange(nt) * dt
    lo, hi = 64.0 * n / 4096.0, 4032.0 * n / 4096.0
    xs = h * np.round(np.linspace(lo, hi, nshots_total))
    xr = np.linspace(0.05, 0.95, nrec) * (n - 1) * h
    return dict(n=n, nt=nt, h=h, dt=dt, f0=f0, vp=vp32, rho=rho32, t=t, xs=xs, xr=xr, halo=20, check_freq=max(2, int(math.isqrt(nt))),
                nshots_total=nshots_total)


def n_refwd(nt: int, cf: int) -> int:
    """re-forwarded steps of the LinearCheckpointer schedule (src/utils/checkpointers.jl:87-105, acou_gradient.jl:141-176)."""
    if cf == 1:
        return 0
    last = (nt // cf) * cf
    curr, cnt = last, 0
    for it in range(nt, 0, -1):
        if not (curr <= it - 1 <= curr + cf or (it - 1) % cf == 0):
            curr -= cf
            cnt += cf - 1
    return cnt


# ---------------------------------------------------------------------------------------------------------
# clocks sampling during the timed region (B200_PROFILING.md)
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / throttle-reason samples during the timed region, taken in-process through NVML (nvidia_ml_py).
    A separate `nvidia-smi -lms` process perturbed the launch stream of the benchmark itself (driver lock contention:
    the kernels here are ~100 us each and launched back to back), so the sampling is a light NVML poll every 250 ms."""

    def __init__(self, device_index: int, period_s: float = 0.25):
        self.idx, self.period = device_index, period_s
        self.samples, self._stop, self.thr, self.err = [], threading.Event(), None, None

    def start(self):
        try:
            import pynvml as N

            N.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.idx]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else self.idx
            self.h = N.nvmlDeviceGetHandleByIndex(phys)
            self.N = N
        except Exception as e:  # noqa: BLE001
            self.err = f"NVML unavailable: {e}"
            return
        self.thr = threading.Thread(target=self._run, daemon=True)
        self.thr.start()

    def _run(self):
        N = self.N
        while not self._stop.is_set():
            try:
                sm = N.nvmlDeviceGetClockInfo(self.h, N.NVML_CLOCK_SM)
                smax = N.nvmlDeviceGetMaxClockInfo(self.h, N.NVML_CLOCK_SM)
                reasons = N.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(N, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else N.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                power = N.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.samples.append((time.perf_counter(), sm, smax, power, int(reasons)))
            except Exception as e:  # noqa: BLE001
                self.err = str(e)
                return
            self._stop.wait(self.period)

    def stop(self, t_from=None, t_to=None):
        self._stop.set()
        if self.thr is not None:
            self.thr.join(timeout=2)
        sel = [x for x in self.samples if (t_from is None or x[0] >= t_from) and (t_to is None or x[0] <= t_to)]
        if not sel:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.err or "no samples"]}
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "hw_power_brake_slowdown": 0x80}
        reasons = sorted(k for k, b in bits.items() if any(x[4] & b for x in sel))
        return {"sm_mhz": float(np.median([x[1] for x in sel])), "sm_max_mhz": float(max(x[2] for x in sel)), "power_w_max": float(max(x[3] for x in sel)),
                "samples": len(sel), "reasons": reasons, "how": "NVML poll every 250 ms during the timed region"}


# ---------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle's OpenMP build, which keeps the reference's Threads-backend sweep structure
# (separate p / inject / v / record / correlate passes).  Julia is not installed, so the
# reference itself cannot run (kind = "port").
# ---------------------------------------------------------------------------------------------------------
def cpu_sample_problem(n=4096, nt=40, cf=6):
    """the C2 problem itself (grid, model, halo, free surface, 512 receivers) with the time axis shortened so that one shot's
    gradient costs the host a few seconds (the driver times 25 of them): nt = 40, check_freq = isqrt(40) = 6"""
    prob = c2_problem(n=n, nt=nt, nshots_total=64, nrec=512)
    prob["check_freq"] = cf
    return prob


CPU_SAMPLE = ("per step: one shot's gradient of the C2 workload itself (4096x4096 Float32 model, halo 20, free surface, 1 source, 512 receivers, "
              "L2 misfit vs zeros) with the time axis shortened to nt=40, check_freq=6")


def cpu_threads() -> int:
    """All host cores for the OpenMP oracle.  torch.distributed.run exports OMP_NUM_THREADS=1 to its workers and libgomp reads the
    variable only once, so the count is set through the library (oracle.set_threads -> omp_set_num_threads)."""
    from oracle import oracle as O

    O.build()
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = str(cores)
    return O.set_threads(cores)


def cpu_gradient_once(prob, shot_index=0):
    """one shot's gradient with the oracle (OpenMP build); returns (seconds, cell_updates)."""
    from oracle import oracle as O

    O.build()
    O.use_openmp(True)
    T = np.float32
    n, nt, h = prob["n"], prob["nt"], prob["h"]
    params = O.Params(nt=nt, dt=prob["dt"], gridsize=(n, n), spacing=(h, h), halo=prob["halo"], rcoef=1e-4, freetop=True, dtype=T)
    sim = O.build_wavesim("acoustic_vd", params, gradient=True, check_freq=prob["check_freq"])
    t0s = 2.0 / prob["f0"]
    tf = np.asfortranarray((1000.0 * O.gaussderivstf(prob["t"], t0s, prob["f0"])).astype(T).reshape(nt, 1))
    sp = np.asfortranarray(np.array([[prob["xs"][shot_index % len(prob["xs"])], 2 * h]], dtype=T))
    rp = np.asfortranarray(np.stack([prob["xr"], np.full_like(prob["xr"], 3 * h)], axis=1).astype(T))
    shot = O.ScalarShot(src_positions=sp, src_tf=tf, domfreq=T(prob["f0"]), rec_positions=rp)
    mis = O.L2Misfit(observed=np.zeros((nt, rp.shape[0]), dtype=T, order="F"))
    sim.set_matprop(prob["vp"], prob["rho"])
    c0 = sim.cell_updates
    t0 = time.perf_counter()
    sim.init_shot(shot)
    sim.gradient_1shot(shot, mis)
    dt = time.perf_counter() - t0
    O.use_openmp(False)
    return dt, sim.cell_updates - c0


def cpu_leg(warmup: int, steps: int):
    """(Gcell-updates/s, seconds per step, threads, kind, sample): mean over `steps` gradients after `warmup` untimed ones.  The real
    reference (Julia, parall = :threads) is timed when it is installed (baseline/julia_ref.py); otherwise the OpenMP oracle."""
    prob = cpu_sample_problem()
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    try:
        import julia_ref

        if julia_ref.available():
            import tempfile

            case = dict(kind="acoustic_vd", n=(prob["n"], prob["n"]), nt=prob["nt"], dt=prob["dt"], h=prob["h"], halo=prob["halo"], rcoef=1e-4, freetop=True,
                        dtype=np.dtype(np.float32), vp=prob["vp"], rho=prob["rho"],
                        shots=[dict(src_positions=np.array([[prob["xs"][0], 2 * prob["h"]]]), rec_positions=np.stack([prob["xr"], np.full_like(prob["xr"], 3 * prob["h"])], axis=1),
                                    src_tf=(1000.0 * _gaussderiv(prob["t"], 2.0 / prob["f0"], prob["f0"])).reshape(-1, 1), domfreq=prob["f0"])])
            with tempfile.TemporaryDirectory() as d:
                julia_ref.write_problem(d, case, check_freq=prob["check_freq"], mute_src=3)
                r = julia_ref.run(d, "time", warmup, steps)
            units = prob["n"] ** 2 * (2 * prob["nt"] + n_refwd(prob["nt"], prob["check_freq"]))
            return units / r["seconds_per_step"] / 1e9, r["seconds_per_step"], int(r["threads"]), "reference", CPU_SAMPLE + " (SeismicWaves.jl, parall=:threads)"
    except Exception as e:  # noqa: BLE001 -- fall back to the port, say why
        print(f"[bench] Julia reference not usable ({e}); timing the OpenMP oracle", file=sys.stderr)
    threads = cpu_threads()
    for w in range(warmup):
        cpu_gradient_once(prob, w)
    tot_t, tot_u = 0.0, 0
    for k in range(steps):
        dt, u = cpu_gradient_once(prob, warmup + k)
        tot_t += dt
        tot_u += u
    return tot_u / tot_t / 1e9, tot_t / max(steps, 1), threads, "port", CPU_SAMPLE + " (OpenMP build of the CPU restatement of the reference's Threads backend; Julia is not installed)"


def _gaussderiv(t, t0, f0):
    """gaussderivstf (src/utils/utils.jl:11-15)"""
    return (t - t0) * np.exp(-((np.pi * f0 * (t - t0)) ** 2))


def run_reference(args):
    """--impl reference: the reference's CPU path on the host cores (rank 0 only): SeismicWaves.jl's Threads backend when Julia is
    installed, else the OpenMP build of the oracle."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    val, sec, threads, kind, sample = cpu_leg(args.warmup, args.steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sec, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(None, args, note="the CPU arm runs this workload with the time axis shortened (nt=40, check_freq=6): same grid, model, "
                                                                       "boundaries, sources and receivers; the metric is a rate (cell-updates/s), so the step count does not enter it"),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(_, args, note=None):
    c = {"workload": "C2: 2D acoustic variable-density 4096x4096 Float32, nt=1000, C-PML halo 20 + free surface, check_freq=31, "
                     "1 source + 512 receivers per shot, full adjoint gradient (vp, rho) with checkpoint re-forwarding",
         "shots_per_gpu_per_step": 1, "parallelism": f"shots sharded over {args.gpus} GPU(s), NCCL all-reduce of the gradients",
         "l2_policy": "working set per sweep (>= 600 MB) exceeds the 126 MB L2, no flush needed"}
    if args.grid != 4096 or args.nt != 1000:
        c["workload"] = f"REDUCED (not the headline config): 2D acoustic VD {args.grid}x{args.grid} Float32 nt={args.nt}"
    elif args.check_freq and args.check_freq != 31:
        c["workload"] = c["workload"].replace("check_freq=31", f"check_freq={args.check_freq} (NOT the headline check_freq=31)")
    if note:
        c["note"] = note
    return c


# ---------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------
class C2Bench:
    """The C2 workload bound to this rank's GPU: one wavesim per arithmetic mode, the shot factory, the timed loops."""

    def __init__(self, args, S, torch, dist, world, rank, local):
        self.args, self.S, self.torch, self.dist = args, S, torch, dist
        self.world, self.rank, self.local = world, rank, local
        self.lib = S._lib.load()
        prob = c2_problem(n=args.grid, nt=args.nt)
        if args.check_freq:
            prob["check_freq"] = args.check_freq
        self.prob = prob
        n, nt, h = prob["n"], prob["nt"], prob["h"]
        T = np.float32
        bc = S.CPMLBoundaryConditionParameters(halo=prob["halo"], rcoef=T(1e-4), freeboundtop=True)
        self.params = S.InputParametersAcoustic(nt, T(prob["dt"]), (n, n), (T(h), T(h)), bc, dtype=np.dtype(T))
        self.gradparams = S.GradParameters(mute_radius_src=3, mute_radius_rec=0, compute_misfit=True, check_freq=prob["check_freq"])

        def pinned(a):
            t = torch.empty(a.shape[::-1], dtype=torch.float32, pin_memory=True)  # reversed shape: C-order tensor == F-order array
            v = t.numpy().T
            v[...] = a
            return v, t

        vp, self._k1 = pinned(prob["vp"])
        rho, self._k2 = pinned(prob["rho"])
        self.matprop = S.VpRhoAcousticVDMaterialProperties(vp, rho)
        self.tf = np.asfortranarray((1000.0 * S.gaussderivstf(prob["t"], 2.0 / prob["f0"], prob["f0"])).astype(T).reshape(nt, 1))
        self.rp = np.stack([prob["xr"], np.full_like(prob["xr"], 3 * h)], axis=1).astype(T)
        self.nrec = self.rp.shape[0]
        self.zeros_obs = np.zeros((nt, self.nrec), dtype=T, order="F")
        self.comm = None
        if world > 1:
            idbuf = torch.zeros(128, dtype=torch.uint8)
            if rank == 0:
                raw = (S._lib.C.c_ubyte * 128)()
                S._lib.check(self.lib.swb_comm_unique_id(raw))
                idbuf = torch.tensor(list(raw), dtype=torch.uint8)
            idbuf = idbuf.cuda()
            dist.broadcast(idbuf, 0)
            raw = (S._lib.C.c_ubyte * 128)(*idbuf.cpu().tolist())
            self.comm = S._lib.C.c_void_p()
            S._lib.check(self.lib.swb_comm_create(raw, world, rank, local, S._lib.C.byref(self.comm)))

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def shot(self, g):
        S, T, prob = self.S, np.float32, self.prob
        sp = np.array([[prob["xs"][g % prob["nshots_total"]], 2 * prob["h"]]], dtype=T)
        return S.ScalarShot(srcs=S.ScalarSources(sp, self.tf, T(prob["f0"])), recs=S.ScalarReceivers(self.rp, prob["nt"], dtype=np.dtype(T)))

    def wavesim(self, fast_f32):
        S = self.S
        rp = S.RunParameters(parall="B200", device=self.local, erroronPPW=False, fast_f32=bool(fast_f32))
        ws = S.build_wavesim(self.params, self.matprop, runparams=rp, gradparams=self.gradparams, gradient=True)
        ws.check_sim_consistency(self.matprop, [self.shot(0)])
        ws.set_wavesim_matprop(self.matprop)
        ws.zero_total_gradient()
        return ws

    def one_step(self, ws, g, misfit_val):
        """one shot's gradient with everything resident on the device (the `value` arm)"""
        S, lib, gp = self.S, self.lib, self.gradparams
        shot = self.shot(g)
        ws.init_shot(shot)  # C-PML profiles (host, O(halo)) + upload
        ws._bind(shot)  # nearest grid points, STF scaling (host, O(nt)) + upload of a few KB
        S._lib.check(lib.swb_sim_gradient_l2(ws._h, None, None, S._lib.C.byref(misfit_val)))
        sp, rpp = shot.srcs.positions, shot.recs.positions
        S._lib.check(lib.swb_sim_accumulate_gradient(ws._h, sp.shape[0], S.api._vp(np.asfortranarray(sp)), gp.mute_radius_src,
                                                     rpp.shape[0], S.api._vp(np.asfortranarray(rpp)), gp.mute_radius_rec))

    def timed(self, ws, shot_ids, warm_ids, sampler=None):
        """device-timed loop over this rank's shots `shot_ids` (+ the closing all-reduce); returns a dict of raw measurements
        (ms = max over ranks, cell-updates = sum over ranks)"""
        S, lib, torch, dist = self.S, self.lib, self.torch, self.dist
        stream_ptr = S._lib.C.c_void_p()
        S._lib.check(lib.swb_sim_stream(ws._h, S._lib.C.byref(stream_ptr)))
        ext = torch.cuda.ExternalStream(stream_ptr.value, device=torch.device("cuda", self.local))
        misfit_val = S._lib.C.c_double()
        for g in warm_ids:
            self.one_step(ws, g, misfit_val)
        if self.comm is not None:  # warm-up of the collective too (NCCL sets up its channels on the first call)
            S._lib.check(lib.swb_sim_allreduce_total_gradient(ws._h, self.comm))
        ws.zero_total_gradient()
        ws.kernel_timing(1)
        self.barrier()
        t_from = time.perf_counter()
        cu0, l0 = ws.cell_updates(), lib.swb_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        for g in shot_ids:
            self.one_step(ws, g, misfit_val)
        if self.comm is not None:
            S._lib.check(lib.swb_sim_allreduce_total_gradient(ws._h, self.comm))
        e1.record(ext)
        self.barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop(t_from, time.perf_counter()) if sampler is not None else None
        cu = ws.cell_updates() - cu0
        launches = lib.swb_launch_count() - l0
        ws.kernel_timing(0)
        kt = [ws.kernel_timing_class(c) for c in range(3)]
        tms = torch.tensor([ms], dtype=torch.float64, device="cuda")
        tcu = torch.tensor([float(cu)], dtype=torch.float64, device="cuda")
        if self.world > 1:
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            dist.all_reduce(tcu, op=dist.ReduceOp.SUM)
        ms_max, cu_all = float(tms.item()), float(tcu.item())
        return dict(ms=ms_max, cell_updates=cu_all, value=cu_all / (ms_max * 1e-3) / 1e9, launches=int(launches), kt=kt, clocks=clocks,
                    misfit_last=float(misfit_val.value), nshots_rank=len(shot_ids))

    def roofline(self, kt, peak, peak_src, fast):
        """dominant kernel = the fused VD step (forward / re-forward sweeps): 9 arrays x 4 B per cell; the adjoint launch
        (adjoint step + injection + three correlations, 17 arrays x 4 B) is reported alongside"""
        (kt_ms, kt_n), (ka_ms, ka_n), (_, kr_n) = kt
        n = self.prob["n"]
        if kt_n <= 0:
            return None
        dur = kt_ms / kt_n * 1e-3
        ach = 36 * n * n / dur / 1e9
        full = self.args.grid == 4096
        roof = {"bound": "hbm", "kernel": "vd_fused_kernel<float,%s,0,16> (v update + p update + inject + record, one launch per time step)" % ("float" if fast else "double"),
                "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": 582.0e6 if (full and fast) else None,
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one forward launch, ncu --set full, cold caches (profiles/r2_ncu_vd_fwd.txt)",
                "peak_source": peak_src, "frac_of_8TBs_nominal": ach / 8000.0, "avg_launch_us": dur * 1e6, "timed_launches": kt_n,
                "algorithmic_bytes_per_launch": 36 * n * n,
                "sampling": "CUDA events on the engine's stream around each forward-sweep graph (nt step launches + the checkpoint "
                            "copies) inside the timed region; duration = elapsed / nt"}
        if ka_n > 0:  # the adjoint graph also holds the re-forward launches (class 2 count); they are forward-step launches
            dur_a = (ka_ms - kr_n * (kt_ms / kt_n)) / ka_n * 1e-3
            ach_a = 68 * n * n / dur_a / 1e9
            roof["adjoint_kernel"] = {"achieved": ach_a, "frac": ach_a / peak, "avg_launch_us": dur_a * 1e6, "timed_launches": ka_n,
                                      "algorithmic_bytes_per_launch": 68 * n * n, "traffic": 1112.5e6 if (full and fast) else None,
                                      "traffic_source": "profiles/r2_ncu_vd_adj.txt"}
        return roof

    def e2e(self, ws, nshots):
        """the public API call on host buffers: swgradient(wavesim, matprop, shots, misfit), wall clock, max over ranks"""
        S, torch, dist, args = self.S, self.torch, self.dist, self.args
        n, nt, nrec = self.prob["n"], self.prob["nt"], self.nrec
        shots = [self.shot(self.rank * (args.steps + args.warmup) + args.warmup + k) for k in range(nshots)]
        misfits = [S.L2Misfit(observed=self.zeros_obs) for _ in shots]
        cu0 = ws.cell_updates()
        self.barrier()
        t0 = time.perf_counter()
        grad, mis = S.swgradient(ws, self.matprop, shots, misfits)
        if self.comm is not None:
            S._lib.check(self.lib.swb_sim_allreduce_total_gradient(ws._h, self.comm))
            grad = ws.get_total_gradient()
        self.barrier()
        t_e2e = time.perf_counter() - t0
        cu_e = ws.cell_updates() - cu0
        te = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
        tce = torch.tensor([float(cu_e)], dtype=torch.float64, device="cuda")
        if self.world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
            dist.all_reduce(tce, op=dist.ReduceOp.SUM)
        field_b = n * n * 4
        h2d = (2 * field_b) / nshots + (nt * 4 + 2 * 8 + nrec * 2 * 8 + nt * nrec * 4)  # material (once per call) + per shot: STF, positions, observed
        d2h = nt * nrec * 4 + (2 * field_b * (2 if self.comm is not None else 1)) / nshots  # seismograms per shot + gradient (vp, rho) per call
        assert np.isfinite(grad["vp"]).all(), "gradient is not finite"
        if self.args.grid == 4096 and self.args.nt == 1000:
            assert float(np.abs(grad["vp"]).max()) > 0, "gradient is empty"
        return {"value": float(tce.item()) / float(te.item()) / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "seconds": float(te.item())}

    def parity_check(self, ws, fast):
        """Re-computes shot 32 of the workload through the public API and compares it with the CPU oracle's result for the same shot
        (tests/golden/c2_fullsize_oracle.npz, generator tests/golden/make_c2_fullsize.py): seismograms (every 4th receiver), misfit,
        vp / rho gradients (every 16th cell per axis + full norms).  Raises if any relative-L2 error exceeds north_star's 1e-4."""
        S = self.S
        path = os.path.join(ROOT, "tests", "golden", "c2_fullsize_oracle.npz")
        if not (self.args.grid == 4096 and self.args.nt == 1000 and self.prob["check_freq"] == 31 and os.path.exists(path)):
            return None
        gold = np.load(path)
        shot = self.shot(int(gold["shot"]))
        grad, mis = S.swgradient(ws, self.matprop, [shot], [S.L2Misfit(observed=self.zeros_obs)])
        rs, cs = int(gold["rec_stride"]), int(gold["cell_stride"])

        def rel(a, b):
            a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
            return float(np.linalg.norm(a - b) / np.linalg.norm(b))

        out = {"shot": int(gold["shot"]), "seismograms_rel_l2": rel(shot.recs.seismograms[:, ::rs], gold["seis"]),
               "misfit_rel": abs(float(mis) - float(gold["misfit"])) / float(gold["misfit"])}
        for k in ("vp", "rho"):
            out[f"grad_{k}_rel_l2"] = rel(grad[k][::cs, ::cs], gold[f"grad_{k}"])
            out[f"grad_{k}_norm_rel"] = abs(float(np.linalg.norm(grad[k].astype(np.float64))) - float(gold[f"grad_{k}_norm"])) / float(gold[f"grad_{k}_norm"])
        out["tolerance"] = 1e-4
        out["against"] = "CPU oracle, tests/golden/c2_fullsize_oracle.npz (the same shot, full size; the oracle's output, not the reference's)"
        out["arith"] = "fast_f32" if fast else "f64 intermediates"
        bad = {k: v for k, v in out.items() if isinstance(v, float) and k != "tolerance" and not (v <= 1e-4)}
        assert not bad, f"benchmarked path disagrees with the oracle at full size: {bad}"
        return out

    def close(self):
        if self.comm is not None:
            self.lib.swb_comm_destroy(self.comm)
            self.comm = None


def run_b200(args):
    import torch
    import torch.distributed as dist

    import swb200 as S

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}"
    S._lib.require_device()  # no CPU fallback
    torch.cuda.set_device(local)
    import logging

    logging.getLogger("seismicwaves_b200").setLevel(logging.ERROR)  # the points-per-wavelength warning repeats for every shot
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = C2Bench(args, S, torch, dist, world, rank, local)
    n, nt = B.prob["n"], B.prob["nt"]
    headline = args.grid == 4096 and args.nt == 1000

    # ---- headline: device-resident arm (`value`) and the end-to-end arm (`e2e`), weak scaling: one shot per GPU and step ------------
    ws = B.wavesim(args.fast_f32)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler is not None:
        sampler.start()  # started before the warm-up so that its start-up cost is not inside the timed region
    base = rank * (args.steps + args.warmup)
    main = B.timed(ws, [base + args.warmup + k for k in range(args.steps)], [base + w for w in range(args.warmup)], sampler)
    e2e = B.e2e(ws, args.steps)
    parity = B.parity_check(ws, bool(args.fast_f32)) if rank == 0 else None
    if world > 1:
        B.barrier()

    # ---- strong scaling: the 64 shots of BASELINE's C2 sharded over the ranks (distribsrcs rule), gradients all-reduced ----------------
    strong = None
    if headline and not args.no_extras:
        from swb200.multigpu import shot_group

        mine = list(shot_group(B.prob["nshots_total"], world, rank))
        r = B.timed(ws, mine, [])
        strong = {"shots_total": B.prob["nshots_total"], "shots_this_rank": len(mine), "seconds": r["ms"] * 1e-3, "value": r["value"], "unit": UNIT,
                  "scaling": "strong", "note": "BASELINE config 2 as written: 64 shots sharded over the GPUs (contiguous groups, reference's distribsrcs rule), "
                                               "one NCCL all-reduce of the gradients at the end; device-timed, max over ranks"}
    dominant_bytes = ws.device_bytes()
    ws.close()

    # ---- the reference-faithful arithmetic (Float64 intermediates; bit-identical to the oracle) on the same workload ----------------
    faithful = None
    if not args.no_extras and args.fast_f32:
        ws2 = B.wavesim(0)
        r = B.timed(ws2, [base + 1 + k for k in range(3)], [base])
        faithful = {"arith": "f32 storage, f64 intermediates (the reference's promotion rule; bit-identical to the CPU oracle at full size, "
                             "profiles/r2_parity_benchmarked_modes.jsonl)", "value": r["value"], "unit": UNIT, "ms_per_step": r["ms"] / 3, "steps": 3, "warmup": 1,
                    "_kt": r["kt"]}
        if rank == 0:
            faithful["parity_check"] = B.parity_check(ws2, False)
        if world > 1:
            B.barrier()
        ws2.close()

    slab = slab_entry(S, torch, dist, world, rank, local) if (world > 1 and headline and not args.no_extras) else None

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md 6.65 TB/s)"
        if faithful is not None:
            faithful["roofline"] = B.roofline(faithful.pop("_kt"), peak, peak_src, False)
        line = {
            "metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": main["ms"] / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(None, args),
            "arith": "f32 storage, f32 arithmetic (SWB_FLAG_FAST_F32)" if args.fast_f32 else "f32 storage, f64 intermediates (the reference's promotion rule)",
            "clocks": main["clocks"], "e2e": e2e, "gpu_launches": main["launches"], "roofline": B.roofline(main["kt"], peak, peak_src, bool(args.fast_f32)),
            "misfit_last_shot": main["misfit_last"],
            "useful_Gcell_per_s": 2.0 * n * n * nt * args.steps * world / (main["ms"] * 1e-3) / 1e9,
            "device_bytes": dominant_bytes, "parity_check": parity, "reference_faithful_mode": faithful, "strong_scaling": strong,
        }
        if slab is not None:
            if "algorithmic_GBps_per_gpu" in slab:
                slab["frac_of_peak_16B_per_cell"] = slab["algorithmic_GBps_per_gpu"] / peak
            line["slab_c5"] = slab
        if world == 1 and not args.no_extras and headline:
            line["other_configs"] = other_configs(peak)
        if not args.no_cpu:
            val, sec, threads, kind, sample = cpu_leg(1, 2)
            line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample + "; mean of 2 after 1 warm-up", "seconds_per_step": sec}
        print(json.dumps(line), flush=True)
    B.close()
    if world > 1:
        dist.destroy_process_group()


def slab_entry(S, torch, dist, world, rank, local):
    """BASELINE config 5 under --gpus N > 1: 3D acoustic CD forward run of a 2048 x 2048 x (128 N) grid cut into N z slabs (one per GPU),
    halo exchange over NVLink every step; preceded by a bitwise check of the decomposed run against the single-GPU engine on a small
    twin (tools/slab_check.py).  Device-timed on the engines' streams, max over ranks."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import slab_check

    try:
        ok, detail = slab_check.bitwise_twin(S, torch, dist, world, rank, local, exchange="p2p")
        ok2, detail2 = slab_check.bitwise_twin(S, torch, dist, world, rank, local, twins=slab_check.TWINS[:1], exchange="nccl")
        res = slab_check.throughput(S, torch, dist, world, rank, local, grid=(2048, 2048, 128 * world), nt=100, exchange="p2p")
        alt = slab_check.throughput(S, torch, dist, world, rank, local, grid=(2048, 2048, 128 * world), nt=100, exchange="nccl")
        res["bitwise_equal_to_single_gpu_on_twin"] = bool(ok and ok2)
        res["twin"] = detail + detail2
        res["nccl_exchange_for_comparison"] = {k: alt[k] for k in ("ms_per_step", "value", "per_gpu_Gcell_per_s", "exchange")}
        return res
    except Exception as e:  # noqa: BLE001 -- the headline line must survive a failing extra
        return {"error": str(e)[:300]}


def other_configs(peak):
    """Per-launch timings of the other fused engines on the BASELINE.json configurations that are not the headline: C3 elastic at full
    size (both precisions), 2D constant density at roofline size, and C4 itself (3D CD 768^3, nt = 500, check_freq = 50: 155 GB of device
    checkpoints).  CUDA events over whole replayed sweeps (tools/bench_sim.py).  Informational: `value` above is C2 only."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_sim

    c3 = "C3 2D elastic P-SV 4096x2048, nt=1000, check_freq=31, off-grid moment-tensor source, 10 vector receivers, forward + rho/lambda/mu adjoint gradient"
    runs = [(c3 + " -- Float32 (f32 arithmetic)", ["--kind ela --n 4096 2048 --nt 1000 --check-freq 31 --dtype f32 --fast-f32 1 --nrec 10 --reps 1"]),
            (c3 + " -- Float32 (f64 intermediates, reference-faithful)", ["--kind ela --n 4096 2048 --nt 1000 --check-freq 31 --dtype f32 --fast-f32 0 --nrec 10 --reps 1"]),
            (c3 + " -- Float64", ["--kind ela --n 4096 2048 --nt 1000 --check-freq 31 --dtype f64 --nrec 10 --reps 1"]),
            ("2D acoustic CD 4096x4096 Float32 (C1 physics at roofline size)", ["--kind cd --n 4096 4096 --nt 100 --check-freq 10 --reps 2"]),
            ("C4 3D acoustic CD 768^3 Float32, forward + adjoint gradient, nt=500, check_freq=50, 1024 receivers",
             ["--kind cd --n 768 768 768 --nt 500 --check-freq 50 --nrec 1024 --reps 1", "--kind cd --n 768 768 768 --nt 60 --check-freq 10 --nrec 1024 --reps 1"])]
    out = []
    for name, variants in runs:
        rec = None
        for i, argv in enumerate(variants):
            try:
                r = bench_sim.measure(bench_sim.parse_args(argv.split()))
                rec = {"config": name if i == 0 else name + f" -- FALLBACK ARGS ({argv}) after: {rec['error']}", "device_GB": r["device_GB"],
                       "gradient_wall_s": r.get("wall_s_last"), "Gcell_per_s_wall": r.get("Gcell_per_s_wall_last")}
                for k in ("fwd", "adj"):
                    if k in r:
                        rec[k] = {"us_per_step": r[k]["us"], "Gcell_per_s": r[k]["Gcell_s"], "algorithmic_GBps": r[k]["GBps"], "frac_of_peak": r[k]["GBps"] / peak}
                break
            except Exception as e:  # noqa: BLE001 -- the headline line must survive a failing extra
                rec = {"config": name, "error": str(e)[:200]}
        out.append(rec)
    return out


def main():
    import faulthandler

    faulthandler.enable()  # a fatal signal in a native library leaves a traceback on stderr instead of a silent exit
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"  # NCCL_DEBUG=VERSION prints a banner on stdout, in front of the JSON line
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", type=int, default=4096, help="override the grid size (non-headline runs only)")
    ap.add_argument("--nt", type=int, default=1000, help="override the number of time steps (non-headline runs only)")
    ap.add_argument("--check-freq", type=int, default=0)
    ap.add_argument("--fast-f32", type=int, default=1,
                    help="1 (default) = Float32 storage and Float32 arithmetic (SWB_FLAG_FAST_F32; within the 1e-4 Float32 tolerance); "
                         "0 = Float64 intermediates, bit-faithful to the reference's promotion rule")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the per-launch timings of the non-headline configurations (other_configs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
