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
def cpu_sample_problem(n=1536, nt=60, cf=7):
    prob = c2_problem(n=n, nt=nt, nshots_total=64, nrec=128)
    prob["check_freq"] = cf
    return prob


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


def run_reference(args):
    """--impl reference: the CPU restatement of the reference's Threads backend on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    prob = cpu_sample_problem()
    for w in range(args.warmup):
        cpu_gradient_once(prob, w)
    tot_t, tot_u = 0.0, 0
    for k in range(args.steps):
        dt, u = cpu_gradient_once(prob, args.warmup + k)
        tot_t += dt
        tot_u += u
    val = tot_u / tot_t / 1e9
    sample = f"per step: one shot's gradient on a {prob['n']}x{prob['n']} crop-equivalent of the C2 model, nt={prob['nt']}, check_freq={prob['check_freq']}, 128 receivers"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot_t / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(c2_problem.__defaults__, args, note="reference arm runs a bounded sample of this workload on the host CPU"),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": int(os.environ["OMP_NUM_THREADS"]), "kind": "port", "sample": sample},
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
    lib = S._lib.load()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    prob = c2_problem(n=args.grid, nt=args.nt)
    if args.check_freq:
        prob["check_freq"] = args.check_freq
    n, nt, h = prob["n"], prob["nt"], prob["h"]
    T = np.float32
    bc = S.CPMLBoundaryConditionParameters(halo=prob["halo"], rcoef=T(1e-4), freeboundtop=True)
    params = S.InputParametersAcoustic(nt, T(prob["dt"]), (n, n), (T(h), T(h)), bc, dtype=np.dtype(T))
    runparams = S.RunParameters(parall="B200", device=local, erroronPPW=False, fast_f32=bool(args.fast_f32))
    gradparams = S.GradParameters(mute_radius_src=3, mute_radius_rec=0, compute_misfit=True, check_freq=prob["check_freq"])

    def pinned(a):
        t = torch.empty(a.shape[::-1], dtype=torch.float32, pin_memory=True)  # reversed shape: C-order tensor == F-order array
        v = t.numpy().T
        v[...] = a
        return v, t

    vp, _k1 = pinned(prob["vp"])
    rho, _k2 = pinned(prob["rho"])
    matprop = S.VpRhoAcousticVDMaterialProperties(vp, rho)
    tf = np.asfortranarray((1000.0 * S.gaussderivstf(prob["t"], 2.0 / prob["f0"], prob["f0"])).astype(T).reshape(nt, 1))
    rp = np.stack([prob["xr"], np.full_like(prob["xr"], 3 * h)], axis=1).astype(T)

    def make_shot(k):
        g = (rank * (args.steps + args.warmup) + k) % prob["nshots_total"]
        sp = np.array([[prob["xs"][g], 2 * h]], dtype=T)
        return S.ScalarShot(srcs=S.ScalarSources(sp, tf, T(prob["f0"])), recs=S.ScalarReceivers(rp, nt, dtype=np.dtype(T)))

    nrec = rp.shape[0]
    zeros_obs = np.zeros((nt, nrec), dtype=T, order="F")
    wavesim = S.build_wavesim(params, matprop, runparams=runparams, gradparams=gradparams, gradient=True)
    comm = None
    if world > 1:
        idbuf = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            raw = (S._lib.C.c_ubyte * 128)()
            S._lib.check(lib.swb_comm_unique_id(raw))
            idbuf = torch.tensor(list(raw), dtype=torch.uint8)
        idbuf = idbuf.cuda()
        dist.broadcast(idbuf, 0)
        raw = (S._lib.C.c_ubyte * 128)(*idbuf.cpu().tolist())
        comm = S._lib.C.c_void_p()
        S._lib.check(lib.swb_comm_create(raw, world, rank, local, S._lib.C.byref(comm)))

    stream_ptr = S._lib.C.c_void_p()
    S._lib.check(lib.swb_sim_stream(wavesim._h, S._lib.C.byref(stream_ptr)))
    ext = torch.cuda.ExternalStream(stream_ptr.value, device=torch.device("cuda", local))

    # ---- device-resident arm (`value`) -----------------------------------------------------------------
    wavesim.check_sim_consistency(matprop, [make_shot(0)])
    wavesim.set_wavesim_matprop(matprop)
    wavesim.zero_total_gradient()
    misfit_val = S._lib.C.c_double()

    def one_step(k):
        shot = make_shot(k)
        wavesim.init_shot(shot)  # C-PML profiles (host, O(halo)) + upload
        wavesim._bind(shot)  # nearest grid points, STF scaling (host, O(nt)) + upload of a few KB
        S._lib.check(lib.swb_sim_gradient_l2(wavesim._h, None, None, S._lib.C.byref(misfit_val)))
        sp, rpp = shot.srcs.positions, shot.recs.positions
        S._lib.check(lib.swb_sim_accumulate_gradient(wavesim._h, sp.shape[0], S.api._vp(np.asfortranarray(sp)), gradparams.mute_radius_src,
                                                     rpp.shape[0], S.api._vp(np.asfortranarray(rpp)), gradparams.mute_radius_rec))

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # started before the warm-up so that nvidia-smi's start-up cost is not inside the timed region
    for w in range(args.warmup):
        one_step(w)
    if comm is not None:  # warm-up of the collective too (NCCL sets up its channels on the first call)
        S._lib.check(lib.swb_sim_allreduce_total_gradient(wavesim._h, comm))
    wavesim.zero_total_gradient()
    wavesim.kernel_timing(1)
    barrier()
    t_from = time.perf_counter()
    cu0, l0 = wavesim.cell_updates(), lib.swb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for k in range(args.steps):
        one_step(args.warmup + k)
    if comm is not None:
        S._lib.check(lib.swb_sim_allreduce_total_gradient(wavesim._h, comm))
    e1.record(ext)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop(t_from, time.perf_counter()) if rank == 0 else None
    cu = wavesim.cell_updates() - cu0
    launches = lib.swb_launch_count() - l0
    wavesim.kernel_timing(0)
    (kt_ms, kt_n), (ka_ms, ka_n), (_, kr_n) = wavesim.kernel_timing_class(0), wavesim.kernel_timing_class(1), wavesim.kernel_timing_class(2)
    tms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    tcu = torch.tensor([float(cu)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tcu, op=dist.ReduceOp.SUM)
    ms_max, cu_all = float(tms.item()), float(tcu.item())
    value = cu_all / (ms_max * 1e-3) / 1e9

    # ---- end-to-end arm (`e2e`): the public API on host buffers -------------------------------------------
    shots = [make_shot(args.warmup + k) for k in range(args.steps)]
    misfits = [S.L2Misfit(observed=zeros_obs) for _ in shots]
    cu0 = wavesim.cell_updates()
    barrier()
    t0 = time.perf_counter()
    grad, mis = S.swgradient(wavesim, matprop, shots, misfits)
    if comm is not None:
        S._lib.check(lib.swb_sim_allreduce_total_gradient(wavesim._h, comm))
        grad = wavesim.get_total_gradient()
    barrier()
    t_e2e = time.perf_counter() - t0
    cu_e = wavesim.cell_updates() - cu0
    te = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
    tce = torch.tensor([float(cu_e)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.all_reduce(tce, op=dist.ReduceOp.SUM)
    e2e_val = float(tce.item()) / float(te.item()) / 1e9
    field_b = n * n * 4
    h2d = (2 * field_b) / args.steps + (nt * 4 + 2 * 8 + nrec * 2 * 8 + nt * nrec * 4)  # material (once per call) + per shot: STF, positions, observed
    d2h = nt * nrec * 4 + (2 * field_b * (2 if comm is not None else 1)) / args.steps  # seismograms per shot + gradient (vp, rho) per call
    assert np.isfinite(grad["vp"]).all(), "gradient is not finite"
    if args.grid == 4096 and args.nt == 1000:
        assert float(np.abs(grad["vp"]).max()) > 0, "gradient is empty"

    dominant_kernel, device_bytes = wavesim.dominant_kernel_name(), wavesim.device_bytes()
    wavesim.close()
    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md 6.65 TB/s)"
        # dominant kernel: the fused VD step (forward / re-forward sweeps); algorithmic bytes = 9 arrays x 4 B per cell
        bytes_per_cell = 9 * 4
        roof = None
        if kt_n > 0:
            dur = kt_ms / kt_n * 1e-3
            ach = bytes_per_cell * n * n / dur / 1e9
            roof = {"bound": "hbm", "kernel": dominant_kernel, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": 583.2e6 if (args.grid == 4096 and args.fast_f32) else None,
                    "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one forward launch, ncu --set full (profiles/r1_ncu_vd_fwd_v6.txt)",
                    "peak_source": peak_src, "frac_of_8TBs_nominal": ach / 8000.0, "avg_launch_us": dur * 1e6, "timed_launches": kt_n,
                    "algorithmic_bytes_per_launch": bytes_per_cell * n * n,
                    "sampling": "CUDA events on the engine's stream around each forward-sweep graph (nt step launches + the checkpoint "
                                "copies) inside the timed region; duration = elapsed / nt"}
            if ka_n > 0:  # the adjoint launch: adjoint step + injection + three correlations, 17 arrays x 4 B per cell
                # the adjoint graph also holds the re-forward launches (class 2 count); they are forward-step launches
                dur_a = (ka_ms - kr_n * (kt_ms / kt_n)) / ka_n * 1e-3
                ach_a = 17 * 4 * n * n / dur_a / 1e9
                roof["adjoint_kernel"] = {"achieved": ach_a, "frac": ach_a / peak, "avg_launch_us": dur_a * 1e6, "timed_launches": ka_n,
                                          "algorithmic_bytes_per_launch": 17 * 4 * n * n}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(None, args),
            "arith": "f32 storage, f32 arithmetic (SWB_FLAG_FAST_F32)" if args.fast_f32 else "f32 storage, f64 intermediates (the reference's promotion rule)",
            "clocks": clocks, "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "seconds": float(te.item())},
            "gpu_launches": int(launches), "roofline": roof, "misfit_last_shot": float(misfit_val.value),
            "useful_Gcell_per_s": 2.0 * n * n * nt * args.steps * world / (ms_max * 1e-3) / 1e9,
            "device_bytes": device_bytes,
        }
        if world == 1 and not args.no_extras and args.grid == 4096 and args.nt == 1000:
            line["other_configs"] = other_configs(peak)
        if not args.no_cpu:
            cores = os.cpu_count() or 1
            os.environ.setdefault("OMP_NUM_THREADS", str(cores))
            sp = cpu_sample_problem()
            cpu_gradient_once(sp, 0)  # warm-up (page faults, OpenMP pool)
            best = None
            for _ in range(2):
                dtc, u = cpu_gradient_once(sp, 1)
                r = u / dtc / 1e9
                best = r if best is None else max(best, r)
            line["cpu_baseline"] = {"value": best, "unit": UNIT, "cores": int(os.environ["OMP_NUM_THREADS"]), "kind": "port",
                                    "sample": f"one shot's gradient, {sp['n']}x{sp['n']} version of the C2 model, nt={sp['nt']}, check_freq={sp['check_freq']}, best of 2 "
                                              "(OpenMP build of the CPU restatement of the reference's Threads backend; Julia is not installed)"}
        print(json.dumps(line), flush=True)
    if comm is not None:
        lib.swb_comm_destroy(comm)
    if world > 1:
        dist.destroy_process_group()


def other_configs(peak):
    """Per-launch timings of the other fused engines on the BASELINE.json configurations that are not the headline (C3 elastic at
    full size, 2D / 3D constant-density at sizes that take seconds): CUDA events over whole replayed sweeps, a few dozen steps
    each (tools/bench_sim.py).  Informational: `value` above is C2 only."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_sim

    runs = [("C3 elastic P-SV 4096x2048 Float32 (f32 arithmetic)", "--kind ela --n 4096 2048 --nt 60 --check-freq 10 --dtype f32 --fast-f32 1 --nrec 10 --reps 2"),
            ("C3 elastic P-SV 4096x2048 Float64", "--kind ela --n 4096 2048 --nt 60 --check-freq 10 --dtype f64 --nrec 10 --reps 2"),
            ("2D acoustic CD 4096x4096 Float32 (C1 physics at roofline size)", "--kind cd --n 4096 4096 --nt 100 --check-freq 10 --reps 2"),
            ("C4-like 3D acoustic CD 512^3 Float32 (768^3 needs 155 GB of checkpoints: profiles/r1_cd_c4_fullsize.log)", "--kind cd --n 512 512 512 --nt 30 --check-freq 10 --reps 2")]
    out = []
    for name, argv in runs:
        try:
            r = bench_sim.measure(bench_sim.parse_args(argv.split()))
            rec = {"config": name, "device_GB": r["device_GB"]}
            for k in ("fwd", "adj"):
                if k in r:
                    rec[k] = {"us_per_step": r[k]["us"], "Gcell_per_s": r[k]["Gcell_s"], "algorithmic_GBps": r[k]["GBps"], "frac_of_peak": r[k]["GBps"] / peak}
            out.append(rec)
        except Exception as e:  # noqa: BLE001 -- the headline line must survive a failing extra
            out.append({"config": name, "error": str(e)[:200]})
    return out


def main():
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
