"""world_size-2 CPU test (gloo) of the shot-parallel host logic: the shot split is the reference's distribsrcs, and the
all-reduced per-rank gradient sums equal the sequential shot loop.  The per-shot compute here is the CPU oracle (tests may
use it); on the GPU box the same ShotParallel object reduces the engine's device-resident totals over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nshots, outdir):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import swb200 as S
    from swb200.multigpu import ShotParallel, shot_group
    from cases import acoustic_case, case_params_oracle, make_observed, matprop_list, oracle_forward, oracle_shots
    from oracle import oracle as O

    case = acoustic_case(kind="acoustic_cd", n=(40, 36), nt=60, halo=5, nshots=nshots, seed=3)
    syn, _ = oracle_forward(case)
    obs = make_observed(case, syn)
    sp = ShotParallel()
    group = shot_group(nshots, world, rank)
    shots = oracle_shots(case)
    sim = O.build_wavesim("acoustic_cd", case_params_oracle(case), gradient=True, check_freq=1)
    local = {"vp": np.zeros(case["n"], order="F")}
    local_mis = 0.0
    if len(group) > 0:
        res = O.swgradient(sim, matprop_list(case), [shots[s] for s in group], [O.L2Misfit(observed=obs[s]) for s in group], compute_misfit=True)
        local, local_mis = res
    tot, mis = sp.allreduce_host(local, float(local_mis))
    np.save(os.path.join(outdir, f"grad_{rank}.npy"), tot["vp"])
    np.save(os.path.join(outdir, f"mis_{rank}.npy"), np.array([mis, group.start if len(group) else -1, len(group)]))
    dist.destroy_process_group()


@pytest.mark.parametrize("nshots", [3, 1])
def test_shot_sharding_allreduce_equals_sequential_loop(tmp_path, nshots):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, nshots, str(tmp_path)), nprocs=world, join=True)
    from cases import acoustic_case, case_params_oracle, make_observed, matprop_list, oracle_forward, oracle_shots
    from oracle import oracle as O

    case = acoustic_case(kind="acoustic_cd", n=(40, 36), nt=60, halo=5, nshots=nshots, seed=3)
    syn, _ = oracle_forward(case)
    obs = make_observed(case, syn)
    sim = O.build_wavesim("acoustic_cd", case_params_oracle(case), gradient=True, check_freq=1)
    ref, refmis = O.swgradient(sim, matprop_list(case), oracle_shots(case), [O.L2Misfit(observed=o) for o in obs], compute_misfit=True)
    g0, g1 = np.load(tmp_path / "grad_0.npy"), np.load(tmp_path / "grad_1.npy")
    m0, m1 = np.load(tmp_path / "mis_0.npy"), np.load(tmp_path / "mis_1.npy")
    assert np.array_equal(g0, g1)  # every rank holds the same reduced gradient
    assert np.allclose(g0, ref["vp"], rtol=1e-13, atol=0)
    assert abs(m0[0] - refmis) <= 1e-12 * abs(refmis) and m0[0] == m1[0]
    # contiguous groups, first ranks get the extra shots (distribsrcs)
    counts = [int(m0[2]), int(m1[2])]
    assert sum(counts) == nshots and counts[0] >= counts[1]


def test_distribsrcs_matches_reference_rule():
    from swb200 import distribsrcs

    assert [list(g) for g in distribsrcs(7, 3)] == [[0, 1, 2], [3, 4], [5, 6]]
    assert [list(g) for g in distribsrcs(64, 8)] == [list(range(8 * k, 8 * k + 8)) for k in range(8)]
    assert [list(g) for g in distribsrcs(2, 4)] == [[0], [1]]


def _slab_worker(rank, world, port, outdir):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import swb200 as S
    from swb200.multigpu import ShotParallel

    sp = ShotParallel()
    h = S._lib.swb_slab_handle()
    h.nz, h.plane_elems, h.device, h.pid = 100 + rank, 4096, rank, os.getpid()
    for k in range(3):
        h.raw[k] = 1000 * (rank + 1) + k
    for k in range(192):
        h.ipc[k] = (rank * 7 + k) % 256
    got = sp.allgather_bytes(bytes(h))
    hs = [S._lib.swb_slab_handle.from_buffer_copy(b) for b in got]
    np.save(os.path.join(outdir, f"slab_{rank}.npy"), np.array([[x.nz, x.device, x.raw[2], x.ipc[191], x.pid] for x in hs], dtype=np.int64))
    dist.destroy_process_group()


def test_slab_handles_travel_in_rank_order(tmp_path):
    """the host side of the peer-memory slab exchange: every rank receives every rank's 240-byte handle, in rank order"""
    world = 2
    mp.spawn(_slab_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    a, b = np.load(tmp_path / "slab_0.npy"), np.load(tmp_path / "slab_1.npy")
    assert np.array_equal(a, b) and a.shape == (2, 5)
    assert list(a[:, 0]) == [100, 101] and list(a[:, 1]) == [0, 1] and list(a[:, 2]) == [1002, 2002]
    assert list(a[:, 3]) == [191 % 256, (7 + 191) % 256] and a[0, 4] != a[1, 4]


def test_slab_partition_and_point_routing():
    from swb200.multigpu import slab_local_planes, slab_range, slab_route_points

    for nz, world in ((1024, 8), (91, 3), (75, 4), (61, 2)):
        own = [slab_range(nz, world, r) for r in range(world)]
        assert own[0].start == 0 and own[-1].stop == nz and all(own[r].stop == own[r + 1].start for r in range(world - 1))
        assert max(len(o) for o in own) - min(len(o) for o in own) <= 1
        for r in range(world):
            loc = slab_local_planes(nz, world, r)
            assert loc.start == own[r].start - (r > 0) and loc.stop == own[r].stop + (r < world - 1)
        idx = np.stack([np.zeros(nz, dtype=np.int64), np.zeros(nz, dtype=np.int64), np.arange(nz)], axis=1)
        routed = [slab_route_points(idx, nz, world, r) for r in range(world)]
        assert sorted(np.concatenate(routed).tolist()) == list(range(nz))  # every plane's points have exactly one owner
