"""Worker of the z-slab decomposition test: launched with torch.distributed.run, one rank per GPU.

Every rank builds the same seeded global models, keeps its slab, runs the decomposed forward simulation (SlabForward3D, halo exchange
over NVLink) and rank 0 compares the gathered seismograms with the same shot computed on a single GPU by the ordinary engine
(tools/slab_check.py: the per-cell arithmetic is identical, so the traces must agree bit for bit)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    import torch
    import torch.distributed as dist

    import slab_check
    import swb200 as S

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    ok, detail = slab_check.bitwise_twin(S, torch, dist, world, rank, local, exchange="p2p")
    ok2, detail2 = slab_check.bitwise_twin(S, torch, dist, world, rank, local, exchange="nccl")
    ok, detail = ok and ok2, detail + detail2
    if rank == 0:
        for d in detail:
            print(f"slab test [{d['exchange']}] {d['dtype']} fast={d['fast_f32']} n={tuple(d['grid'])} world={world}: bitwise equal = {d['bitwise_equal']}, live traces = {d['live_traces']}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
