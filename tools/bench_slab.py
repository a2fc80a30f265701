#!/usr/bin/env python
"""tools/bench_slab.py -- forward-only throughput of the z-slab decomposition (BASELINE config 5 family), one rank per GPU:

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_slab.py --grid 2048 2048 1024 --nt 50

Each rank builds only its slab of the layered velocity model.  Rank 0 prints one JSON line: whole-job Gcell-updates/s
(global cells x nt / max-over-ranks device time of the timed forward run; tools/slab_check.py)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, nargs=3, default=[2048, 2048, 1024])
    ap.add_argument("--nt", type=int, default=50)
    ap.add_argument("--halo", type=int, default=20)
    ap.add_argument("--fast-f32", type=int, default=1)
    ap.add_argument("--check", action="store_true", help="also run the bitwise twin check first")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    import slab_check
    import swb200 as S

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if "RANK" in os.environ:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    out = {}
    if args.check and world > 1:
        ok, detail = slab_check.bitwise_twin(S, torch, dist, world, rank, local)
        out["bitwise_equal_to_single_gpu_on_twin"], out["twin"] = ok, detail
    out.update(slab_check.throughput(S, torch, dist, world, rank, local, grid=tuple(args.grid), nt=args.nt, halo=args.halo, fast_f32=bool(args.fast_f32)))
    if rank == 0:
        print(json.dumps(out), flush=True)
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
