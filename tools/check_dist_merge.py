#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/check_dist_merge.py [tiles_x tiles_y [probability|area]]: the N-GPU merge (NCCL seam exchange) must
reproduce the single-GPU merge bit for bit (kept nuclei and their nuclei_id)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

import nuhtc_b200 as nb
from nuhtc_b200 import synth
from nuhtc_b200.seam import merge_distributed
from nuhtc_b200.slide import shard_by_rows


def main():
    tx, ty = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (32, 32)
    strategy = sys.argv[3] if len(sys.argv) > 3 else "probability"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    slide = synth.slide_nuclei(tx, ty, per_tile=23, seed=7)
    sh = shard_by_rows(slide, rank, world)
    xy, voff, score = (torch.from_numpy(sh[k]).to(dev) for k in ("xy", "voff", "score"))
    kept, ids = merge_distributed(xy, voff, score, sh, rank, world, 0.05, strategy, return_ids=True)   # warm-up (NCCL channels, allocator)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    kept, ids = merge_distributed(xy, voff, score, sh, rank, world, 0.05, strategy, return_ids=True)
    e1.record()
    torch.cuda.synchronize()
    tms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    gids = torch.from_numpy(sh["gid"]).to(dev)[kept]
    n = torch.tensor([kept.numel()], device=dev)
    ns = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(ns, n)
    mx = int(max(x.item() for x in ns))
    pad = torch.full((mx, 2), -1, dtype=torch.int64, device=dev)
    pad[: kept.numel(), 0] = gids
    pad[: kept.numel(), 1] = ids
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    ok = True
    if rank == 0:
        allp = torch.cat([o[: int(c.item())] for o, c in zip(outs, ns)]).cpu().numpy()
        by_id = allp[np.argsort(allp[:, 1])]
        full = [torch.from_numpy(slide[k]).to(dev) for k in ("xy", "voff", "score")]
        ref = nb.merge_arrays(*full, 0.05, strategy).cpu().numpy()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        nb.merge_arrays(*full, 0.05, strategy)
        s1.record()
        torch.cuda.synchronize()
        ok = len(by_id) == len(ref) and (by_id[:, 1] == np.arange(len(ref))).all() and (by_id[:, 0] == ref).all()
        n_all = len(slide["score"])
        print(f"dist-merge[{strategy}] world={world} nuclei={n_all} kept={len(ref)} match={ok} | {world}-GPU merge {float(tms.item()):.2f} ms "
              f"({n_all / float(tms.item()) / 1e3:.1f} M nuclei/s), single-GPU merge {s0.elapsed_time(s1):.2f} ms", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
