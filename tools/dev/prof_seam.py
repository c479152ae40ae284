"""torchrun --nproc-per-node 2 tools/dev/prof_seam.py: torch.profiler table of one distributed merge at the bench's size."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
from nuhtc_b200 import synth
from nuhtc_b200.seam import merge_distributed
from nuhtc_b200.slide import shard_by_rows
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
slide = synth.slide_nuclei(16, 20 * world, per_tile=190, seed=7)     # ~61k nuclei per rank like the 20-step bench run
sh = shard_by_rows(slide, rank, world)
xy, voff, score = (torch.from_numpy(sh[k]).to(dev) for k in ("xy", "voff", "score"))
sh["gid"] = torch.from_numpy(sh["gid"]).to(dev)
for _ in range(3):
    merge_distributed(xy, voff, score, sh, rank, world, 0.05)
dist.barrier(); torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(5):
    merge_distributed(xy, voff, score, sh, rank, world, 0.05)
torch.cuda.synchronize()
if rank == 0: print("own", score.numel(), "ms per merge", (time.perf_counter() - t0) / 5 * 1e3)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    merge_distributed(xy, voff, score, sh, rank, world, 0.05)
    torch.cuda.synchronize()
if rank == 0:
    print(prof.key_averages().table(sort_by="cpu_time_total", row_limit=28, max_name_column_width=44))
dist.barrier(); dist.destroy_process_group()
