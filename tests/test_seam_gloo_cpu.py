"""CPU: the multi-rank merge protocol (band exchange, halo graph, state exchange, global nuclei_id) with
world_size 2 and 3 over gloo.  The two device calls are replaced by an oracle-backed engine (tests may use the
oracle; the product engine is CUDA only), so what is exercised here is the host-side logic of nuhtc_b200/seam.py.
The result must equal the single-process merge bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleEngine:
    """graph / rounds with the CPU oracle's polygon IoU (small inputs only)."""

    def graph(self, xy, voff, score, thr):
        from oracle import cpu as O
        xy, voff, score = xy.numpy(), voff.numpy(), score.numpy()
        N = len(score)
        rings = [xy[voff[i]:voff[i + 1]] for i in range(N)]
        env = np.array([[r[:, 0].min(), r[:, 1].min(), r[:, 0].max(), r[:, 1].max()] for r in rings]) if N else np.zeros((0, 4))
        order = sorted(range(N), key=lambda i: (-score[i], i))
        rank = np.empty(N, dtype=np.int64)
        rank[order] = np.arange(N)
        ins = [[] for _ in range(N)]
        for a in range(N):
            for b in range(N):
                if a == b or rank[a] > rank[b]:
                    continue
                if env[a, 0] > env[b, 2] or env[a, 2] < env[b, 0] or env[a, 1] > env[b, 3] or env[a, 3] < env[b, 1]:
                    continue
                if O.poly_iou(rings[a], rings[b]) > thr:
                    ins[b].append(a)
        indeg = torch.tensor([len(x) for x in ins], dtype=torch.int32)
        in_off = torch.zeros(N + 1, dtype=torch.int32)
        in_off[1:] = torch.cumsum(indeg, 0)
        in_list = torch.tensor([a for x in ins for a in x] + [0], dtype=torch.int32)
        return indeg, in_off, in_list

    def rounds(self, in_off, indeg, in_list, frozen, state, nrounds, remaining=None):
        rem = 0
        for _ in range(nrounds):
            rem = 0
            snap = state.clone()
            for i in range(state.numel()):
                if snap[i] != 0 or (frozen is not None and frozen[i]):
                    continue
                sup = in_list[in_off[i]: in_off[i] + indeg[i]].long()
                st = snap[sup]
                if (st == 1).any():
                    state[i] = 2
                elif (st == 2).all():
                    state[i] = 1
                else:
                    rem += 1
        if remaining is not None:
            remaining.fill_(rem)
            return remaining
        return torch.tensor([rem], dtype=torch.int64)


def _worker(rank, world, port, tiles, strategy, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from nuhtc_b200 import synth
        from nuhtc_b200.seam import merge_distributed
        from nuhtc_b200.slide import shard_by_rows
        slide = synth.slide_nuclei(tiles[0], tiles[1], per_tile=6, seed=4)
        sh = shard_by_rows(slide, rank, world)
        kept, ids = merge_distributed(torch.from_numpy(sh["xy"]), torch.from_numpy(sh["voff"]), torch.from_numpy(sh["score"]), sh,
                                      rank, world, 0.05, merge_strategy=strategy, engine=OracleEngine(), return_ids=True)
        ret[rank] = (sh["gid"][kept.numpy()].tolist(), ids.tolist())
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("strategy", ["probability", "area"])
@pytest.mark.parametrize("world,tiles", [(2, (3, 4)), (3, (2, 5)), (4, (2, 3))])   # the last one leaves rank 3 without tiles
def test_distributed_merge_equals_single_process(oracle, world, tiles, strategy):
    from nuhtc_b200 import synth
    slide = synth.slide_nuclei(tiles[0], tiles[1], per_tile=6, seed=4)
    ref = oracle.merge_overlap_arrays(slide["xy"], slide["voff"], slide["score"], 0.05, strategy)
    if strategy == "area":   # the two strategies keep different nuclei on this slide, some of them picked across the seam
        assert ref.tolist() != oracle.merge_overlap_arrays(slide["xy"], slide["voff"], slide["score"], 0.05).tolist()
    assert len(ref) < len(slide["score"])  # there are cross-tile duplicates, some of them across the stripe seam
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), tiles, strategy, ret), nprocs=world, join=True)
    got = {}
    for r in range(world):
        gids, ids = ret[r]
        for g, i in zip(gids, ids):
            got[i] = g
    assert sorted(got) == list(range(len(ref)))                     # nuclei_id 0..k-1, each assigned once
    assert [got[i] for i in range(len(ref))] == ref.tolist()        # same nuclei, same ids as the one-process merge


def test_stripes_cover_the_slide():
    from nuhtc_b200.slide import shard_by_rows, stripe_rows
    from nuhtc_b200 import synth
    assert [stripe_rows(208, r, 8) for r in range(8)] == [(26 * r, 26 * r + 26) for r in range(8)]
    assert [stripe_rows(5, r, 3) for r in range(3)] == [(0, 2), (2, 4), (4, 5)]
    slide = synth.slide_nuclei(4, 7, per_tile=5, seed=1)
    parts = [shard_by_rows(slide, r, 3) for r in range(3)]
    allg = np.sort(np.concatenate([p["gid"] for p in parts]))
    assert (allg == np.arange(len(slide["score"]))).all()
    for p in parts:
        for j, g in enumerate(p["gid"][:50]):
            a = slide["xy"][slide["voff"][g]:slide["voff"][g + 1]]
            assert (p["xy"][p["voff"][j]:p["voff"][j + 1]] == a).all()
