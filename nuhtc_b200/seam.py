"""Multi-GPU cross-tile merge: each rank owns the nuclei of its stripe of tile rows; only the nuclei that can
touch another stripe ("band" nuclei) travel, with one variable-length all-gather, and only their 1-byte states
travel during the resolve.

Why this is exact.  The greedy merge of tools/nuclei_merge.py:114-150 is the unique fixed point of
    kept(b)       <=> every overlapping (IoU > thr) nucleus that outranks b is suppressed
    suppressed(b) <=> some overlapping nucleus that outranks b is kept
over the global score order.  A nucleus lies inside its tile, so two nuclei of different ranks can only overlap if
each intersects the other stripe's extent -- i.e. both are band nuclei.  After the band exchange every rank
therefore holds ALL suppressors of its own nuclei (own nuclei + halo copies of foreign band nuclei) and can build
its part of the suppression graph locally (csrc/merge.cu, same kernels as the single-GPU merge).  The resolve rounds
update own nuclei only; between rounds the states of the band nuclei are all-gathered so halo copies follow their
owners.  The loop ends when no rank has an undecided own nucleus; the result equals the single-GPU merge bit for bit.
The final ``nuclei_id`` (rank of a kept nucleus among all kept, nuclei_merge.py:201) comes from one all-gather of
the kept (score, id) pairs.

``engine`` abstracts the two device calls so that the host protocol can be exercised on CPU with gloo (tests inject an
oracle-backed engine); the product engine is CUDA only.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import os
import time

import torch
import torch.distributed as dist

from . import _lib as L

__all__ = ["CudaMergeEngine", "merge_distributed", "stripe_extent"]


class CudaMergeEngine:
    """nuhtc_merge_graph / nuhtc_merge_rounds of libnuhtc_b200.so."""

    def graph(self, xy: torch.Tensor, voff: torch.Tensor, score: torch.Tensor, thr: float):
        L.require_cuda(xy, "xy")
        dev = xy.device
        N = score.numel()
        lib = L.lib()
        indeg = torch.zeros(max(N, 1), dtype=torch.int32, device=dev)
        in_off = torch.zeros(N + 1, dtype=torch.int32, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        cap = 8 * N + 1024
        import ctypes
        for _ in range(8):
            in_list = torch.empty(cap, dtype=torch.int32, device=dev)
            wsb = lib.nuhtc_merge_workspace_bytes(N, xy.shape[0], cap)
            ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
            npairs = ctypes.c_int64(0)
            with torch.cuda.device(dev):
                rc = lib.nuhtc_merge_graph(xy.data_ptr(), voff.data_ptr(), score.data_ptr(), N, xy.shape[0], float(thr), cap,
                                           indeg.data_ptr(), in_off.data_ptr(), in_list.data_ptr(), ctypes.byref(npairs),
                                           status.data_ptr(), ws.data_ptr(), wsb, L.stream_ptr(dev))
            if rc == -4:
                cap = int(npairs.value) + 1024
                continue
            L.check(rc, "merge_graph")
            L.count("merge")
            return indeg[:N], in_off, in_list
        raise L.NuhtcError("merge_graph: candidate-pair capacity could not be satisfied")

    def yextent(self, xy: torch.Tensor, voff: torch.Tensor):
        N = voff.numel() - 1
        ymin = torch.empty(N, dtype=torch.float64, device=xy.device)
        ymax = torch.empty(N, dtype=torch.float64, device=xy.device)
        with torch.cuda.device(xy.device):
            rc = L.lib().nuhtc_ring_yextent(xy.data_ptr(), voff.data_ptr(), N, ymin.data_ptr(), ymax.data_ptr(), L.stream_ptr(xy.device))
        L.check(rc, "ring_yextent")
        return ymin, ymax

    def rounds(self, in_off, indeg, in_list, frozen, state, nrounds: int, remaining: Optional[torch.Tensor] = None) -> torch.Tensor:
        dev = state.device
        if remaining is None:
            remaining = torch.zeros(1, dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            rc = L.lib().nuhtc_merge_rounds(in_off.data_ptr(), indeg.data_ptr(), in_list.data_ptr(), state.numel(),
                                            L.ptr(frozen), state.data_ptr(), remaining.data_ptr(), int(nrounds), L.stream_ptr(dev))
        L.check(rc, "merge_rounds")
        return remaining


def stripe_extent(shard_meta: Dict, rows: Tuple[int, int]) -> Optional[Tuple[float, float]]:
    """y-range covered by the tiles of a stripe of tile rows [r0, r1)."""
    r0, r1 = rows
    if r1 <= r0:
        return None
    return float(r0 * shard_meta["stride"]), float((r1 - 1) * shard_meta["stride"] + shard_meta["tile"])


def _all_gather_flat(out: torch.Tensor, inp: torch.Tensor, world: int, group) -> None:
    """out [world * n] <- every rank's inp [n] (one collective; the list form only where the backend lacks the flat one)."""
    if hasattr(dist, "all_gather_into_tensor") and inp.device.type == "cuda":
        dist.all_gather_into_tensor(out, inp, group=group)
    else:
        dist.all_gather(list(out.view(world, -1).unbind(0)), inp, group=group)


def _all_gather_ragged(t: torch.Tensor, counts: List[int], group) -> List[torch.Tensor]:
    """all-gather of per-rank tensors with different leading sizes: pad to the maximum, gather, trim."""
    mx = max(max(counts), 1)
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    outs = [torch.empty_like(pad) for _ in counts]
    dist.all_gather(outs, pad, group=group)
    return [o[:c] for o, c in zip(outs, counts)]


def _ring_area(xy: torch.Tensor, voff: torch.Tensor) -> torch.Tensor:
    """|shoelace| / 2 of every ring (closed or open), fp64."""
    cnt = voff[1:] - voff[:-1]
    n = xy.shape[0]
    idx = torch.arange(n, device=xy.device)
    seg = torch.repeat_interleave(torch.arange(cnt.numel(), device=xy.device), cnt)
    nxt = torch.where(idx + 1 == voff[1:][seg], voff[:-1][seg], idx + 1)
    cr = xy[:, 0] * xy[nxt, 1] - xy[nxt, 0] * xy[:, 1]
    return torch.segment_reduce(cr, "sum", lengths=cnt, unsafe=True).abs() * 0.5


def _best_per_group(group: torch.Tensor, area: torch.Tensor, score: torch.Tensor, gid: torch.Tensor):
    """For every distinct value of `group`: the candidate with the largest area, ties to the better global rank (score
    descending, then id) -- nuclei_merge.py:143-150 walks the hits in rank order and keeps the first maximum.  Returns
    (group values, candidate positions)."""
    p = torch.argsort(gid, stable=True)
    p = p[torch.argsort(-score[p], stable=True)]
    p = p[torch.argsort(-area[p], stable=True)]
    p = p[torch.argsort(group[p], stable=True)]
    g = group[p]
    first = torch.ones(g.numel(), dtype=torch.bool, device=g.device)
    first[1:] = g[1:] != g[:-1]
    return g[first], p[first]


def _area_picks(xy, voff, score, gid, a_score, a_gid, state, indeg, in_off, in_list, bidx, ncounts, rank, group):
    """merge_strategy='area' (nuclei_merge.py:143-150): every kept nucleus q is replaced by the largest nucleus among those it
    suppressed FIRST, i.e. among the suppressed c whose lowest-ranked kept suppressor is q.  Returns the flags of the own
    nuclei that are in the final set.

    Every suppressor of an own nucleus is in the local set, so owner(c) is a local computation.  The children of a kept q
    live on q's rank or are band nuclei of other ranks; one more exchange of the band records -- each carrying its owner's
    id, its own (area, score) and, for a kept nucleus, the best child found on its own rank -- lets every rank work out
    the same pick for every kept band nucleus."""
    dev = score.device
    N, M = score.numel(), a_score.numel()
    world = len(ncounts)
    area = _ring_area(xy, voff)
    o1 = torch.argsort(a_gid, stable=True)
    order = o1[torch.argsort(-a_score[o1], stable=True)]            # local set in global rank order
    lrank = torch.empty(M, dtype=torch.int64, device=dev)
    lrank[order] = torch.arange(M, device=dev)
    own_state = state[:N]
    # ---- owner(c) of every suppressed own nucleus: the best-ranked kept entry of its suppressor list
    deg = indeg[:N].to(torch.int64) * (own_state == 2)
    dst = torch.repeat_interleave(torch.arange(N, device=dev), deg)
    eoff = torch.cumsum(deg, 0) - deg
    src = in_list[in_off[:N].to(torch.int64)[dst] + (torch.arange(dst.numel(), device=dev) - eoff[dst])].to(torch.int64)
    live = state[src] == 1
    big = torch.iinfo(torch.int64).max
    owner_lrank = torch.full((N,), big, dtype=torch.int64, device=dev)
    owner_lrank.scatter_reduce_(0, dst[live], lrank[src[live]], "amin")
    child = (owner_lrank < big).nonzero().squeeze(1)
    owner = order[owner_lrank[child]]                               # local-set index of each child's owner
    owner_of = torch.full((N,), -1, dtype=torch.int64, device=dev)
    owner_of[child] = owner
    # ---- best own child of every local-set nucleus
    lb = torch.full((M,), -1, dtype=torch.int64, device=dev)
    if child.numel():
        q, pos = _best_per_group(owner, area[child], score[child], gid[child])
        lb[q] = child[pos]
    # ---- band records of every rank
    nb = bidx.numel()
    rec = torch.full((nb, 8), -1.0, dtype=torch.float64, device=dev)
    if nb:
        rec[:, 0] = gid[bidx].to(torch.float64)
        rec[:, 1] = score[bidx]
        rec[:, 2] = area[bidx]
        rec[:, 3] = own_state[bidx].to(torch.float64)
        ob = owner_of[bidx]
        rec[:, 4] = torch.where(ob >= 0, a_gid[ob.clamp(min=0)].to(torch.float64), rec[:, 4])
        lbb = lb[bidx]
        has = lbb >= 0
        lc = lbb.clamp(min=0)
        rec[:, 5] = torch.where(has, gid[lc].to(torch.float64), rec[:, 5])
        rec[:, 6] = torch.where(has, area[lc], rec[:, 6])
        rec[:, 7] = torch.where(has, score[lc], rec[:, 7])
    T = torch.cat(_all_gather_ragged(rec, ncounts, group)) if sum(ncounts) else rec
    t_rank = torch.repeat_interleave(torch.arange(world, device=dev), torch.as_tensor(ncounts, device=dev))
    t_gid = T[:, 0].to(torch.int64)
    # ---- candidates of the kept band nuclei: (a) best child on the owner's rank, (b) band children on other ranks
    a_sel = (T[:, 3] == 1) & (T[:, 5] >= 0)
    cand_q = [t_gid[a_sel]]
    cand = [(T[a_sel, 6], T[a_sel, 7], T[a_sel, 5].to(torch.int64))]
    b_sel = ((T[:, 3] == 2) & (T[:, 4] >= 0)).nonzero().squeeze(1)
    if b_sel.numel():
        sg, sp = torch.sort(t_gid)
        og = T[b_sel, 4].to(torch.int64)
        j = torch.searchsorted(sg, og).clamp(max=sg.numel() - 1)
        found = sg[j] == og                                          # the owner is a band nucleus ...
        foreign = found & (t_rank[sp[j]] != t_rank[b_sel])           # ... of another rank than the child's
        b_sel = b_sel[foreign]
        cand_q.append(og[foreign])
        cand.append((T[b_sel, 2], T[b_sel, 1], t_gid[b_sel]))
    cq = torch.cat(cand_q)
    c_area, c_score, c_gid = (torch.cat([c[i] for c in cand]) for i in range(3))
    flagged = torch.zeros(N, dtype=torch.bool, device=dev)
    band = torch.zeros(N, dtype=torch.bool, device=dev)
    band[bidx] = True
    kept = own_state == 1
    # kept nuclei away from the seam: children are all local
    inner = kept & ~band
    lbo = lb[:N]
    flagged |= inner & (lbo < 0)
    flagged[lbo[inner & (lbo >= 0)]] = True
    # kept band nuclei: the pick comes out of the exchanged records (the same on every rank); the picked nucleus' own rank flags it
    if cq.numel():
        qg, pos = _best_per_group(cq, c_area, c_score, c_gid)
        flagged |= torch.isin(gid, c_gid[pos])
        flagged |= kept & band & ~torch.isin(gid, qg)
    else:
        flagged |= kept & band
    return flagged


def merge_distributed(xy: torch.Tensor, voff: torch.Tensor, score: torch.Tensor, shard: Dict, rank: int, world: int,
                      overlap_threshold: float = 0.05, merge_strategy: str = "probability", engine=None, group=None,
                      return_ids: bool = False):
    """Returns this rank's kept LOCAL indices (score-descending); with ``return_ids`` also their global nuclei_id."""
    if merge_strategy not in ("probability", "area"):
        raise ValueError(f"Invalid merge strategy: {merge_strategy}. Use 'probability' or 'area'.")
    from .slide import stripe_rows
    engine = engine or CudaMergeEngine()
    dev = xy.device
    trace = os.environ.get("NUHTC_SEAM_TRACE") == "1" and rank == 0   # host wall-clock per phase (synchronises: debugging only)
    marks = []

    def mark(name):
        if trace:
            if dev.type == "cuda":
                torch.cuda.synchronize(dev)
            marks.append((name, time.perf_counter()))

    mark("start")
    N = score.numel()
    gid = torch.as_tensor(shard["gid"], dtype=torch.int64, device=dev)
    cnt = voff[1:] - voff[:-1]
    if hasattr(engine, "yextent"):
        ymin, ymax = engine.yextent(xy.contiguous(), voff)                      # one launch (torch.segment_reduce: 2 x 160 us at 100 k rings)
    else:
        ycol = xy[:, 1].contiguous()
        ymin = torch.segment_reduce(ycol, "min", lengths=cnt, unsafe=True)
        ymax = torch.segment_reduce(ycol, "max", lengths=cnt, unsafe=True)
    extents = [stripe_extent(shard, stripe_rows(shard["tiles_y"], q, world)) for q in range(world)]

    # ---- band = own nuclei that reach into another rank's stripe
    others = [e for q, e in enumerate(extents) if q != rank and e is not None]
    if others and N:
        E = torch.tensor(others, dtype=torch.float64, device=dev)                   # [world-1, 2]
        band = ((ymax[:, None] >= E[None, :, 0]) & (ymin[:, None] <= E[None, :, 1])).any(dim=1)
    else:
        band = torch.zeros(N, dtype=torch.bool, device=dev)
    bidx = band.nonzero().squeeze(1)
    nb = int(bidx.numel())
    bcnt = cnt[bidx]
    bvoff = torch.zeros(nb + 1, dtype=torch.int64, device=dev)
    bvoff[1:] = torch.cumsum(bcnt, 0)
    nv = int(bvoff[-1].item()) if nb else 0
    if nb:
        bseg = torch.repeat_interleave(torch.arange(nb, device=dev), bcnt, output_size=nv)
        src = voff[bidx][bseg] + (torch.arange(nv, device=dev) - bvoff[:-1][bseg])
        bxy = xy[src]
    else:
        bxy = xy[:0]
    mark("band")
    # ---- one exchange of the band nuclei: the counts, then ONE flat buffer per rank [5 doubles per nucleus | 2 per vertex]
    sizes = torch.tensor([nb, nv], dtype=torch.int64, device=dev)
    all_sizes_t = torch.empty(2 * world, dtype=torch.int64, device=dev)
    _all_gather_flat(all_sizes_t, sizes, world, group)
    all_sizes = all_sizes_t.view(world, 2).tolist()
    ncounts = [s[0] for s in all_sizes]
    vcounts = [s[1] for s in all_sizes]
    mxn, mxv = max(max(ncounts), 1), max(max(vcounts), 1)
    rec_len = 5 * mxn + 2 * mxv
    send = torch.zeros(rec_len, dtype=torch.float64, device=dev)
    if nb:
        sm = send[: 5 * nb].view(nb, 5)
        sm[:, 0] = score[bidx]
        sm[:, 1] = gid[bidx]
        sm[:, 2] = bcnt
        sm[:, 3] = ymin[bidx]
        sm[:, 4] = ymax[bidx]
        send[5 * mxn: 5 * mxn + 2 * nv].view(nv, 2).copy_(bxy)
    recv = torch.empty(world * rec_len, dtype=torch.float64, device=dev)
    _all_gather_flat(recv, send, world, group)
    recv = recv.view(world, rec_len)
    g_meta = [recv[q, : 5 * ncounts[q]].view(ncounts[q], 5) for q in range(world)]
    g_xy = [recv[q, 5 * mxn: 5 * mxn + 2 * vcounts[q]].view(vcounts[q], 2) for q in range(world)]

    mark("exchange")
    # ---- halo = foreign band nuclei that reach into MY stripe (one pass over the concatenated records of all ranks)
    me = extents[rank]
    f_meta = torch.cat(g_meta) if sum(ncounts) else torch.zeros((0, 5), dtype=torch.float64, device=dev)
    f_xy = torch.cat(g_xy) if sum(vcounts) else xy[:0]
    f_rank = torch.repeat_interleave(torch.arange(world, device=dev), torch.tensor(ncounts, device=dev), output_size=sum(ncounts))
    if me is not None and f_meta.shape[0]:
        take = (f_meta[:, 4] >= me[0]) & (f_meta[:, 3] <= me[1]) & (f_rank != rank)
        halo_flat = take.nonzero().squeeze(1)          # positions in the concatenated band list (rank-major, band order)
    else:
        halo_flat = torch.zeros(0, dtype=torch.int64, device=dev)
    H = int(halo_flat.numel())
    if H:
        c = f_meta[:, 2].to(torch.int64)
        off = torch.zeros(c.numel() + 1, dtype=torch.int64, device=dev)
        off[1:] = torch.cumsum(c, 0)
        tc = c[halo_flat]
        toff = torch.zeros(H + 1, dtype=torch.int64, device=dev)
        toff[1:] = torch.cumsum(tc, 0)
        nhv = int(toff[-1].item())
        tseg = torch.repeat_interleave(torch.arange(H, device=dev), tc, output_size=nhv)
        vsrc = off[:-1][halo_flat][tseg] + (torch.arange(tseg.numel(), device=dev) - toff[:-1][tseg])
        a_score = torch.cat([score, f_meta[halo_flat, 0]])
        a_gid = torch.cat([gid, f_meta[halo_flat, 1].to(torch.int64)])
        a_cnt = torch.cat([cnt, tc])
        a_xy = torch.cat([xy, f_xy[vsrc]])
    else:
        a_score, a_gid, a_cnt, a_xy = score, gid, cnt, xy
    M = N + H
    # The local set is own nuclei followed by the halo copies.  Scores are distinct by contract (ties are the one thing the
    # reference leaves open: pandas' quicksort), so no re-ordering by global id is needed: the graph kernels rank by score.
    p_voff = torch.zeros(M + 1, dtype=torch.int64, device=dev)
    p_voff[1:] = torch.cumsum(a_cnt, 0)
    p_xy = a_xy.contiguous()
    p_score = a_score.contiguous()

    mark("halo+permute")
    indeg, in_off, in_list = engine.graph(p_xy, p_voff, p_score, overlap_threshold)
    mark("graph")
    frozen = torch.zeros(M, dtype=torch.uint8, device=dev)
    frozen[N:] = 1                                   # halo copies: read, never written (their owners decide)
    state = (indeg[:M] == 0).to(torch.uint8)
    state[N:] = 0
    band_pos = bidx

    # ---- resolve: sweep own nuclei, exchange band states, until nobody has an undecided own nucleus.  ONE collective per
    # iteration: every rank appends its count of undecided own nuclei (8 bytes) to its band states, so the termination test
    # rides on the state exchange instead of a separate all-reduce (on 8 GPUs every extra collective is another point where
    # the slowest host holds everybody up).
    # Every rank sends one fixed-size message: [undecided own nuclei (int64) | band states | padding]; the halo copies
    # read their owners' states straight out of the gathered buffer through an index computed once.
    mx = 8 * ((8 + max(max(ncounts), 1) + 7) // 8)
    msg = torch.zeros(mx, dtype=torch.uint8, device=dev)
    msg_remaining = msg[:8].view(torch.int64)              # the rounds kernel writes its counter straight into the message
    msg_band = msg[8:8 + nb]
    gathered = torch.empty(world * mx, dtype=torch.uint8, device=dev)
    g_remaining = gathered.view(world, mx)[:, :8]
    state_halo = state[N:]
    if H:
        starts = torch.zeros(world + 1, dtype=torch.int64, device=dev)
        starts[1:] = torch.cumsum(torch.as_tensor(ncounts, device=dev), 0)
        hr = f_rank[halo_flat]
        halo_src = hr * mx + 8 + (halo_flat - starts[hr])      # position of each halo nucleus' state in `gathered`
    blind = 3          # iterations between two looks at the termination counter: the host only waits once per `blind` exchanges
    done = False
    iters = 0
    for _ in range(1 << 18):
        for _k in range(blind):
            # the first call settles the interior (dependency chains inside a crowded tile overlap run a few dozen deep);
            # afterwards only what hangs on halo states is left
            engine.rounds(in_off, indeg, in_list, frozen, state, 16 if iters == 0 else 4, msg_remaining)
            if nb:
                torch.index_select(state, 0, band_pos, out=msg_band)
            _all_gather_flat(gathered, msg, world, group)
            if H:
                torch.index_select(gathered, 0, halo_src, out=state_halo)
            iters += 1
        # an iteration that found nothing undecided anywhere leaves every later one a no-op, so looking at the last one is enough
        if int(g_remaining.contiguous().view(torch.int64).sum().item()) == 0:
            done = True
            break
    assert done
    mark("resolve")
    if trace:
        print("seam trace (ms):", ", ".join(f"{b[0]} {1e3 * (b[1] - a[1]):.2f}" for a, b in zip(marks, marks[1:])),
              f"| own {N} band {nb} halo {H} resolve iterations {iters}", flush=True)
    if merge_strategy == "area":
        flagged = _area_picks(xy, voff, score, gid, a_score, a_gid, state, indeg, in_off, in_list, bidx, ncounts, rank, group)
        mark("area picks")
    else:
        flagged = state[:N] == 1
    kept_local = flagged.nonzero().squeeze(1)
    order = torch.argsort(-score[kept_local], stable=True)
    kept_local = kept_local[order]
    if not return_ids:
        return kept_local
    # ---- global nuclei_id = rank among all kept nuclei (score desc, then id)
    kc = torch.tensor([int(kept_local.numel())], dtype=torch.int64, device=dev)
    all_kc = [torch.empty_like(kc) for _ in range(world)]
    dist.all_gather(all_kc, kc, group=group)
    all_kc = [int(k.item()) for k in all_kc]
    rec = torch.stack([score[kept_local], gid[kept_local].to(torch.float64)], dim=1)
    g_rec = torch.cat(_all_gather_ragged(rec, all_kc, group))
    o1 = torch.argsort(g_rec[:, 1], stable=True)
    o2 = torch.argsort(-g_rec[o1, 0], stable=True)
    glob_order = o1[o2]                                  # positions sorted by (score desc, id asc)
    pos = torch.empty_like(glob_order)
    pos[glob_order] = torch.arange(glob_order.numel(), device=dev)
    start = sum(all_kc[:rank])
    return kept_local, pos[start: start + all_kc[rank]]
