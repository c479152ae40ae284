"""Slide-level sharding: tiles are split over ranks by slide region (row stripes of the tile grid); the RoI stage,
paste and per-tile mask NMS need no communication; the cross-tile merge exchanges only the nuclei near stripe seams.

The reference runs tools/infer_wsi.py in one process per slide and merges afterwards with tools/nuclei_merge.py
(no communication anywhere); this module is the B200-side equivalent for N GPUs of one node.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from .nuclei_merge import merge_arrays

__all__ = ["stripe_rows", "shard_by_rows", "merge_sharded"]


def stripe_rows(tiles_y: int, rank: int, world: int):
    """Contiguous tile-row range [r0, r1) owned by `rank` (8 stripes of 26 rows for the 208-row cfg-4 slide)."""
    base, rem = divmod(tiles_y, world)
    r0 = rank * base + min(rank, rem)
    return r0, r0 + base + (1 if rank < rem else 0)


def shard_by_rows(slide: Dict, rank: int, world: int) -> Dict:
    """Nuclei whose TILE lies in this rank's stripe, with their global ids (index into the unsharded arrays)."""
    tiles_x, tiles_y = slide["tiles_x"], slide["tiles_y"]
    r0, r1 = stripe_rows(tiles_y, rank, world)
    trow = slide["tile_id"] // tiles_x
    sel = np.nonzero((trow >= r0) & (trow < r1))[0]
    cnt = np.diff(slide["voff"])[sel]
    voff = np.zeros(len(sel) + 1, dtype=np.int64)
    voff[1:] = np.cumsum(cnt)
    if len(sel):
        src = np.concatenate([np.arange(slide["voff"][i], slide["voff"][i + 1]) for i in sel]) if len(sel) < 4096 else \
            _ragged_gather(slide["voff"], sel, cnt)
        xy = slide["xy"][src]
    else:
        xy = np.zeros((0, 2), dtype=np.float64)
    return dict(xy=xy, voff=voff, score=slide["score"][sel], gid=sel.astype(np.int64), tile_id=slide["tile_id"][sel],
                rows=(r0, r1), tiles_x=tiles_x, tiles_y=tiles_y, stride=slide["stride"], tile=slide["tile"])


def _ragged_gather(voff: np.ndarray, sel: np.ndarray, cnt: np.ndarray) -> np.ndarray:
    starts = voff[sel]
    out_off = np.zeros(len(sel) + 1, dtype=np.int64)
    out_off[1:] = np.cumsum(cnt)
    idx = np.arange(out_off[-1], dtype=np.int64)
    seg = np.repeat(np.arange(len(sel)), cnt)
    return starts[seg] + (idx - out_off[seg])


def merge_sharded(xy: torch.Tensor, voff: torch.Tensor, score: torch.Tensor, shard: Dict, rank: int, world: int,
                  overlap_threshold: float = 0.05, merge_strategy: str = "probability") -> Optional[torch.Tensor]:
    """Merge the nuclei of a slide that is sharded over `world` ranks.  Returns this rank's kept LOCAL indices in
    score order.  world == 1 is the plain single-GPU merge."""
    if world == 1:
        return merge_arrays(xy, voff, score, overlap_threshold, merge_strategy)
    from .seam import merge_distributed
    return merge_distributed(xy, voff, score, shard, rank, world, overlap_threshold, merge_strategy)
