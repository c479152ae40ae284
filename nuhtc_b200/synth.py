"""Seeded synthetic workloads for the BASELINE.json configs.  Host-side numpy/torch only;
used by tests/ and bench.py -- there is no network for slides or checkpoints."""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np
import torch

FPN_STRIDES = (4, 8, 16, 32)


def fpn_levels(B: int, C: int, frame: int = 512, seed: int = 0, device="cpu") -> List[torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(B, C, frame // s, frame // s, generator=g).to(device) for s in FPN_STRIDES]


def proposals(B: int, n_per: int, dist: str = "nuclei", frame: int = 512, seed: int = 0) -> torch.Tensor:
    """rois [B*n_per, 5] in the network frame.  'nuclei': side U(16,64) px (all route to level 0);
    'routed': side log-U(16,512) px (all four levels); 'clustered': like 'nuclei', but the proposals of a tile sit on ~150
    nucleus centres with a few px of jitter, the way RPN proposals that survived NMS(0.7) crowd around objects (analysis
    distribution for window sharing, see tools/window_sharing.py; not a bench default)."""
    g = torch.Generator().manual_seed(seed + 17)
    K = B * n_per
    ctr = torch.rand(K, 2, generator=g) * frame
    if dist == "nuclei":
        wh = 16 + torch.rand(K, 2, generator=g) * 48
    elif dist == "clustered":
        nn = 150
        cen = torch.rand(B, nn, 2, generator=g) * frame
        size = 16 + torch.rand(B, nn, 2, generator=g) * 48
        pick = torch.randint(0, nn, (B, n_per), generator=g)
        bi = torch.arange(B)[:, None].expand(B, n_per)
        ctr = (cen[bi, pick] + torch.randn(B, n_per, 2, generator=g) * 2.0).reshape(K, 2)
        wh = (size[bi, pick] * (0.9 + 0.2 * torch.rand(B, n_per, 2, generator=g))).reshape(K, 2)
    elif dist == "routed":
        side = torch.exp(torch.rand(K, 1, generator=g) * (math.log(512) - math.log(16)) + math.log(16))
        wh = side * (0.75 + 0.5 * torch.rand(K, 2, generator=g))
    else:
        raise ValueError(dist)
    x1y1 = (ctr - wh / 2).clamp(0, frame)
    x2y2 = (ctr + wh / 2).clamp(0, frame)
    bidx = torch.arange(B).repeat_interleave(n_per).to(torch.float32)[:, None]
    return torch.cat([bidx, x1y1, x2y2], dim=1)


def nms_boxes(N: int, num_classes: int = 5, seed: int = 0, density: float = 30.0) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Boxes w,h~U(8,32) with centres uniform over a square sized so that on average `density` same-class boxes
    cover a point (keep ratio ~30-60% at IoU 0.5); scores are a random permutation of a linspace (distinct)."""
    g = torch.Generator().manual_seed(seed + 31)
    side = math.sqrt(N / num_classes * 400.0 / density)
    ctr = torch.rand(N, 2, generator=g) * side
    wh = 8 + torch.rand(N, 2, generator=g) * 24
    boxes = torch.cat([ctr - wh / 2, ctr + wh / 2], dim=1)
    scores = torch.linspace(0.001, 0.999, N)[torch.randperm(N, generator=g)]
    labels = torch.randint(0, num_classes, (N,), generator=g)
    return boxes, scores, labels


def nuclei_masks(n: int, frame: int = 256, mask_size: int = 28, seed: int = 0, dup_frac: float = 0.3):
    """n detections in a frame: boxes [n,4] (frame px), 28x28 probability maps of an ellipse (already sigmoid),
    scores [n] distinct.  A fraction are jittered duplicates of earlier nuclei so mask NMS has work."""
    rng = np.random.default_rng(seed + 47)
    base = max(1, int(round(n * (1 - dup_frac))))
    cx = rng.uniform(8, frame - 8, base)
    cy = rng.uniform(8, frame - 8, base)
    a = rng.uniform(4, 12, base)
    b = rng.uniform(4, 12, base)
    pick = rng.integers(0, base, n - base)
    cx = np.concatenate([cx, cx[pick] + rng.uniform(-2, 2, n - base)])
    cy = np.concatenate([cy, cy[pick] + rng.uniform(-2, 2, n - base)])
    a = np.concatenate([a, a[pick] * rng.uniform(0.9, 1.1, n - base)])
    b = np.concatenate([b, b[pick] * rng.uniform(0.9, 1.1, n - base)])
    boxes = np.stack([cx - a - 1.5, cy - b - 1.5, cx + a + 1.5, cy + b + 1.5], 1).astype(np.float32)
    u = (np.arange(mask_size) + 0.5) / mask_size
    bw = boxes[:, 2] - boxes[:, 0]
    bh = boxes[:, 3] - boxes[:, 1]
    px = boxes[:, 0, None] + u[None] * bw[:, None]
    py = boxes[:, 1, None] + u[None] * bh[:, None]
    r2 = ((py[:, :, None] - cy[:, None, None]) / b[:, None, None]) ** 2 + ((px[:, None, :] - cx[:, None, None]) / a[:, None, None]) ** 2
    probs = 1.0 / (1.0 + np.exp(-6.0 * (1.0 - r2)))
    scores = rng.permutation(np.linspace(0.36, 0.999, n)).astype(np.float32)
    return torch.from_numpy(boxes), torch.from_numpy(probs.astype(np.float32))[:, None], torch.from_numpy(scores)


def ellipse_polygons(cx, cy, a, b, nverts: int = 32):
    """Integer-vertex rings of ellipses (contour-like polygons): returns (xy [sumV,2] float64, voff [N+1] int64).
    Consecutive duplicate vertices produced by rounding are removed."""
    N = len(cx)
    th = np.linspace(0, 2 * np.pi, nverts, endpoint=False)
    X = np.rint(cx[:, None] + a[:, None] * np.cos(th)[None]).astype(np.int64)
    Y = np.rint(cy[:, None] + b[:, None] * np.sin(th)[None]).astype(np.int64)
    keep = np.ones((N, nverts), dtype=bool)
    keep[:, 1:] = (X[:, 1:] != X[:, :-1]) | (Y[:, 1:] != Y[:, :-1])
    keep[:, 0] &= ~((X[:, 0] == X[:, -1]) & (Y[:, 0] == Y[:, -1]) & (keep.sum(1) > 1))
    counts = keep.sum(1)
    voff = np.zeros(N + 1, dtype=np.int64)
    voff[1:] = np.cumsum(counts)
    xy = np.stack([X[keep], Y[keep]], 1).astype(np.float64)
    return xy, voff


def slide_nuclei(tiles_x: int, tiles_y: int, per_tile: int = 23, tile: int = 256, stride: int = 192, seed: int = 0,
                 nverts: int = 32):
    """Synthetic slide for the merge (cfg 4/5): `per_tile` elliptical nuclei per tile in slide coordinates; a nucleus
    that falls in the overlap band of a neighbouring tile is also reported by that tile with a jittered contour and an
    independent score (the cross-tile duplicates nuclei_merge.py removes).
    Returns dict(xy, voff, score, tile_id) with distinct float64 scores."""
    rng = np.random.default_rng(seed + 101)
    T = tiles_x * tiles_y
    tx = np.repeat(np.arange(tiles_x)[None], tiles_y, 0).reshape(-1)
    ty = np.repeat(np.arange(tiles_y)[:, None], tiles_x, 1).reshape(-1)
    n0 = T * per_tile
    tid = np.repeat(np.arange(T), per_tile)
    # keep nuclei inside the tile minus a 12 px margin (infer_wsi.py drops detections that touch the tile margin)
    lx = rng.uniform(14, tile - 14, n0)
    ly = rng.uniform(14, tile - 14, n0)
    cx = tx[tid] * stride + lx
    cy = ty[tid] * stride + ly
    a = rng.uniform(4, 12, n0)
    b = rng.uniform(4, 12, n0)
    cxs, cys, as_, bs, tids = [cx], [cy], [a], [b], [tid]
    for dx, dy in ((1, 0), (-1, 0), (0, 1), (0, -1), (1, 1), (1, -1), (-1, 1), (-1, -1)):
        ntx, nty = tx[tid] + dx, ty[tid] + dy
        ok = (ntx >= 0) & (ntx < tiles_x) & (nty >= 0) & (nty < tiles_y)
        nlx, nly = cx - ntx * stride, cy - nty * stride
        ok &= (nlx >= 14) & (nlx <= tile - 14) & (nly >= 14) & (nly <= tile - 14)
        m = int(ok.sum())
        cxs.append(cx[ok] + rng.uniform(-2, 2, m))
        cys.append(cy[ok] + rng.uniform(-2, 2, m))
        as_.append(a[ok] * rng.uniform(0.92, 1.08, m))
        bs.append(b[ok] * rng.uniform(0.92, 1.08, m))
        tids.append((nty * tiles_x + ntx)[ok])
    cx, cy, a, b, tid = map(np.concatenate, (cxs, cys, as_, bs, tids))
    N = len(cx)
    perm = rng.permutation(N)
    cx, cy, a, b, tid = cx[perm], cy[perm], a[perm], b[perm], tid[perm]
    xy, voff = ellipse_polygons(cx, cy, a, b, nverts)
    score = rng.permutation(np.linspace(0.36, 0.999, N))
    return dict(xy=xy, voff=voff, score=score.astype(np.float64), tile_id=tid.astype(np.int64), tiles_x=tiles_x,
                tiles_y=tiles_y, stride=stride, tile=tile)


class SyntheticHeads:
    """Stand-ins for the bbox / mask heads (stock cuDNN modules, outside this package's target): their outputs are
    seeded tensors that do not depend on the pooled features, so the RoI-stage ops see realistic shapes and value
    ranges (class logits, small box deltas, blob-shaped mask logits) without any convolution in the timed region."""

    def __init__(self, K: int, num_classes: int = 5, num_stages: int = 3, pool: int = 509, mask_size: int = 28, seed: int = 0,
                 device="cpu"):
        g = torch.Generator().manual_seed(seed + 211)
        self.cls = [(torch.randn(K, num_classes + 1, generator=g) * 1.5).to(device) for _ in range(num_stages)]
        self.reg = [(torch.randn(K, 4, generator=g) * 0.5).to(device) for _ in range(num_stages)]
        u = (torch.arange(mask_size, dtype=torch.float32) + 0.5) / mask_size * 2 - 1
        a = 0.55 + 0.4 * torch.rand(pool, 1, 1, generator=g)
        b = 0.55 + 0.4 * torch.rand(pool, 1, 1, generator=g)
        r2 = (u[None, :, None] / b) ** 2 + (u[None, None, :] / a) ** 2
        self.pool = (6.0 * (1.0 - r2))[:, None].contiguous().to(device)   # logits; sigmoid > 0.5 inside the ellipse
        self.pool_n = pool

    def bbox_heads(self):
        return [(lambda feats, i=i: (self.cls[i], self.reg[i])) for i in range(len(self.cls))]

    def mask_head(self, feats, det_cand):
        return self.pool[det_cand % self.pool_n]

    def to(self, device):
        self.cls = [t.to(device) for t in self.cls]
        self.reg = [t.to(device) for t in self.reg]
        self.pool = self.pool.to(device)
        return self
