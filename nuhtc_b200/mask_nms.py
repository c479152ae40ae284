"""Per-tile mask NMS: drop-in for ``mask_nms`` of /root/reference/tools/infer_wsi.py:60-84.

The reference RLE-encodes every mask with pycocotools, computes the n x n IoU matrix on the host and
runs a Python double loop.  Here the masks become bit rows on the GPU, the IoUs are popcounts and the
greedy scan is the shared suppression kernel (csrc/mask_nms.cu); results (kept indices and their
order) are identical.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch

from . import _lib as L

__all__ = ["mask_nms", "mask_nms_device", "pack_masks", "rle_encode"]


def pack_masks(masks: torch.Tensor):
    """uint8/bool [n,h,w] CUDA -> (bits int64 [n,h,ceil(w/64)], area int32 [n], bbox int32 [n,4])."""
    L.require_cuda(masks, "masks")
    if masks.dtype == torch.bool:
        masks = masks.view(torch.uint8)
    assert masks.dtype == torch.uint8 and masks.dim() == 3
    masks = masks.contiguous()
    n, h, w = masks.shape
    dev = masks.device
    bits = torch.empty((n, h, (w + 63) // 64), dtype=torch.int64, device=dev)
    area = torch.empty(n, dtype=torch.int32, device=dev)
    bbox = torch.empty((n, 4), dtype=torch.int32, device=dev)
    if n:
        with torch.cuda.device(dev):
            rc = L.lib().nuhtc_pack_masks(masks.data_ptr(), n, h, w, bits.data_ptr(), area.data_ptr(), bbox.data_ptr(),
                                          L.stream_ptr(dev))
        L.check(rc, "pack_masks")
        L.count("pack")
    return bits, area, bbox


def mask_nms_device(bits: torch.Tensor, area: torch.Tensor, bbox: torch.Tensor, scores: torch.Tensor, width: int, thr: float,
                    tile: Optional[torch.Tensor] = None, num_tiles: int = 1, max_tile_size: Optional[int] = None):
    """Batched over tiles, no host sync.  Returns (keep [n] int32, tile_start [T] int32, tile_count [T] int32,
    status [1] int32): tile t's kept original indices in score order are keep[tile_start[t] : +tile_count[t]]."""
    L.require_cuda(bits, "bits")
    n, h, wpm = bits.shape
    dev = bits.device
    assert wpm == (width + 63) // 64
    scores = scores.to(torch.float32).contiguous()
    if tile is not None:
        tile = tile.to(torch.int32).contiguous()
    mts = n if max_tile_size is None else int(max_tile_size)
    mts = max(1, min(mts, max(n, 1)))
    keep = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    tstart = torch.empty(num_tiles, dtype=torch.int32, device=dev)
    tcount = torch.empty(num_tiles, dtype=torch.int32, device=dev)
    status = torch.empty(1, dtype=torch.int32, device=dev)
    lib = L.lib()
    wsb = lib.nuhtc_mask_nms_workspace_bytes(n, num_tiles, mts)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.nuhtc_mask_nms(bits.data_ptr(), area.data_ptr(), bbox.data_ptr(), scores.data_ptr(), L.ptr(tile), n, num_tiles,
                                mts, h, int(width), float(thr), keep.data_ptr(), tstart.data_ptr(), tcount.data_ptr(),
                                status.data_ptr(), ws.data_ptr(), wsb, L.stream_ptr(dev))
    L.check(rc, "mask_nms")
    L.count("mask_nms")
    return keep, tstart, tcount, status


def _runs(mask: np.ndarray) -> np.ndarray:
    """Column-major run lengths of one [h,w] mask, starting with a run of zeros (possibly empty): pycocotools rleEncode."""
    flat = np.asarray(mask, dtype=bool).T.reshape(-1)
    change = np.flatnonzero(flat[1:] != flat[:-1]) + 1
    edges = np.concatenate([[0], change, [flat.size]])
    runs = np.diff(edges)
    if flat.size and flat[0]:
        runs = np.concatenate([[0], runs])
    return runs.astype(np.int64)


def rle_counts_to_string(counts) -> bytes:
    """pycocotools' compressed `counts` (common/maskApi.c rleToString): every count from the fourth on is stored as its
    difference to the count two places back, each value in little-endian groups of 5 bits + a continuation bit, as ASCII
    characters 48..111."""
    out = bytearray()
    cnts = [int(c) for c in counts]
    for i, x in enumerate(cnts):
        if i > 2:
            x -= cnts[i - 2]
        more = True
        while more:
            c = x & 0x1F
            x >>= 5                                  # arithmetic shift: negative differences end on x == -1
            more = (x != -1) if (c & 0x10) else (x != 0)
            if more:
                c |= 0x20
            out.append(c + 48)
    return bytes(out)


def rle_string_to_counts(s: bytes) -> List[int]:
    """Inverse of rle_counts_to_string (maskApi.c rleFrString)."""
    cnts: List[int] = []
    p = 0
    while p < len(s):
        x, k, more = 0, 0, True
        while more:
            c = s[p] - 48
            x |= (c & 0x1F) << (5 * k)
            more = bool(c & 0x20)
            p += 1
            k += 1
            if not more and (c & 0x10):
                x |= -1 << (5 * k)
        if len(cnts) > 2:
            x += cnts[-2]
        cnts.append(x)
    return cnts


def rle_encode(mask: np.ndarray, compressed: bool = True) -> dict:
    """`maskUtils.encode(np.asfortranarray(mask))` for one [h,w] mask on the host: {'size': [h, w], 'counts': bytes}
    (compressed=False: the run lengths as a list, pycocotools' "uncompressed RLE")."""
    h, w = mask.shape
    runs = _runs(mask)
    return {"size": [h, w], "counts": rle_counts_to_string(runs) if compressed else runs.tolist()}


def rle_decode(rle: dict) -> np.ndarray:
    """`maskUtils.decode` of one RLE dict -> [h,w] uint8."""
    h, w = rle["size"]
    counts = rle["counts"]
    if isinstance(counts, (bytes, str)):
        counts = rle_string_to_counts(counts.encode() if isinstance(counts, str) else counts)
    flat = np.zeros(h * w, dtype=np.uint8)
    pos, v = 0, 0
    for c in counts:
        if v:
            flat[pos:pos + c] = 1
        pos += c
        v ^= 1
    return flat.reshape(w, h).T.copy()


def mask_nms(masks, pred_scores, thr: float = 0.9, min_area=None) -> Tuple[List[dict], np.ndarray]:
    """Same call as the reference: ``masks`` is a list/array of [h,w] uint8 masks (host or CUDA),
    ``pred_scores`` their scores.  Returns (kept masks as pycocotools RLE dicts with compressed `counts`, kept indices in
    score order)."""
    if isinstance(masks, torch.Tensor):
        m = masks
    else:
        m = torch.from_numpy(np.ascontiguousarray(np.asarray(masks)))
    if not m.is_cuda:
        m = m.cuda(non_blocking=True)
    s = pred_scores if isinstance(pred_scores, torch.Tensor) else torch.from_numpy(np.asarray(pred_scores, dtype=np.float32))
    s = s.to(m.device, torch.float32)
    n = m.shape[0]
    if n == 0:
        return [], np.zeros(0, dtype=np.int64)
    bits, area, bbox = pack_masks(m)
    keep, _, tcount, _ = mask_nms_device(bits, area, bbox, s, m.shape[2], thr)
    k = int(tcount.item())
    idx = keep[:k].cpu().numpy().astype(np.int64)
    host = m.cpu().numpy() if isinstance(masks, torch.Tensor) else np.asarray(masks)
    return [rle_encode(host[i]) for i in idx], idx
