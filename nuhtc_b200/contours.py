"""Mask -> contour -> GeoJSON, the tile post-processing between the mask NMS and the cross-tile merge.

Reference: tools/infer_wsi.py
  :51-54   mask2inst: cv2.findContours(mask, RETR_TREE, CHAIN_APPROX_SIMPLE)[0][0] + its first point again
  :528-533 per kept nucleus: contour, drop len < 3, add the tile coordinate
  :536     boxes shifted by the tile coordinate
  :541-585 QuPath Feature dicts (Polygon + centre Point), dumped as a flat list (:661-664)
The contours are traced on the GPU from the bit-row masks (`nuhtc_mask_contours`), so the dense masks never leave the device;
`rings_for_merge` lays the kept contours out as the fp64 ring arrays `nuclei_merge.merge_arrays` consumes.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import NuhtcError, count, lib

__all__ = ["mask_contours", "mask2inst", "rings_for_merge", "tile_features", "write_sidecar", "read_sidecar"]


def _stream(t: torch.Tensor):
    return torch.cuda.current_stream(t.device).cuda_stream


def mask_contours(bits: torch.Tensor, w: int, max_pts: int = 256, check: bool = True, bbox: Optional[torch.Tensor] = None,
                  select: Optional[torch.Tensor] = None):
    """bits [n,h,ceil(w/64)] int64 bit rows (RoIStageResult.mask_bits / mask_nms.pack_masks) -> (xy [n,max_pts,2] int32,
    count [n] int32, status [1] int32), all on the device.  ``bbox`` [n,4] int32: the tight boxes paste / pack return
    (optional, saves a scan).  ``select`` [n] uint8/bool: only these masks are traced (the others get count 0).  ``check``
    reads the status word (one small D2H)."""
    if not bits.is_cuda:
        raise NuhtcError("mask_contours: CUDA tensors only (no CPU fallback)")
    assert bits.dtype == torch.int64 and bits.dim() == 3 and bits.is_contiguous()
    n, h, wpm = bits.shape
    assert wpm == (w + 63) // 64
    xy = torch.empty((n, max_pts, 2), dtype=torch.int32, device=bits.device)
    cnt = torch.empty((n,), dtype=torch.int32, device=bits.device)
    status = torch.empty((1,), dtype=torch.int32, device=bits.device)
    if bbox is not None:
        assert bbox.dtype == torch.int32 and bbox.shape == (n, 4) and bbox.is_contiguous() and bbox.device == bits.device
    if select is not None:
        if select.dtype == torch.bool:
            select = select.view(torch.uint8)
        assert select.dtype == torch.uint8 and select.shape == (n,) and select.is_contiguous() and select.device == bits.device
    with torch.cuda.device(bits.device):
        rc = lib().nuhtc_mask_contours(bits.data_ptr(), 0 if bbox is None else bbox.data_ptr(),
                                       0 if select is None else select.data_ptr(), n, h, w, max_pts, xy.data_ptr(), cnt.data_ptr(), status.data_ptr(),
                                       _stream(bits))
    _lib.check(rc, "nuhtc_mask_contours")
    count("contours")
    if check:
        st = int(status.item())
        if st == 1:
            raise NuhtcError(f"mask_contours: a contour has more than max_pts={max_pts} points (max {int(cnt.max())})")
        if st:
            raise NuhtcError("mask_contours: a mask larger than 64 px in a frame that does not fit shared memory")
    return xy, cnt, status


def mask2inst(inst_map) -> np.ndarray:
    """tools/infer_wsi.py:51-54 for one dense mask: [n+1,1,2] int32 contour, first point repeated at the end."""
    from .mask_nms import pack_masks
    m = torch.as_tensor(np.ascontiguousarray(inst_map)).to(torch.uint8).cuda()[None]
    bits, _, bbox = pack_masks(m)
    xy, cnt, _ = mask_contours(bits, m.shape[2], bbox=bbox, max_pts=max(8, 2 * int(m.shape[1] + m.shape[2]) + int(m.sum().item())))
    c = xy[0, : int(cnt[0])].cpu().numpy().reshape(-1, 1, 2)
    return np.concatenate([c, c[[0]]], axis=0)


def rings_for_merge(xy: torch.Tensor, cnt: torch.Tensor, select: Optional[torch.Tensor] = None,
                    origin: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Closed fp64 rings of the selected masks, shifted by their tile coordinate.

    select [n] bool (None = all); masks whose closed contour has fewer than 3 points are dropped like infer_wsi.py:530.
    origin [n,2] int32 tile (x, y) per mask.  Returns (ring_xy [sumV,2] fp64, voff [m+1] int64, index [m] int64 = the
    selected masks' indices).  Synchronises (sizes the output)."""
    n, max_pts, _ = xy.shape
    ok = (cnt + 1) >= 3
    if select is not None:
        ok = ok & select
    k = torch.where(ok, cnt.to(torch.int64) + 1, torch.zeros_like(cnt, dtype=torch.int64))
    voff_all = torch.zeros((n + 1,), dtype=torch.int64, device=xy.device)
    torch.cumsum(k, 0, out=voff_all[1:])
    index = ok.nonzero().squeeze(1)
    total = int(voff_all[-1].item())
    out = torch.empty((total, 2), dtype=torch.float64, device=xy.device)
    if origin is not None:
        origin = origin.to(torch.int32).contiguous()
    with torch.cuda.device(xy.device):
        rc = lib().nuhtc_contour_rings(xy.data_ptr(), cnt.data_ptr(), voff_all.data_ptr(), 0 if origin is None else origin.data_ptr(),
                                       n, max_pts, out.data_ptr(), _stream(xy))
    _lib.check(rc, "nuhtc_contour_rings")
    count("rings")
    voff = torch.cat([voff_all[index], voff_all[-1:]])
    return out, voff, index


def tile_features(ring_xy: np.ndarray, voff: np.ndarray, boxes: np.ndarray, labels: np.ndarray, scores: np.ndarray,
                  classes: Sequence[str], colors: Sequence[Sequence[int]]) -> Tuple[List[dict], List[dict]]:
    """The QuPath wire format of tools/infer_wsi.py:541-585: (polygon features, centre-point features).  ``boxes`` are
    already in slide coordinates (infer_wsi.py:536)."""
    geo, pts = [], []
    for i in range(len(voff) - 1):
        ring = ring_xy[voff[i]: voff[i + 1]]
        ring = ring.astype(np.int64) if np.all(ring == np.round(ring)) else ring
        props = lambda: {"objectType": "annotation", "label": int(labels[i]), "score": float(scores[i]),
                         "classification": {"name": classes[int(labels[i])], "color": list(colors[int(labels[i])])},
                         "isLocked": False}
        geo.append({"type": "Feature", "geometry": {"type": "Polygon", "coordinates": [ring.tolist()]}, "properties": props()})
        pts.append({"type": "Feature",
                    "geometry": {"type": "Point", "coordinates": [float((boxes[i][0] + boxes[i][2]) / 2),
                                                                  float((boxes[i][1] + boxes[i][3]) / 2)]},
                    "properties": props()})
    return geo, pts


def write_sidecar(path: str, ring_xy: np.ndarray, voff: np.ndarray, scores: np.ndarray, labels: np.ndarray,
                  boxes: Optional[np.ndarray] = None) -> None:
    """Binary sidecar of a slide's nuclei (vertices / offsets / score / label arrays): what `nuclei_merge.merge_arrays` reads
    directly, instead of parsing a GeoJSON of 10^6 features."""
    np.savez(path, xy=np.asarray(ring_xy, dtype=np.float64), voff=np.asarray(voff, dtype=np.int64),
             score=np.asarray(scores, dtype=np.float64), label=np.asarray(labels, dtype=np.int64),
             bbox=np.zeros((0, 4), np.float32) if boxes is None else np.asarray(boxes, dtype=np.float32))


def read_sidecar(path: str):
    z = np.load(path)
    return z["xy"], z["voff"], z["score"], z["label"], z["bbox"]
