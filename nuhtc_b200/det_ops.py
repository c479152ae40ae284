"""Detection glue ops over the C ABI (det_glue.cu): box decode, multiclass candidate expansion, fixed detection slots and
the tile margin / min-area filter.  One launch each instead of the chains of small torch kernels the reference runs
(mmdet delta_xywh_bbox_coder.py:163-260, nuhtc/models/bbox_head.py:12-102, tools/infer_wsi.py:510-521)."""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import torch

from . import _lib as L

__all__ = ["delta2bbox", "multiclass_candidates", "detection_slots", "tile_filter", "keep_flags"]

_F4 = ctypes.c_float * 4


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    L.require_cuda(t, name)
    assert t.dtype == torch.float32, f"{name}: float32 expected"
    return t.contiguous()


def delta2bbox(rois: torch.Tensor, deltas: torch.Tensor, means: Sequence[float] = (0., 0., 0., 0.),
               stds: Sequence[float] = (1., 1., 1., 1.), max_shape=None, wh_ratio_clip: float = 16 / 1000,
               divide_by: float = 1.0) -> torch.Tensor:
    """Class-agnostic ``delta2bbox`` (delta_xywh_bbox_coder.py:163-260).  rois [K,4] or [K,5] (batch index in column 0,
    copied through); returns the same shape.  ``divide_by``: the rescale of get_bboxes (bbox_head.py:373-376)."""
    rois, deltas = _f32c(rois, "rois"), _f32c(deltas, "deltas")
    K = rois.shape[0]
    assert rois.dim() == 2 and rois.shape[1] in (4, 5) and deltas.shape == (K, 4)
    out = torch.empty_like(rois)
    mh, mw = (int(max_shape[0]), int(max_shape[1])) if max_shape is not None else (0, 0)
    with torch.cuda.device(rois.device):
        rc = L.lib().nuhtc_delta2bbox(rois.data_ptr(), int(rois.shape[1] == 5), deltas.data_ptr(), K, _F4(*[float(m) for m in means]),
                                      _F4(*[float(s) for s in stds]), mh, mw, float(wh_ratio_clip), float(divide_by),
                                      out.data_ptr(), L.stream_ptr(rois.device))
    L.check(rc, "nuhtc_delta2bbox")
    L.count("glue")
    return out


def multiclass_candidates(boxes: torch.Tensor, scores: torch.Tensor, rois: torch.Tensor, num_classes: int, score_thr: float):
    """boxes [K,4] (may be a column slice of a [K,5] tensor), scores [K,>=num_classes] (background last), rois [K,5]
    (column 0 = tile) -> (cand_boxes [K*C,4], cand_scores, cand_labels int64, cand_tile int32, groups int32 = tile or -1 below
    the threshold)."""
    scores, rois = _f32c(scores, "scores"), _f32c(rois, "rois")
    L.require_cuda(boxes, "boxes")
    assert boxes.dtype == torch.float32 and boxes.dim() == 2 and boxes.shape[1] == 4
    if boxes.stride(1) != 1:
        boxes = boxes.contiguous()
    K, C = boxes.shape[0], num_classes
    dev = boxes.device
    cb = torch.empty((K * C, 4), dtype=torch.float32, device=dev)
    cs = torch.empty((K * C,), dtype=torch.float32, device=dev)
    cl = torch.empty((K * C,), dtype=torch.int64, device=dev)
    ct = torch.empty((K * C,), dtype=torch.int32, device=dev)
    gr = torch.empty((K * C,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = L.lib().nuhtc_multiclass_candidates(boxes.data_ptr(), boxes.stride(0), scores.data_ptr(), scores.shape[1], rois.data_ptr(), rois.shape[1],
                                                 K, C, float(score_thr), cb.data_ptr(), cs.data_ptr(), cl.data_ptr(), ct.data_ptr(),
                                                 gr.data_ptr(), L.stream_ptr(dev))
    L.check(rc, "nuhtc_multiclass_candidates")
    L.count("glue")
    return cb, cs, cl, ct, gr


def detection_slots(keep: torch.Tensor, gstart: torch.Tensor, gcount: torch.Tensor, max_per_img: int, cand_boxes: torch.Tensor,
                    cand_scores: torch.Tensor, cand_labels: torch.Tensor, cand_tile: torch.Tensor, scale_factor: float = 1.0):
    """``dets[:max_num]`` per tile as fixed slots -> (det_boxes [B*M,4], det_scores, det_labels, det_tile, det_valid bool,
    det_cand int64, mask_rois [B*M,5])."""
    dev = keep.device
    B, M = gstart.numel(), int(max_per_img)
    n = B * M
    db = torch.empty((n, 4), dtype=torch.float32, device=dev)
    ds = torch.empty((n,), dtype=torch.float32, device=dev)
    dl = torch.empty((n,), dtype=torch.int64, device=dev)
    dt = torch.empty((n,), dtype=torch.int32, device=dev)
    dv = torch.empty((n,), dtype=torch.uint8, device=dev)
    dc = torch.empty((n,), dtype=torch.int64, device=dev)
    mr = torch.empty((n, 5), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = L.lib().nuhtc_detection_slots(keep.data_ptr(), gstart.data_ptr(), gcount.data_ptr(), B, M, cand_boxes.data_ptr(),
                                           cand_scores.data_ptr(), cand_labels.data_ptr(), cand_tile.data_ptr(), float(scale_factor),
                                           db.data_ptr(), ds.data_ptr(), dl.data_ptr(), dt.data_ptr(), dv.data_ptr(), dc.data_ptr(),
                                           mr.data_ptr(), L.stream_ptr(dev))
    L.check(rc, "nuhtc_detection_slots")
    L.count("glue")
    return db, ds, dl, dt, dv.view(torch.bool), dc, mr


def tile_filter(det_boxes: torch.Tensor, area: torch.Tensor, det_tile: torch.Tensor, margin: int, img_h: int, img_w: int,
                min_area: int) -> torch.Tensor:
    """tools/infer_wsi.py:510-521: tile id of the detections inside the margin with at least min_area pixels, -1 otherwise."""
    out = torch.empty_like(det_tile)
    with torch.cuda.device(det_boxes.device):
        rc = L.lib().nuhtc_tile_filter(det_boxes.data_ptr(), area.data_ptr(), det_tile.data_ptr(), det_boxes.shape[0], int(margin),
                                       int(img_h), int(img_w), int(min_area), out.data_ptr(), L.stream_ptr(det_boxes.device))
    L.check(rc, "nuhtc_tile_filter")
    L.count("glue")
    return out


def keep_flags(keep: torch.Tensor, tile_start: torch.Tensor, tile_count: torch.Tensor, max_tile_size: int, n: int) -> torch.Tensor:
    """uint8 [n]: 1 for the masks listed in the mask-NMS keep lists, 0 elsewhere (infer_wsi.py:527-531 selection)."""
    flags = torch.empty((n,), dtype=torch.uint8, device=keep.device)
    with torch.cuda.device(keep.device):
        rc = L.lib().nuhtc_keep_flags(keep.data_ptr(), tile_start.data_ptr(), tile_count.data_ptr(), tile_start.numel(),
                                      int(max_tile_size), n, flags.data_ptr(), L.stream_ptr(keep.device))
    L.check(rc, "nuhtc_keep_flags")
    L.count("glue")
    return flags
