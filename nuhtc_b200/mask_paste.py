"""Mask paste: drop-in for mmdet's ``_do_paste_mask`` / ``FCNMaskHead.get_seg_masks`` hot part.

Reference: /root/reference/thirdparty/mmdetection/mmdet/models/roi_heads/mask_heads/fcn_mask_head.py
  * ``_do_paste_mask(masks, boxes, img_h, img_w, skip_empty=True)``  :344-412
  * ``get_seg_masks(mask_pred, det_bboxes, det_labels, rcnn_test_cfg, ori_shape, scale_factor, rescale)``  :179-310
The kernel fuses grid generation, bilinear resampling, the ``>= thr`` and the narrow store; see csrc/paste.cu.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib as L

__all__ = ["paste_masks_dense_bits", "_do_paste_mask", "paste_masks", "get_seg_masks", "get_seg_masks_device"]


def paste_masks(masks: torch.Tensor, boxes: torch.Tensor, img_h: int, img_w: int, thr: Optional[float] = None,
                kind: str = "prob", want_stats: bool = False):
    """masks [N,1,h,w] or [N,h,w] fp32 probabilities, boxes [N,>=4] -> pasted tensor.

    kind 'prob' : fp32 [N,img_h,img_w] (what ``_do_paste_mask`` returns)
    kind 'bin'  : bool [N,img_h,img_w] = prob >= thr (what ``get_seg_masks`` builds)
    kind 'bits' : int64 [N,img_h,ceil(img_w/64)] bit rows of the same binary mask (8x fewer bytes)
    want_stats  : also return (area [N] int32, bbox [N,4] int32) of the binary mask."""
    L.require_cuda(masks, "masks")
    L.require_cuda(boxes, "boxes")
    if masks.dim() == 4:
        assert masks.size(1) == 1, "class-agnostic mask [N,1,h,w] expected (select the label channel first)"
        masks = masks[:, 0]
    N, mh, mw = masks.shape
    dev = masks.device
    masks = masks.to(torch.float32).contiguous()
    boxes = boxes[:, :4].to(torch.float32).contiguous()
    img_h, img_w = int(img_h), int(img_w)
    if kind == "prob":
        out = torch.empty((N, img_h, img_w), dtype=torch.float32, device=dev)
        k = L.PASTE_PROB
    elif kind == "bin":
        out = torch.empty((N, img_h, img_w), dtype=torch.bool, device=dev)
        k = L.PASTE_BIN
    elif kind == "bits":
        out = torch.empty((N, img_h, (img_w + 63) // 64), dtype=torch.int64, device=dev)
        k = L.PASTE_BITS
    else:
        raise ValueError(kind)
    if kind != "prob" and thr is None:
        raise ValueError("a threshold is required for binary output")
    area = bbox = None
    if want_stats and kind != "prob":
        area = torch.empty(N, dtype=torch.int32, device=dev)
        bbox = torch.empty((N, 4), dtype=torch.int32, device=dev)
    if N:
        with torch.cuda.device(dev):
            rc = L.lib().nuhtc_paste_masks(masks.data_ptr(), boxes.data_ptr(), N, mh, mw, img_h, img_w,
                                           float(0.0 if thr is None else thr), k, out.data_ptr(), L.ptr(area), L.ptr(bbox),
                                           L.stream_ptr(dev))
        L.check(rc, "paste_masks")
        L.count("paste")
    return (out, area, bbox) if want_stats else out


def paste_masks_dense_bits(masks: torch.Tensor, boxes: torch.Tensor, img_h: int, img_w: int, thr: float):
    """Dense bool frames AND bit rows (+ area, tight box) from one evaluation of every mask (img_w % 16 == 0):
    returns (dense bool [N,H,W], bits int64 [N,H,ceil(W/64)], area int32 [N], bbox int32 [N,4])."""
    L.require_cuda(masks, "masks")
    L.require_cuda(boxes, "boxes")
    if masks.dim() == 4:
        assert masks.size(1) == 1, "class-agnostic mask [N,1,h,w] expected (select the label channel first)"
        masks = masks[:, 0]
    N, mh, mw = masks.shape
    dev = masks.device
    masks = masks.to(torch.float32).contiguous()
    boxes = boxes[:, :4].to(torch.float32).contiguous()
    img_h, img_w = int(img_h), int(img_w)
    dense = torch.empty((N, img_h, img_w), dtype=torch.bool, device=dev)
    bits = torch.empty((N, img_h, (img_w + 63) // 64), dtype=torch.int64, device=dev)
    area = torch.empty(N, dtype=torch.int32, device=dev)
    bbox = torch.empty((N, 4), dtype=torch.int32, device=dev)
    if N:
        with torch.cuda.device(dev):
            rc = L.lib().nuhtc_paste_masks_dense_bits(masks.data_ptr(), boxes.data_ptr(), N, mh, mw, img_h, img_w, float(thr),
                                                      dense.data_ptr(), bits.data_ptr(), area.data_ptr(), bbox.data_ptr(),
                                                      L.stream_ptr(dev))
        L.check(rc, "paste_masks_dense_bits")
        L.count("paste_dual")
    return dense, bits, area, bbox


def _do_paste_mask(masks: torch.Tensor, boxes: torch.Tensor, img_h: int, img_w: int, skip_empty: bool = True):
    """Same contract as the reference: returns (pasted fp32 [N,h',w'], (slice_y, slice_x) or ()).

    With ``skip_empty`` the reference pastes only the region that tightly bounds all boxes (+1 px) and
    returns its slices; the values inside are identical to the full-frame paste, so the region is cut
    out of the full-frame result here."""
    full = paste_masks(masks, boxes, img_h, img_w, kind="prob")
    if not skip_empty:
        return full, ()
    x0_int, y0_int = torch.clamp(boxes.min(dim=0).values.floor()[:2] - 1, min=0).to(dtype=torch.int32)
    x1_int = torch.clamp(boxes[:, 2].max().ceil() + 1, max=img_w).to(dtype=torch.int32)
    y1_int = torch.clamp(boxes[:, 3].max().ceil() + 1, max=img_h).to(dtype=torch.int32)
    x0_int, y0_int, x1_int, y1_int = int(x0_int), int(y0_int), int(x1_int), int(y1_int)
    return full[:, y0_int:y1_int, x0_int:x1_int], (slice(y0_int, y1_int), slice(x0_int, x1_int))


def _paste_frame(det_bboxes, ori_shape, scale_factor, rescale) -> Tuple[torch.Tensor, int, int]:
    bboxes = det_bboxes[:, :4]
    if not isinstance(scale_factor, torch.Tensor):
        if isinstance(scale_factor, float):
            scale_factor = np.array([scale_factor] * 4)
        assert isinstance(scale_factor, np.ndarray)
        scale_factor = torch.Tensor(scale_factor)
    if rescale:
        img_h, img_w = ori_shape[:2]
        bboxes = bboxes / scale_factor.to(bboxes)
    else:
        w_scale, h_scale = scale_factor[0], scale_factor[1]
        img_h = np.round(ori_shape[0] * h_scale.item()).astype(np.int32)
        img_w = np.round(ori_shape[1] * w_scale.item()).astype(np.int32)
    return bboxes, int(img_h), int(img_w)


def get_seg_masks_device(mask_pred, det_bboxes, det_labels, mask_thr_binary, ori_shape, scale_factor, rescale,
                         class_agnostic: bool = True, kind: str = "bin", want_stats: bool = False):
    """Device-resident core of get_seg_masks: returns the [N,H,W] bool (or bit-row) tensor without the
    per-instance host copies of fcn_mask_head.py:308-309."""
    if isinstance(mask_pred, torch.Tensor):
        mask_pred = mask_pred.sigmoid()
    else:
        mask_pred = det_bboxes.new_tensor(mask_pred)  # AugTest branch: already activated (fcn_mask_head.py:228-232)
    bboxes, img_h, img_w = _paste_frame(det_bboxes, ori_shape, scale_factor, rescale)
    if not class_agnostic:
        mask_pred = mask_pred[range(len(mask_pred)), det_labels][:, None]
    if mask_thr_binary < 0:
        raise NotImplementedError("mask_thr_binary < 0 (debug visualisation) is outside the NuHTC inference path")
    return paste_masks(mask_pred, bboxes, img_h, img_w, thr=mask_thr_binary, kind=kind, want_stats=want_stats)


def get_seg_masks(mask_pred, det_bboxes, det_labels, rcnn_test_cfg, ori_shape, scale_factor, rescale, num_classes: int,
                  class_agnostic: bool = True):
    """FCNMaskHead.get_seg_masks: list[num_classes] of lists of host bool arrays [H,W]."""
    thr = rcnn_test_cfg["mask_thr_binary"] if isinstance(rcnn_test_cfg, dict) else rcnn_test_cfg.mask_thr_binary
    im_mask = get_seg_masks_device(mask_pred, det_bboxes, det_labels, thr, ori_shape, scale_factor, rescale, class_agnostic)
    host = im_mask.cpu().numpy()  # one D2H copy instead of N
    labels = det_labels.cpu().numpy()
    cls_segms = [[] for _ in range(num_classes)]
    for i in range(host.shape[0]):
        cls_segms[labels[i]].append(host[i])
    return cls_segms
