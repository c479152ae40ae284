"""CPU restatement of the RoI stage as the reference drives it  --  TEST INFRASTRUCTURE ONLY.

Follows stock ``HybridTaskCascadeRoIHead.simple_test`` (thirdparty/mmdetection/mmdet/models/roi_heads/htc_roi_head.py:330-503)
with ``SingleRoIExtractor`` + the per-tile post-processing of tools/infer_wsi.py:510-526: per-image Python loops,
per-level RoIAlign calls with scatter, per-image multiclass_nms, per-image get_seg_masks, per-tile mask_nms.
Used by the parity tests and as the timed CPU baseline of bench.py (the mmcv CPU kernels are single threaded;
``nthreads`` > 1 lets the oracle split RoIs / masks over host threads for the "all host threads" reference arm).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Sequence

import numpy as np
import torch

from . import cpu as O


def roi_stage_cpu(feats: Sequence[torch.Tensor], rois: torch.Tensor, bbox_heads: Sequence[Callable], mask_head: Callable,
                  cfg, nthreads: int = 1, score_fn=None) -> List[Dict]:
    """Returns one dict per tile: det_boxes [d,4], det_scores, det_labels, masks [d,H,W] bool, keep (indices into d,
    mask-NMS survivors in score order)."""
    B = feats[0].shape[0]
    C = cfg.num_classes
    score_fn = score_fn or (lambda s: torch.softmax(s, dim=-1))

    def extract(r, P, sr):
        if cfg.extractor == "single":
            return O.single_roi_extract(feats[: len(cfg.featmap_strides)], r, cfg.featmap_strides, P, sr, cfg.finest_scale, nthreads)
        return O.sum_roi_extract(feats[: cfg.sum_levels], r, cfg.featmap_strides[: cfg.sum_levels], P, sr, nthreads)

    ms_scores = []
    bbox_pred = None
    for i in range(cfg.num_stages):
        bbox_feats = extract(rois, cfg.bbox_out, cfg.bbox_sampling_ratio)
        cls_score, bbox_pred = bbox_heads[i](bbox_feats)
        ms_scores.append(cls_score)
        if i < cfg.num_stages - 1:
            new = O.delta2bbox(rois[:, 1:], bbox_pred, stds=cfg.stage_stds[i], max_shape=cfg.img_shape)
            rois = torch.cat([rois[:, :1], new], dim=1)
    cls_score = sum(ms_scores) / float(len(ms_scores))
    scores = score_fn(cls_score)
    bboxes = O.delta2bbox(rois[:, 1:], bbox_pred, stds=cfg.stage_stds[-1], max_shape=cfg.img_shape) / cfg.scale_factor
    tile_of_roi = rois[:, 0].long()
    out = []
    H, W = cfg.ori_shape
    for b in range(B):
        idx = (tile_of_roi == b).nonzero().squeeze(1)
        sc = torch.cat([scores[idx, :C], scores.new_zeros(idx.numel(), 1)], dim=1)  # last column = background, ignored
        dets, labels, cand = O.multiclass_nms(bboxes[idx], sc, cfg.score_thr, dict(type="nms", iou_threshold=cfg.nms_iou),
                                              cfg.max_per_img)
        d = dets.shape[0]
        det_boxes = dets[:, :4].contiguous()
        mask_rois = torch.cat([torch.full((d, 1), float(b)), det_boxes * cfg.scale_factor], dim=1)
        mask_feats = extract(mask_rois, cfg.mask_out, cfg.mask_sampling_ratio)
        # global candidate id of each detection (roi index * C + label), what the GPU driver hands its mask head
        det_cand = idx[cand // C] * C + (cand % C)
        logits = mask_head(mask_feats, det_cand)
        masks = O.get_seg_masks(logits.sigmoid(), det_boxes, H, W, np.array([1.0] * 4, dtype=np.float32), True, cfg.mask_thr_binary)
        area = masks.sum((1, 2))
        ok = ((det_boxes[:, 0] >= cfg.margin) & (det_boxes[:, 1] >= cfg.margin) & (det_boxes[:, 2] <= W - cfg.margin) &
              (det_boxes[:, 3] <= H - cfg.margin) & (area >= cfg.min_area))
        sel = ok.nonzero().squeeze(1)
        if sel.numel():
            k = O.mask_nms(masks[sel].numpy().astype(np.uint8), dets[sel, 4].numpy(), thr=cfg.mask_nms_thr)
            keep = sel[torch.from_numpy(np.ascontiguousarray(k))]
        else:
            keep = sel
        # tools/infer_wsi.py:528-533: contour of every kept nucleus, contours shorter than 3 points dropped
        mk = masks.numpy().astype(np.uint8)
        contours = [O.mask2inst(mk[int(i)]) for i in keep]
        out.append(dict(det_boxes=det_boxes, det_scores=dets[:, 4], det_labels=labels, det_cand=det_cand, masks=masks, keep=keep,
                        contours=contours, mask_prob=logits.sigmoid()))
    return out
