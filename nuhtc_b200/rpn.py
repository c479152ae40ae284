"""RPN proposal post-processing: per-level top-k, delta2bbox, batched NMS with the level id as the
label, top ``max_per_img`` -- same flow as RPNHead._get_bboxes_single / _bbox_post_process
(/root/reference/thirdparty/mmdetection/mmdet/models/dense_heads/rpn_head.py:103-236), on the B200 NMS kernels.

``get_bboxes_single`` / ``bbox_post_process`` keep the reference's per-image semantics; ``proposals_batched`` runs all
images of a batch through ONE grouped NMS launch (image = group, level = class segment).
"""
from __future__ import annotations

import copy
from typing import List, Sequence

import torch

import ctypes
import os

from . import _lib as L
from .mmcv_ops import batched_nms, nms_groups
from .det_ops import delta2bbox

__all__ = ["get_bboxes_single", "bbox_post_process", "proposals_batched"]


def _cfg_get(cfg, key, default=None):
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


def _level_topk(cls_score: torch.Tensor, bbox_pred: torch.Tensor, anchors: torch.Tensor, nms_pre: int, use_sigmoid_cls: bool):
    assert cls_score.size()[-2:] == bbox_pred.size()[-2:]
    cls_score = cls_score.permute(1, 2, 0)
    if use_sigmoid_cls:
        scores = cls_score.reshape(-1).sigmoid()
    else:
        scores = cls_score.reshape(-1, 2).softmax(dim=1)[:, 0]
    bbox_pred = bbox_pred.permute(1, 2, 0).reshape(-1, 4)
    if 0 < nms_pre < scores.shape[0]:
        ranked_scores, rank_inds = scores.sort(descending=True)
        topk_inds = rank_inds[:nms_pre]
        scores = ranked_scores[:nms_pre]
        bbox_pred = bbox_pred[topk_inds, :]
        anchors = anchors[topk_inds, :]
    return scores, bbox_pred, anchors


def bbox_post_process(mlvl_scores, mlvl_bboxes, mlvl_valid_anchors, level_ids, cfg, img_shape):
    """RPNHead._bbox_post_process (rpn_head.py:189-236): returns dets [n,5]."""
    scores = torch.cat(mlvl_scores)
    anchors = torch.cat(mlvl_valid_anchors)
    rpn_bbox_pred = torch.cat(mlvl_bboxes)
    proposals = delta2bbox(anchors.contiguous(), rpn_bbox_pred.contiguous(), stds=(1., 1., 1., 1.), max_shape=img_shape)
    ids = torch.cat(level_ids)
    min_bbox_size = _cfg_get(cfg, "min_bbox_size", -1)
    if min_bbox_size >= 0:
        w = proposals[:, 2] - proposals[:, 0]
        h = proposals[:, 3] - proposals[:, 1]
        valid_mask = (w > min_bbox_size) & (h > min_bbox_size)
        if not valid_mask.all():
            proposals, scores, ids = proposals[valid_mask], scores[valid_mask], ids[valid_mask]
    if proposals.numel() > 0:
        dets, _ = batched_nms(proposals, scores, ids, _cfg_get(cfg, "nms"))
    else:
        return proposals.new_zeros(0, 5)
    return dets[:_cfg_get(cfg, "max_per_img")]


def get_bboxes_single(cls_score_list: Sequence[torch.Tensor], bbox_pred_list: Sequence[torch.Tensor],
                      mlvl_anchors: Sequence[torch.Tensor], img_shape, cfg, use_sigmoid_cls: bool = True):
    """RPNHead._get_bboxes_single (rpn_head.py:103-187) for one image."""
    cfg = copy.deepcopy(cfg)
    nms_pre = _cfg_get(cfg, "nms_pre", -1)
    level_ids, mlvl_scores, mlvl_bbox_preds, mlvl_valid_anchors = [], [], [], []
    for level_idx in range(len(cls_score_list)):
        scores, bbox_pred, anchors = _level_topk(cls_score_list[level_idx], bbox_pred_list[level_idx], mlvl_anchors[level_idx],
                                                 nms_pre, use_sigmoid_cls)
        mlvl_scores.append(scores)
        mlvl_bbox_preds.append(bbox_pred)
        mlvl_valid_anchors.append(anchors)
        level_ids.append(scores.new_full((scores.size(0),), level_idx, dtype=torch.long))
    return bbox_post_process(mlvl_scores, mlvl_bbox_preds, mlvl_valid_anchors, level_ids, cfg, img_shape)


@torch.no_grad()
def proposals_batched(cls_scores: Sequence[torch.Tensor], bbox_preds: Sequence[torch.Tensor], mlvl_anchors: Sequence[torch.Tensor],
                      img_shape, cfg, use_sigmoid_cls: bool = True, impl: str = "auto") -> List[torch.Tensor]:
    """All B images at once: ``cls_scores[l]`` is [B, A(*2), H_l, W_l], ``bbox_preds[l]`` [B, A*4, H_l, W_l].
    Same result per image as ``get_bboxes_single`` (below mmcv's split_thr the NMS runs on offset boxes; decoded boxes are
    clamped to the frame, so the level segments are exact), one NMS launch for the batch.
    impl: 'auto' = the fused top-k + decode kernel (`nuhtc_rpn_topk_decode`, one launch for every level and image) when the
    levels fit its shared-memory select, else the per-level torch sort; 'sort' forces the latter (A/B, parity tests)."""
    B = cls_scores[0].shape[0]
    nms_cfg = dict(_cfg_get(cfg, "nms"))
    assert nms_cfg.pop("type", "nms") == "nms"
    iou_thr = nms_cfg.pop("iou_threshold")
    nms_pre = _cfg_get(cfg, "nms_pre", -1)
    max_per_img = _cfg_get(cfg, "max_per_img")
    min_bbox_size = _cfg_get(cfg, "min_bbox_size", -1)
    nl = len(cls_scores)
    dev = cls_scores[0].device
    fused = _topk_decode(cls_scores, bbox_preds, mlvl_anchors, img_shape, nms_pre, min_bbox_size, use_sigmoid_cls, impl)
    if fused is not None:
        boxes, scores, labels, groups, per_image, per_level = fused
        return _nms_and_split(boxes, scores, labels, groups, B, nl, per_image, per_level, iou_thr, nms_cfg, max_per_img)
    boxes, scores, labels, groups = [], [], [], []
    per_image = per_level = 0
    img = torch.arange(B, device=dev, dtype=torch.int32)
    for l in range(nl):  # one sort / gather / decode per LEVEL for the whole batch (the reference loops images x levels)
        cs = cls_scores[l].permute(0, 2, 3, 1)
        if use_sigmoid_cls:
            s = cs.reshape(B, -1).sigmoid()
        else:
            s = cs.reshape(B, -1, 2).softmax(dim=2)[:, :, 0]
        d = bbox_preds[l].permute(0, 2, 3, 1).reshape(B, -1, 4)
        a = mlvl_anchors[l]
        n = s.shape[1]
        if 0 < nms_pre < n:
            s, idx = s.sort(dim=1, descending=True)
            s, idx = s[:, :nms_pre], idx[:, :nms_pre]
            d = torch.gather(d, 1, idx[:, :, None].expand(B, nms_pre, 4))
            a = a[idx]                                                   # [B,k,4]
        else:
            a = a[None].expand(B, n, 4)
        k = s.shape[1]
        boxes.append(delta2bbox(a.reshape(-1, 4).contiguous(), d.reshape(-1, 4).contiguous(), stds=(1., 1., 1., 1.),
                                max_shape=img_shape).view(B, k, 4))
        scores.append(s)
        labels.append(torch.full((B, k), l, dtype=torch.long, device=dev))
        groups.append(img[:, None].expand(B, k))
        per_image += k
        per_level = max(per_level, k)
    # image-major order: the NMS breaks score ties by index, like the per-image call on the level-concatenated candidates
    boxes = torch.cat(boxes, 1).reshape(-1, 4)
    scores = torch.cat(scores, 1).reshape(-1)
    labels = torch.cat(labels, 1).reshape(-1)
    groups = torch.cat(groups, 1).reshape(-1).contiguous()
    if min_bbox_size >= 0:
        ok = ((boxes[:, 2] - boxes[:, 0]) > min_bbox_size) & ((boxes[:, 3] - boxes[:, 1]) > min_bbox_size)
        groups = torch.where(ok, groups, torch.full_like(groups, -1))
    return _nms_and_split(boxes, scores, labels, groups, B, nl, per_image, per_level, iou_thr, nms_cfg, max_per_img)


def _nms_and_split(boxes, scores, labels, groups, B, nl, per_image, per_level, iou_thr, nms_cfg, max_per_img):
    split_thr = nms_cfg.pop("split_thr", 10000)
    mode = "offset" if per_image < split_thr else "perclass"  # NB: mmcv decides on the per-image candidate count AFTER the size filter
    keep, gstart, gcount, status = nms_groups(boxes, scores, labels, groups, B, per_level, iou_thr, 0, mode, num_classes=nl)
    host = torch.stack([gstart, torch.clamp(gcount, max=max_per_img), status.to(gstart.dtype).expand_as(gstart)]).cpu()   # one D2H
    if int(host[2, 0]) != 0:
        raise RuntimeError(f"rpn nms status {int(host[2, 0])}")
    out = []
    for b in range(B):
        k = keep[int(host[0, b]): int(host[0, b]) + int(host[1, b])]
        out.append(torch.cat([boxes[k], scores[k, None]], dim=1))
    return out


def _topk_decode(cls_scores, bbox_preds, mlvl_anchors, img_shape, nms_pre, min_bbox_size, use_sigmoid_cls, impl):
    """The fused pre-selection: None when it does not apply (softmax scores, levels too large, impl='sort')."""
    if impl == "sort" or not use_sigmoid_cls or os.environ.get("NUHTC_RPN_TOPK") == "0":
        return None
    nl = len(cls_scores)
    B, A = int(cls_scores[0].shape[0]), int(cls_scores[0].shape[1])
    dev = cls_scores[0].device
    L.require_cuda(cls_scores[0], "cls_scores")
    Hs = (ctypes.c_int * nl)(*[int(c.shape[2]) for c in cls_scores])
    Ws = (ctypes.c_int * nl)(*[int(c.shape[3]) for c in cls_scores])
    lib = L.lib()
    if not lib.nuhtc_rpn_topk_supported(Hs, Ws, nl, A, int(nms_pre)):
        return None
    cls = [c.contiguous() for c in cls_scores]
    reg = [r.contiguous() for r in bbox_preds]
    anc = [a.contiguous() for a in mlvl_anchors]
    for c, r, a in zip(cls, reg, anc):
        assert c.dtype == r.dtype == a.dtype == torch.float32 and r.shape[1] == 4 * A and a.shape[0] == c.shape[2] * c.shape[3] * A
    ks = [min(int(nms_pre), int(a.shape[0])) if nms_pre > 0 else int(a.shape[0]) for a in anc]
    per_image, per_level = sum(ks), max(ks)
    boxes = torch.empty((B * per_image, 4), dtype=torch.float32, device=dev)
    scores = torch.empty(B * per_image, dtype=torch.float32, device=dev)
    labels = torch.empty(B * per_image, dtype=torch.int64, device=dev)
    groups = torch.empty(B * per_image, dtype=torch.int32, device=dev)
    ptr = lambda ts: (ctypes.c_void_p * nl)(*[t.data_ptr() for t in ts])
    with torch.cuda.device(dev):
        rc = lib.nuhtc_rpn_topk_decode(ptr(cls), ptr(reg), ptr(anc), Hs, Ws, nl, B, A, int(nms_pre), 1, int(img_shape[0]), int(img_shape[1]),
                                       16 / 1000, float(min_bbox_size), boxes.data_ptr(), scores.data_ptr(), labels.data_ptr(),
                                       groups.data_ptr(), L.stream_ptr(dev))
    L.check(rc, "rpn_topk_decode")
    L.count("rpn_topk")
    return boxes, scores, labels, groups, per_image, per_level
