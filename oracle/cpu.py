"""CPU oracle for the NuHTC RoI stage + merge  --  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module.  The product
package ``nuhtc_b200`` never does (tests/test_abi_cpu.py::test_product_never_imports_the_oracle checks).

Each function restates one piece of the reference path and cites it.  The
native arithmetic is in ``nuhtc_oracle.c`` (see its header for the parity
status of every piece); the Python glue below restates the reference's own
pure-torch wrappers, which ARE pinned: ``tests/golden/make_golden.py`` executes
the reference's source for those wrappers in the build container and commits
the vectors this module is checked against.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from typing import Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libnuhtc_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "nuhtc_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        c_fp = ctypes.POINTER(ctypes.c_float)
        c_dp = ctypes.POINTER(ctypes.c_double)
        c_i64p = ctypes.POINTER(ctypes.c_int64)
        c_u8p = ctypes.POINTER(ctypes.c_uint8)
        L.oracle_roi_align_fwd.argtypes = [c_fp] + [ctypes.c_int] * 4 + [c_fp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                                        ctypes.c_float, ctypes.c_int, ctypes.c_int, c_fp, ctypes.c_int]
        L.oracle_nms.argtypes = [c_fp, c_fp, ctypes.c_int64, ctypes.c_float, ctypes.c_int, c_i64p]
        L.oracle_nms.restype = ctypes.c_int64
        L.oracle_mask_iou.argtypes = [c_u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_dp]
        L.oracle_mask_area.argtypes = [c_u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_i64p]
        L.oracle_poly_area2.argtypes = [c_dp, ctypes.c_int]
        L.oracle_poly_area2.restype = ctypes.c_double
        L.oracle_poly_inter_area.argtypes = [c_dp, ctypes.c_int, c_dp, ctypes.c_int]
        L.oracle_poly_inter_area.restype = ctypes.c_double
        L.oracle_poly_iou.argtypes = [c_dp, ctypes.c_int, c_dp, ctypes.c_int]
        L.oracle_poly_iou.restype = ctypes.c_double
        L.oracle_merge.argtypes = [c_dp, c_i64p, c_dp, ctypes.c_int64, ctypes.c_double, ctypes.c_int, c_i64p]
        L.oracle_merge.restype = ctypes.c_int64
        L.oracle_merge_check.argtypes = [c_dp, c_i64p, c_dp, ctypes.c_int64, ctypes.c_double, ctypes.c_int, c_i64p, ctypes.c_int, c_dp, ctypes.c_double, c_i64p, ctypes.c_int64, c_i64p]
        L.oracle_merge_check.restype = ctypes.c_int64
        L.oracle_poly_inter_area_slab.argtypes = [c_dp, ctypes.c_int, c_dp, ctypes.c_int]
        L.oracle_poly_inter_area_slab.restype = ctypes.c_double
        L.oracle_paste.argtypes = [c_fp, c_fp] + [ctypes.c_int] * 5 + [c_fp, ctypes.c_int]
        L.oracle_contour0.argtypes = [c_u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int32), ctypes.c_int]
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _i64p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))


def _u8p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))


def _np32(t) -> np.ndarray:
    if isinstance(t, torch.Tensor):
        t = t.detach().cpu().numpy()
    return np.ascontiguousarray(t, dtype=np.float32)


# --------------------------------------------------------------------------- RoIAlign
def roi_align(input, rois, output_size, spatial_scale=1.0, sampling_ratio=0, pool_mode="avg", aligned=True,
              nthreads: int = 1) -> torch.Tensor:
    """mmcv.ops.roi_align CPU forward (avg).  input [B,C,H,W], rois [K,5] -> [K,C,ph,pw].

    Restates mmcv 1.7.2 ``roi_align_forward`` CPU; called by the reference at
    mmdet/models/roi_heads/roi_extractors/single_level_roi_extractor.py:79,96,103 and
    nuhtc/models/roi_extractors_cus.py:198,218."""
    assert pool_mode == "avg"
    ph, pw = (output_size, output_size) if isinstance(output_size, int) else output_size
    x = _np32(input)
    r = _np32(rois)
    B, C, H, W = x.shape
    K = r.shape[0]
    out = np.empty((K, C, ph, pw), dtype=np.float32)
    if K:
        lib().oracle_roi_align_fwd(_fp(x), B, C, H, W, _fp(r), K, ph, pw, float(spatial_scale), int(sampling_ratio),
                                   int(bool(aligned)), _fp(out), int(nthreads))
    return torch.from_numpy(out)


def map_roi_levels(rois: torch.Tensor, num_levels: int, finest_scale: float = 56) -> torch.Tensor:
    """SingleRoIExtractor.map_roi_levels (single_level_roi_extractor.py:36-55), torch CPU ops."""
    scale = torch.sqrt((rois[:, 3] - rois[:, 1]) * (rois[:, 4] - rois[:, 2]))
    lv = torch.floor(torch.log2(scale / finest_scale + 1e-6))
    return lv.clamp(min=0, max=num_levels - 1).long()


def single_roi_extract(feats: Sequence[torch.Tensor], rois: torch.Tensor, featmap_strides: Sequence[int],
                       output_size: int, sampling_ratio: int, finest_scale: float = 56, nthreads: int = 1):
    """SingleRoIExtractor.forward (single_level_roi_extractor.py:58-115): route each RoI to one level."""
    K = rois.shape[0]
    C = feats[0].shape[1]
    out = torch.zeros(K, C, output_size, output_size)
    if len(feats) == 1:
        if K == 0:
            return out
        return roi_align(feats[0], rois, output_size, 1.0 / featmap_strides[0], sampling_ratio, nthreads=nthreads)
    lv = map_roi_levels(rois, len(feats), finest_scale)
    for i in range(len(feats)):
        inds = (lv == i).nonzero(as_tuple=False).squeeze(1)
        if inds.numel() > 0:
            out[inds] = roi_align(feats[i], rois[inds], output_size, 1.0 / featmap_strides[i], sampling_ratio,
                                  nthreads=nthreads)
    return out


def sum_roi_extract(feats: Sequence[torch.Tensor], rois: torch.Tensor, featmap_strides: Sequence[int],
                    output_size: int, sampling_ratio: int, nthreads: int = 1):
    """The RoIAlign branch of AttentionRoIExtractor.forward (nuhtc/models/roi_extractors_cus.py:213-218,246):
    every RoI is pooled on every given level and the level outputs are summed in level order."""
    K = rois.shape[0]
    C = feats[0].shape[1]
    out = torch.zeros(K, C, output_size, output_size)
    for i in range(len(feats)):
        out = out + roi_align(feats[i], rois, output_size, 1.0 / featmap_strides[i], sampling_ratio, nthreads=nthreads)
    return out


def attention_roi_extract(feats: Sequence[torch.Tensor], rois: torch.Tensor, featmap_strides: Sequence[int], output_size: int,
                          sampling_ratio: int, start_level: int = 2, thres: float = 0.0, nthreads: int = 1) -> torch.Tensor:
    """AttentionRoIExtractor.forward, aggregation='sum' (nuhtc/models/roi_extractors_cus.py:195-259), fp32 CPU branch:
    levels < start_level -> RoIAlign on every RoI; levels >= start_level -> cosine-attention pooled vector of the RoI's
    centre cell, broadcast over the bins; all summed in level order."""
    K, C = rois.shape[0], feats[0].shape[1]
    out = torch.zeros(K, C, output_size, output_size)
    if K == 0:
        return out
    for i, f in enumerate(feats):
        if i < start_level:
            t = roi_align(f, rois, output_size, 1.0 / featmap_strides[i], sampling_ratio, nthreads=nthreads)
        else:
            B, _, H, W = f.shape
            sf = 4 * 2 ** i
            rx = torch.div(rois[:, 1] + rois[:, 3], 2 * sf, rounding_mode="floor").clamp(0, W - 1)
            ry = torch.div(rois[:, 2] + rois[:, 4], 2 * sf, rounding_mode="floor").clamp(0, H - 1)
            loc = torch.stack((rois[:, 0], ry, rx), dim=1)
            uni, inv = loc.unique(dim=0, return_inverse=True)
            uni = uni.long()
            vec = f[uni[:, 0], :, uni[:, 1], uni[:, 2]]
            pos = f[uni[:, 0]].permute(0, 2, 3, 1).reshape(-1, H * W, C)
            sim = F.relu(F.cosine_similarity(vec.unsqueeze(1), pos, dim=2) - thres) + thres
            pooled = torch.mean(f[uni[:, 0]] * sim.view(-1, 1, H, W), dim=(2, 3))
            t = pooled[inv][:, :, None, None].expand(K, C, output_size, output_size)
        out = out + t
    return out


# --------------------------------------------------------------------------- NMS
def nms(boxes, scores, iou_threshold, offset=0, score_threshold=0, max_num=-1):
    """mmcv.ops.nms (Python wrapper + nms_cpu).  Returns (dets [k,5], inds [k] int64)."""
    b = _np32(boxes).reshape(-1, 4)
    s = _np32(scores).reshape(-1)
    assert b.shape[0] == s.shape[0] and offset in (0, 1)
    valid = None
    if score_threshold > 0:
        valid = np.nonzero(s > np.float32(score_threshold))[0]
        b, s = np.ascontiguousarray(b[valid]), np.ascontiguousarray(s[valid])
    keep = np.empty(b.shape[0], dtype=np.int64)
    k = lib().oracle_nms(_fp(b), _fp(s), b.shape[0], float(iou_threshold), int(offset), _i64p(keep))
    keep = keep[:k]
    if max_num > 0:
        keep = keep[:max_num]
    dets = np.concatenate([b[keep], s[keep, None]], axis=1)
    if valid is not None:
        keep = valid[keep]
    return torch.from_numpy(dets), torch.from_numpy(keep.astype(np.int64))


def batched_nms(boxes: torch.Tensor, scores: torch.Tensor, idxs: torch.Tensor, nms_cfg: Optional[dict],
                class_agnostic: bool = False):
    """mmcv.ops.batched_nms 1.7.2 (call sites nuhtc/models/bbox_head.py:93,208).  Tie order: lower index first."""
    if nms_cfg is None:
        order = torch.from_numpy(np.argsort(-scores.numpy().astype(np.float64), kind="stable"))
        return torch.cat([boxes[order], scores[order][:, None]], -1), order
    cfg = dict(nms_cfg)
    class_agnostic = cfg.pop("class_agnostic", class_agnostic)
    if class_agnostic:
        boxes_for_nms = boxes
    else:
        max_coordinate = boxes.max()
        offsets = idxs.to(boxes) * (max_coordinate + torch.tensor(1).to(boxes))
        boxes_for_nms = boxes + offsets[:, None]
    assert cfg.pop("type", "nms") == "nms"
    split_thr = cfg.pop("split_thr", 10000)
    if boxes_for_nms.shape[0] < split_thr:
        dets, keep = nms(boxes_for_nms, scores, **cfg)
        out_boxes = boxes[keep]
        out_scores = dets[:, -1]
    else:
        max_num = cfg.pop("max_num", -1)
        total = torch.zeros(scores.shape, dtype=torch.bool)
        after = torch.zeros_like(scores)
        for cid in torch.unique(idxs):
            m = (idxs == cid).nonzero(as_tuple=False).view(-1)
            dets, keep = nms(boxes_for_nms[m], scores[m], **cfg)
            total[m[keep]] = True
            after[m[keep]] = dets[:, -1]
        keep = total.nonzero(as_tuple=False).view(-1)
        sc = after[keep]
        order = torch.from_numpy(np.argsort(-sc.numpy().astype(np.float64), kind="stable"))
        keep = keep[order]
        out_scores = sc[order]
        out_boxes = boxes[keep]
        if max_num > 0:
            keep, out_boxes, out_scores = keep[:max_num], out_boxes[:max_num], out_scores[:max_num]
    return torch.cat([out_boxes, out_scores[:, None]], -1), keep


def delta2bbox(rois, deltas, means=(0., 0., 0., 0.), stds=(1., 1., 1., 1.), max_shape=None, wh_ratio_clip=16 / 1000):
    """mmdet/core/bbox/coder/delta_xywh_bbox_coder.py:163-260, class-agnostic [N,4] form, torch CPU ops."""
    if deltas.shape[0] == 0:
        return deltas
    d = deltas * deltas.new_tensor(stds).view(1, 4) + deltas.new_tensor(means).view(1, 4)
    ctr = (rois[:, :2] + rois[:, 2:]) * 0.5
    size = rois[:, 2:] - rois[:, :2]
    lim = abs(math.log(wh_ratio_clip))
    shift = size * d[:, :2]
    g_ctr = ctr + shift
    g_size = size * d[:, 2:].clamp(min=-lim, max=lim).exp()
    out = torch.cat([g_ctr - g_size * 0.5, g_ctr + g_size * 0.5], dim=-1)
    if max_shape is not None:
        out[:, 0::2].clamp_(min=0, max=max_shape[1])
        out[:, 1::2].clamp_(min=0, max=max_shape[0])
    return out


def multiclass_nms(multi_bboxes, multi_scores, score_thr, nms_cfg, max_num=-1):
    """nuhtc/models/bbox_head.py:12-102 (class-agnostic boxes [n,4], scores [n,C+1]); returns (dets, labels, keep)
    where keep indexes the flattened (n*C) candidate list AFTER the score filter, as the reference's does."""
    n, C = multi_scores.shape[0], multi_scores.shape[1] - 1
    if multi_bboxes.shape[1] > 4:
        bboxes = multi_bboxes.view(n, -1, 4)
    else:
        bboxes = multi_bboxes[:, None].expand(n, C, 4)
    scores = multi_scores[:, :-1]
    labels = torch.arange(C, dtype=torch.long).view(1, -1).expand_as(scores)
    bboxes, scores, labels = bboxes.reshape(-1, 4), scores.reshape(-1), labels.reshape(-1)
    inds = (scores > score_thr).nonzero(as_tuple=False).squeeze(1)
    bboxes, scores, labels = bboxes[inds], scores[inds], labels[inds]
    if bboxes.numel() == 0:
        return torch.cat([bboxes, scores[:, None]], -1), labels, inds
    dets, keep = batched_nms(bboxes, scores, labels, nms_cfg)
    if max_num > 0:
        dets, keep = dets[:max_num], keep[:max_num]
    return dets, labels[keep], inds[keep]


def rpn_bbox_post_process(scores, rpn_bbox_pred, anchors, ids, img_shape, nms_cfg, max_per_img, min_bbox_size=0):
    """RPNHead._bbox_post_process (mmdet/models/dense_heads/rpn_head.py:189-236) on concatenated level tensors."""
    proposals = delta2bbox(anchors, rpn_bbox_pred, max_shape=img_shape)
    if min_bbox_size >= 0:
        w = proposals[:, 2] - proposals[:, 0]
        h = proposals[:, 3] - proposals[:, 1]
        valid = (w > min_bbox_size) & (h > min_bbox_size)
        if not valid.all():
            proposals, scores, ids = proposals[valid], scores[valid], ids[valid]
    if proposals.numel() == 0:
        return proposals.new_zeros(0, 5)
    dets, _ = batched_nms(proposals, scores, ids, nms_cfg)
    return dets[:max_per_img]


def contour0(mask: np.ndarray, approx_simple: bool = True) -> np.ndarray:
    """cv2.findContours(mask, RETR_TREE, CHAIN_APPROX_SIMPLE)[0][0] as [n,2] int32 (x, y); empty mask -> [0,2]."""
    m = np.ascontiguousarray(mask, dtype=np.uint8)
    h, w = m.shape
    cap = 2 * h * w + 8
    out = np.empty((cap, 2), dtype=np.int32)
    n = lib().oracle_contour0(_u8p(m), h, w, int(approx_simple), out.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), cap)
    return out[:n].copy()


def mask2inst(inst_map: np.ndarray) -> np.ndarray:
    """tools/infer_wsi.py:51-54: the first contour, closed by repeating its first point, shape [n+1,1,2]."""
    c = contour0(inst_map).reshape(-1, 1, 2)
    return np.concatenate([c, c[[0]]], axis=0)


# --------------------------------------------------------------------------- paste
def paste_masks(masks: torch.Tensor, boxes: torch.Tensor, img_h: int, img_w: int) -> torch.Tensor:
    """_do_paste_mask(..., skip_empty=False) (mmdet/.../fcn_mask_head.py:344-412): torch CPU grid_sample on the
    normalised grid the reference builds.  masks [N,1,h,w] probs, boxes [N,4] -> [N,img_h,img_w] fp32."""
    N = masks.shape[0]
    bx0, by0, bx1, by1 = [boxes[:, i:i + 1].to(torch.float32) for i in range(4)]
    ys = torch.arange(0, img_h).to(torch.float32) + 0.5
    xs = torch.arange(0, img_w).to(torch.float32) + 0.5
    ny = (ys - by0) / (by1 - by0) * 2 - 1
    nx = (xs - bx0) / (bx1 - bx0) * 2 - 1
    nx[torch.isinf(nx)] = 0
    ny[torch.isinf(ny)] = 0
    grid = torch.stack([nx[:, None, :].expand(N, img_h, img_w), ny[:, :, None].expand(N, img_h, img_w)], dim=3)
    return F.grid_sample(masks.to(torch.float32), grid, align_corners=False)[:, 0]


def paste_masks_c(masks, boxes, img_h, img_w, nthreads=1) -> torch.Tensor:
    """torch-free scalar twin of paste_masks (generic ATen grid_sampler formula), used for CPU timing."""
    m = _np32(masks).reshape(masks.shape[0], masks.shape[-2], masks.shape[-1])
    b = _np32(boxes)[:, :4].copy()
    out = np.empty((m.shape[0], img_h, img_w), dtype=np.float32)
    lib().oracle_paste(_fp(m), _fp(b), m.shape[0], m.shape[1], m.shape[2], img_h, img_w, _fp(out), int(nthreads))
    return torch.from_numpy(out)


def get_seg_masks(mask_prob: torch.Tensor, det_bboxes: torch.Tensor, ori_h: int, ori_w: int, scale_factor,
                  rescale: bool, thr: float = 0.5) -> torch.Tensor:
    """FCNMaskHead.get_seg_masks core (fcn_mask_head.py:248-306), class-agnostic head: [N,1,h,w] already-sigmoid
    probabilities + boxes -> bool [N,H,W] (the per-label regrouping of :308-309 is left to the caller)."""
    boxes = det_bboxes[:, :4]
    sf = torch.as_tensor(scale_factor, dtype=torch.float32)
    if rescale:
        img_h, img_w = ori_h, ori_w
        boxes = boxes / sf.to(boxes)
    else:
        img_h = int(np.round(ori_h * float(sf[1])).astype(np.int32))
        img_w = int(np.round(ori_w * float(sf[0])).astype(np.int32))
    if mask_prob.shape[0] == 0:
        return torch.zeros(0, img_h, img_w, dtype=torch.bool)
    return paste_masks(mask_prob, boxes, img_h, img_w) >= thr


# --------------------------------------------------------------------------- mask NMS
def mask_iou(masks: np.ndarray) -> np.ndarray:
    """maskUtils.iou(rles, rles, [0]*n) for dense uint8 masks [n,h,w] (pycocotools rleIou)."""
    m = np.ascontiguousarray(masks, dtype=np.uint8)
    n, h, w = m.shape
    out = np.zeros((n, n), dtype=np.float64)
    if n:
        lib().oracle_mask_iou(_u8p(m), n, h, w, _dp(out))
    return out


def mask_area(masks: np.ndarray) -> np.ndarray:
    m = np.ascontiguousarray(masks, dtype=np.uint8)
    n, h, w = m.shape
    out = np.zeros(n, dtype=np.int64)
    if n:
        lib().oracle_mask_area(_u8p(m), n, h, w, _i64p(out))
    return out


def mask_nms(masks: np.ndarray, pred_scores: np.ndarray, thr: float = 0.9) -> np.ndarray:
    """tools/infer_wsi.py:60-84 -- returns the kept indices in the order the reference returns them
    (``sort_idx[keep_idx==1]``).  Sorting is np.argsort(scores)[::-1] exactly as the reference writes it."""
    sort_idx = np.argsort(pred_scores)[::-1]
    n = len(masks)
    iou = mask_iou(np.asarray(masks)[sort_idx])
    keep = np.ones(n, dtype=np.uint8)
    for i in range(n):
        if not keep[i]:
            continue
        row = iou[i]
        for j in range(i + 1, n):
            if keep[j] and row[j] > thr:
                keep[j] = 0
    return sort_idx[keep == 1]


# --------------------------------------------------------------------------- polygon merge
def poly_area(xy: np.ndarray) -> float:
    p = np.ascontiguousarray(xy, dtype=np.float64)
    return abs(lib().oracle_poly_area2(_dp(p), p.shape[0])) * 0.5


def poly_inter_area(P: np.ndarray, Q: np.ndarray) -> float:
    p = np.ascontiguousarray(P, dtype=np.float64)
    q = np.ascontiguousarray(Q, dtype=np.float64)
    return lib().oracle_poly_inter_area(_dp(p), p.shape[0], _dp(q), q.shape[0])


def poly_inter_area_slab(P: np.ndarray, Q: np.ndarray) -> float:
    """Independent second algorithm (vertical slabs + even-odd rule) -- anchors poly_inter_area in the tests."""
    p = np.ascontiguousarray(P, dtype=np.float64)
    q = np.ascontiguousarray(Q, dtype=np.float64)
    return lib().oracle_poly_inter_area_slab(_dp(p), p.shape[0], _dp(q), q.shape[0])


def poly_iou(P: np.ndarray, Q: np.ndarray) -> float:
    p = np.ascontiguousarray(P, dtype=np.float64)
    q = np.ascontiguousarray(Q, dtype=np.float64)
    return lib().oracle_poly_iou(_dp(p), p.shape[0], _dp(q), q.shape[0])


def merge_overlap_arrays(xy: np.ndarray, voff: np.ndarray, score: np.ndarray, overlap_threshold: float = 0.01,
                         merge_strategy: str = "probability") -> np.ndarray:
    """tools/nuclei_merge.py:62-174 on flat arrays: returns the ORIGINAL indices of the kept nuclei, ordered by
    their row in the returned frame (row r gets ``nuclei_id`` r, nuclei_merge.py:201)."""
    xy = np.ascontiguousarray(xy, dtype=np.float64)
    voff = np.ascontiguousarray(voff, dtype=np.int64)
    score = np.ascontiguousarray(score, dtype=np.float64)
    N = score.shape[0]
    out = np.empty(N, dtype=np.int64)
    strat = {"probability": 0, "area": 1}[merge_strategy]
    k = lib().oracle_merge(_dp(xy), _i64p(voff), _dp(score), N, float(overlap_threshold), strat, _i64p(out))
    return out[:k]


def merge_overlap_arrays_check(xy, voff, score, overlap_threshold=0.01, merge_strategy="probability", naive=True,
                               slab=False, with_margin=False):
    """Same greedy loop with an O(N^2) candidate search (naive) and / or the slab-decomposition area (slab): the
    independent cross-checks of merge_overlap_arrays."""
    xy = np.ascontiguousarray(xy, dtype=np.float64)
    voff = np.ascontiguousarray(voff, dtype=np.int64)
    score = np.ascontiguousarray(score, dtype=np.float64)
    N = score.shape[0]
    out = np.empty(N, dtype=np.int64)
    strat = {"probability": 0, "area": 1}[merge_strategy]
    margin = ctypes.c_double(0.0)
    ties = np.zeros((4096, 2), dtype=np.int64)
    nties = ctypes.c_int64(0)
    k = lib().oracle_merge_check(_dp(xy), _i64p(voff), _dp(score), N, float(overlap_threshold), strat, _i64p(out),
                                 int(bool(naive)) | (int(bool(slab)) << 1), ctypes.byref(margin), 1e-9, _i64p(ties),
                                 ties.shape[0], ctypes.byref(nties))
    if with_margin:   # min |IoU - thr| over the decisive pairs + the pairs below 1e-9 (outside the parity contract)
        return out[:k], margin.value, ties[: nties.value].copy()
    return out[:k]


def drop_threshold_ties(d: dict, overlap_threshold: float, max_rounds: int = 8):
    """Integer-vertex polygons make IoU a rational that CAN equal the threshold exactly (e.g. 74/3 / 1480/3 = 1/20), and
    what GEOS's double arithmetic decides there is a coin flip: |IoU - thr| < 1e-9 is outside the parity contract
    (SURVEY.md H3).  Golden cases therefore drop the lower-scored nucleus of every such decisive pair -- and the lower-scored
    copies of identical rings (see below).  Returns (cleaned dict, sorted removed original indices)."""
    xy, voff, score = d["xy"], d["voff"], d["score"]
    alive = np.ones(len(score), dtype=bool)
    # identical rings: the reference keys a dict by the shapely polygon (nuclei_merge.py:101-103), so geometrically
    # identical nuclei collide and it keeps the LOWER-scored copy -- out of contract like score ties; keep the top copy only
    best = {}
    for i in range(len(score)):
        key = xy[voff[i]:voff[i + 1]].tobytes()
        j = best.get(key)
        if j is None:
            best[key] = i
        else:
            lo, hi = (i, j) if score[i] < score[j] else (j, i)
            alive[lo] = False
            best[key] = hi
    for _ in range(max_rounds):
        idx = np.nonzero(alive)[0]
        cnt = np.diff(voff)[idx]
        nv = np.zeros(len(idx) + 1, dtype=np.int64)
        nv[1:] = np.cumsum(cnt)
        sel = np.repeat(voff[:-1][idx], cnt) + (np.arange(nv[-1]) - np.repeat(nv[:-1], cnt))
        sub = dict(xy=xy[sel], voff=nv, score=score[idx])
        ties = set()
        for strat in ("probability", "area"):
            _, _, t = merge_overlap_arrays_check(sub["xy"], sub["voff"], sub["score"], overlap_threshold, strat, naive=False,
                                                 with_margin=True)
            ties |= {(int(a), int(b)) for a, b in t}
        if not ties:
            out = dict(d)
            out.update(sub)
            if "tile_id" in d:
                out["tile_id"] = d["tile_id"][idx]
            return out, np.nonzero(~alive)[0]
        for a, b in ties:
            lo = a if sub["score"][a] < sub["score"][b] else b
            alive[idx[lo]] = False
    raise RuntimeError("threshold ties remain")


# --------------------------------------------------------------------------- watershed proposals
def gaussian_blur5(x: torch.Tensor) -> torch.Tensor:
    """torchvision TF.gaussian_blur(x, kernel_size=5), sigma=None -> 0.3*((5-1)*0.5-1)+0.8 = 1.1, reflect padding
    (torchvision/transforms/_functional_tensor.py: _get_gaussian_kernel2d, gaussian_blur)."""
    sigma = 0.3 * ((5 - 1) * 0.5 - 1) + 0.8
    g = torch.linspace(-2.0, 2.0, steps=5, dtype=x.dtype)
    pdf = torch.exp(-0.5 * (g / sigma).pow(2))
    k1 = pdf / pdf.sum()
    k2 = torch.mm(k1[:, None], k1[None, :]).expand(x.shape[1], 1, 5, 5)
    return F.conv2d(F.pad(x, [2, 2, 2, 2], mode="reflect"), k2, groups=x.shape[1])


def watershed_semantic_mask(semantic_pred: torch.Tensor, img_shape, thres: float = 0.0) -> torch.Tensor:
    """First (device) half of _watershed_proposal, nuhtc/models/htc_roi_head_cus.py:285-299 -> [B,H,W] float 0/1."""
    m = F.interpolate(semantic_pred, size=tuple(int(v) for v in img_shape[:2]), mode="bilinear", align_corners=True)
    m = gaussian_blur5(m)
    m = (m > thres).to(semantic_pred.dtype)
    k = torch.ones((1, 1, 5, 5), dtype=m.dtype)
    for _ in range(2):
        m = torch.clamp(F.conv2d(m, k, padding=2) - k.sum() + 1, min=0, max=1)
    for _ in range(2):
        m = torch.clamp(F.conv2d(m, k, padding=2), min=0, max=1)
    return m[:, 0]


def watershed_instances(mask: np.ndarray, min_area: int = 10):
    """Second (host) half, htc_roi_head_cus.py:303-335, for ONE image: binary_fill_holes, EDT, label(distance > 0.25),
    watershed, area filter, _inst_mask_to_bbox (:263-281) -> (boxes [n,5] float32, filled mask).

    skimage is not in this image.  `watershed(-distance, markers, mask=mask)` floods from the markers over the mask; here the
    markers cover the whole mask (checked below: the Euclidean distance of a foreground pixel is >= 1 > 0.25), so there is
    nothing left to flood and it returns the markers."""
    from scipy import ndimage as ndi
    filled = ndi.binary_fill_holes(mask.astype(bool))
    distance = ndi.distance_transform_edt(filled)
    dist_mask = distance > 0.25
    assert (dist_mask == filled).all()
    markers, n = ndi.label(dist_mask)
    inst = markers
    max_area = mask.shape[0] * mask.shape[1] / 4
    boxes = []
    for lab in range(1, n + 1):
        ys, xs = np.nonzero(inst == lab)
        area = ys.size
        if area > min_area and area < max_area:
            boxes.append([xs.min(), ys.min(), xs.max() + 1, ys.max() + 1, 1.0])
    return np.asarray(boxes, dtype=np.float32).reshape(-1, 5), filled


def watershed_proposal(semantic_pred: torch.Tensor, proposal_list=None, img_shape=None, min_area: int = 10, thres: float = 0.0):
    """_watershed_proposal(semantic_pred, proposal_list=..., img_shape=..., min_area, thres) with semantic_dist / sample_num None."""
    m = watershed_semantic_mask(semantic_pred, img_shape, thres).numpy()
    ws = [torch.from_numpy(watershed_instances(m[i], min_area)[0]) for i in range(m.shape[0])]
    if proposal_list is not None:
        proposal_list = list(proposal_list)
        for i, w in enumerate(ws):
            if len(w):
                proposal_list[i] = torch.cat((w, proposal_list[i]), dim=0)
    return proposal_list, ws
