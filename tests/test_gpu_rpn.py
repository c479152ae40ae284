"""RPN proposal post-processing on the GPU against the oracle restatement of
RPNHead._get_bboxes_single / _bbox_post_process (rpn_head.py:103-236) and the golden vector made from the reference source."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


class Cfg(dict):
    __getattr__ = dict.get


def _levels(B, seed, frame=512, A=3, C=1):
    g = torch.Generator().manual_seed(seed)
    cls, reg, anchors = [], [], []
    for l, s in enumerate((4, 8, 16, 32)):
        h = frame // s
        cls.append(torch.randn(B, A * C, h, h, generator=g) * 2)
        reg.append(torch.randn(B, A * 4, h, h, generator=g) * 0.25)
        ys, xs = torch.meshgrid(torch.arange(h), torch.arange(h), indexing="ij")
        ctr = torch.stack([xs, ys], -1).reshape(-1, 1, 2).float() * s
        wh = torch.tensor([[1.0, 1.0], [1.4, 0.7], [0.7, 1.4]]) * (4.0 * s)
        a = torch.cat([ctr - wh / 2, ctr + wh / 2], -1).reshape(-1, 4)   # (h, w, A) order like AnchorGenerator
        anchors.append(a)
    return cls, reg, anchors


def _oracle_single(oracle, cls, reg, anchors, b, cfg):
    sc, dl, an, ids = [], [], [], []
    for l in range(len(cls)):
        s = cls[l][b].permute(1, 2, 0).reshape(-1).sigmoid()
        d = reg[l][b].permute(1, 2, 0).reshape(-1, 4)
        a = anchors[l]
        if 0 < cfg.nms_pre < s.numel():
            rs, ri = s.sort(descending=True)
            s, d, a = rs[:cfg.nms_pre], d[ri[:cfg.nms_pre]], a[ri[:cfg.nms_pre]]
        sc.append(s); dl.append(d); an.append(a); ids.append(torch.full((s.numel(),), l, dtype=torch.long))
    return sc, dl, an, ids


def test_post_process_golden():
    from nuhtc_b200 import rpn
    z = np.load(os.path.join(GOLD, "rpn_post.npz"))
    t = lambda k: torch.from_numpy(z[k]).cuda()
    cfg = Cfg(min_bbox_size=0, nms=dict(type="nms", iou_threshold=0.7), max_per_img=300)
    dets = rpn.bbox_post_process([t("scores")], [t("deltas")], [t("anchors")], [t("ids")], cfg, (512, 512, 3))
    ref = torch.from_numpy(z["dets"])
    assert dets.shape == ref.shape
    assert torch.equal(dets[:, 4].cpu(), ref[:, 4])                       # same survivors, same order
    assert (dets[:, :4].cpu() - ref[:, :4]).abs().max().item() <= 1e-3    # decode uses the device exp


@pytest.mark.parametrize("nms_pre,min_size", [(1000, 0), (300, 6)])
def test_single_and_batched_vs_oracle(oracle, nms_pre, min_size):
    from nuhtc_b200 import rpn
    B = 3
    cls, reg, anchors = _levels(B, seed=11)
    cfg = Cfg(nms_pre=nms_pre, min_bbox_size=min_size, nms=dict(type="nms", iou_threshold=0.7), max_per_img=400)
    gc, gr, ga = [x.cuda() for x in cls], [x.cuda() for x in reg], [x.cuda() for x in anchors]
    batched = rpn.proposals_batched(gc, gr, ga, (512, 512, 3), cfg)
    for b in range(B):
        # the per-level top-k is taken from the GPU's own sigmoid scores (a 1-ulp device/host sigmoid difference could swap
        # the k-th candidate), the NMS then runs on the decoded boxes of the same candidates on both sides
        single = rpn.get_bboxes_single([x[b] for x in gc], [x[b] for x in gr], ga, (512, 512, 3), cfg)
        sc, dl, an, ids = [], [], [], []
        for l in range(4):
            s, d, a = rpn._level_topk(gc[l][b], gr[l][b], ga[l], nms_pre, True)
            sc.append(s.cpu()); dl.append(d.cpu()); an.append(a.cpu()); ids.append(torch.full((s.numel(),), l, dtype=torch.long))
        from nuhtc_b200.roi_stage import delta2bbox
        boxes = torch.cat([delta2bbox(a.cuda(), d.cuda(), (1., 1., 1., 1.), max_shape=(512, 512, 3)).cpu() for a, d in zip(an, dl)])
        scores, lid = torch.cat(sc), torch.cat(ids)
        ok = ((boxes[:, 2] - boxes[:, 0]) > min_size) & ((boxes[:, 3] - boxes[:, 1]) > min_size)
        ref, _ = oracle.batched_nms(boxes[ok], scores[ok], lid[ok], dict(type="nms", iou_threshold=0.7))
        ref = ref[:400]
        assert ref.shape[0] > 50
        assert torch.equal(single.cpu(), ref)
        assert torch.equal(batched[b].cpu(), ref)
        # and the all-CPU restatement agrees up to the device exp in the decode
        full = oracle.rpn_bbox_post_process(scores, torch.cat(dl), torch.cat(an), lid, (512, 512, 3),
                                            dict(type="nms", iou_threshold=0.7), 400, min_size)
        if full.shape == ref.shape:
            assert (full - ref).abs().max().item() <= 1e-3


@pytest.mark.parametrize("nms_pre,quant,min_size", [(1000, 0.0, 0), (2000, 0.0, -1), (1000, 0.25, 6), (100000, 0.0, 0)])
def test_fused_topk_decode_equals_sort_path(nms_pre, quant, min_size):
    """nuhtc_rpn_topk_decode (shared-memory radix select + bitonic sort + decode, one launch) against torch's sigmoid + stable
    sort + gathers + the delta2bbox kernel: every array the NMS reads must be bit-identical, ties included (quantised logits
    give thousands of equal scores; the k-th score then is a tie that must be cut by ascending anchor index)."""
    from nuhtc_b200 import rpn
    from nuhtc_b200.roi_stage import delta2bbox
    B = 4
    cls, reg, anchors = _levels(B, seed=5)
    if quant:
        cls = [torch.round(c / quant) * quant for c in cls]
    gc, gr, ga = [x.cuda() for x in cls], [x.cuda() for x in reg], [x.cuda() for x in anchors]
    fused = rpn._topk_decode(gc, gr, ga, (512, 512, 3), nms_pre, min_size, True, "auto")
    assert fused is not None
    boxes, scores, labels, groups, per_image, per_level = fused
    boxes, scores, labels, groups = boxes.view(B, per_image, 4), scores.view(B, per_image), labels.view(B, per_image), groups.view(B, per_image)
    for b in range(B):
        off = 0
        for l in range(4):
            s = gc[l][b].permute(1, 2, 0).reshape(-1).sigmoid()
            d = gr[l][b].permute(1, 2, 0).reshape(-1, 4)
            a = ga[l]
            if 0 < nms_pre < s.numel():
                rs, ri = s.sort(descending=True, stable=True)
                s, d, a = rs[:nms_pre], d[ri[:nms_pre]], a[ri[:nms_pre]]
            k = s.numel()
            ref_boxes = delta2bbox(a.contiguous(), d.contiguous(), (1., 1., 1., 1.), max_shape=(512, 512, 3))
            assert torch.equal(scores[b, off:off + k], s), (b, l)
            assert torch.equal(boxes[b, off:off + k], ref_boxes), (b, l)
            assert (labels[b, off:off + k] == l).all()
            ok = torch.ones(k, dtype=torch.bool, device="cuda")
            if min_size >= 0:
                ok = ((ref_boxes[:, 2] - ref_boxes[:, 0]) > min_size) & ((ref_boxes[:, 3] - ref_boxes[:, 1]) > min_size)
            assert torch.equal(groups[b, off:off + k], torch.where(ok, b, -1).to(torch.int32))
            off += k
        assert off == per_image


def test_batched_fused_equals_sort_impl():
    from nuhtc_b200 import rpn
    cls, reg, anchors = _levels(16, seed=3)
    cfg = Cfg(nms_pre=1000, min_bbox_size=0, nms=dict(type="nms", iou_threshold=0.7), max_per_img=1000)
    gc, gr, ga = [x.cuda() for x in cls], [x.cuda() for x in reg], [x.cuda() for x in anchors]
    a = rpn.proposals_batched(gc, gr, ga, (512, 512, 3), cfg, impl="auto")
    b = rpn.proposals_batched(gc, gr, ga, (512, 512, 3), cfg, impl="sort")
    assert len(a) == len(b) == 16
    for x, y in zip(a, b):
        assert x.shape[0] > 100 and torch.equal(x, y)
