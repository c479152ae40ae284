"""GPU parity: NMS keep indices must be bit-exact (values and order) vs the CPU oracle (mmcv nms_cpu restatement)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N", [1, 63, 64, 65, 2000, 5000])
def test_nms_keep_exact(oracle, N):
    import nuhtc_b200 as nb
    from nuhtc_b200 import synth
    boxes, scores, _ = synth.nms_boxes(N, seed=N)
    dets_ref, keep_ref = oracle.nms(boxes, scores, 0.5)
    dets, keep = nb.nms(boxes.cuda(), scores.cuda(), 0.5)
    assert torch.equal(keep.cpu(), keep_ref)
    assert torch.equal(dets.cpu(), dets_ref)


def test_nms_options_and_numpy(oracle):
    import nuhtc_b200 as nb
    from nuhtc_b200 import synth
    boxes, scores, _ = synth.nms_boxes(3000, seed=9)
    for kw in (dict(offset=1), dict(score_threshold=0.3), dict(max_num=50), dict(score_threshold=0.2, max_num=77)):
        d_ref, k_ref = oracle.nms(boxes, scores, 0.4, **kw)
        d, k = nb.nms(boxes.cuda(), scores.cuda(), 0.4, **kw)
        assert torch.equal(k.cpu(), k_ref), kw
        assert torch.equal(d.cpu(), d_ref), kw
    d, k = nb.nms(boxes.numpy(), scores.numpy(), 0.5)
    assert isinstance(k, type(boxes.numpy())) and (k == oracle.nms(boxes, scores, 0.5)[1].numpy()).all()
    d, k = nb.nms(torch.zeros(0, 4).cuda(), torch.zeros(0).cuda(), 0.5)
    assert d.shape == (0, 5) and k.numel() == 0


@pytest.mark.parametrize("N", [2000, 5000, 9999, 10000, 20000])
def test_batched_nms_sweep_exact(oracle, N):
    """BASELINE cfg 3 (small half): 5 classes, IoU 0.5; both sides of split_thr=10000."""
    import nuhtc_b200 as nb
    from nuhtc_b200 import synth
    boxes, scores, labels = synth.nms_boxes(N, seed=N + 1)
    cfg = dict(type="nms", iou_threshold=0.5)
    d_ref, k_ref = oracle.batched_nms(boxes, scores, labels, cfg)
    d, k = nb.batched_nms(boxes.cuda(), scores.cuda(), labels.cuda(), cfg)
    assert 0.2 < len(k_ref) / N < 0.8
    assert torch.equal(k.cpu(), k_ref)
    assert torch.equal(d.cpu(), d_ref)
    assert cfg == dict(type="nms", iou_threshold=0.5)  # cfg must not be mutated


def test_batched_nms_variants(oracle):
    import nuhtc_b200 as nb
    from nuhtc_b200 import synth
    boxes, scores, labels = synth.nms_boxes(4000, seed=77)
    for cfg, agn in ((dict(type="nms", iou_threshold=0.5, class_agnostic=True), False),
                     (dict(type="nms", iou_threshold=0.3, max_num=100), False),
                     (dict(type="nms", iou_threshold=0.5, split_thr=1000), False),
                     (dict(type="nms", iou_threshold=0.5, split_thr=1000, max_num=64), False),
                     (dict(type="nms", iou_threshold=0.5, split_thr=1000), True),
                     (dict(iou_threshold=0.6), True)):
        d_ref, k_ref = oracle.batched_nms(boxes, scores, labels, cfg, class_agnostic=agn)
        d, k = nb.batched_nms(boxes.cuda(), scores.cuda(), labels.cuda(), cfg, class_agnostic=agn)
        assert torch.equal(k.cpu(), k_ref), cfg
        assert torch.equal(d.cpu(), d_ref), cfg
    d, k = nb.batched_nms(boxes.cuda(), scores.cuda(), labels.cuda(), None)
    assert torch.equal(k.cpu(), oracle.batched_nms(boxes, scores, labels, None)[1])


def test_offset_arithmetic_is_reproduced(oracle):
    """Large coordinates: the fp32 class offset perturbs the IoUs; keep must still be identical."""
    import nuhtc_b200 as nb
    from nuhtc_b200 import synth
    boxes, scores, labels = synth.nms_boxes(8000, seed=5, density=40.0)
    boxes = boxes + 3000.0
    cfg = dict(type="nms", iou_threshold=0.5)
    _, k_ref = oracle.batched_nms(boxes, scores, labels, cfg)
    _, k = nb.batched_nms(boxes.cuda(), scores.cuda(), labels.cuda(), cfg)
    assert torch.equal(k.cpu(), k_ref)


def test_grouped_nms_equals_per_image_calls(oracle):
    """16 tiles in one launch == 16 reference calls (per-image multiclass_nms loop, htc_roi_head_cus.py:2292)."""
    import nuhtc_b200 as nb
    from nuhtc_b200 import synth
    G, n = 16, 700
    bs, ss, ls, gs = [], [], [], []
    for g in range(G):
        b, s, l = synth.nms_boxes(n + 13 * g, seed=100 + g)
        bs.append(b); ss.append(s); ls.append(l); gs.append(torch.full((b.shape[0],), g, dtype=torch.int32))
    B, S, Lb, Gp = torch.cat(bs), torch.cat(ss), torch.cat(ls), torch.cat(gs)
    perm = torch.randperm(B.shape[0], generator=torch.Generator().manual_seed(1))  # groups need not be contiguous
    B, S, Lb, Gp = B[perm], S[perm], Lb[perm], Gp[perm]
    keep, gstart, gcount, status = nb.nms_groups(B.cuda(), S.cuda(), Lb.cuda(), Gp.cuda(), G, n + 13 * G, 0.5, 0, "offset")
    assert int(status.item()) == 0
    keep, gstart, gcount = keep.cpu(), gstart.cpu(), gcount.cpu()
    for g in range(G):
        idx = (Gp == g).nonzero().squeeze(1)
        _, k_ref = oracle.batched_nms(B[idx], S[idx], Lb[idx], dict(type="nms", iou_threshold=0.5))
        got = keep[gstart[g]: gstart[g] + gcount[g]]
        assert torch.equal(got, idx[k_ref]), g


@pytest.mark.parametrize("N", [50000, 100000, 200000])
def test_batched_nms_large_sweep_exact(oracle, N):
    """BASELINE cfg 3 (large half): keep lists identical (values and order) to the oracle at 50k / 100k / 200k boxes
    (mmcv's per-class loop above split_thr, /root/reference/nuhtc/models/bbox_head.py:93); the oracle needs ~1-16 s."""
    import nuhtc_b200 as nb
    from nuhtc_b200 import synth
    boxes, scores, labels = synth.nms_boxes(N, seed=3)
    cfg = dict(type="nms", iou_threshold=0.5)
    d_ref, k_ref = oracle.batched_nms(boxes, scores, labels, cfg)
    d, k = nb.batched_nms(boxes.cuda(), scores.cuda(), labels.cuda(), cfg)
    assert 0.2 < len(k_ref) / N < 0.8
    assert torch.equal(k.cpu(), k_ref)
    assert torch.equal(d.cpu(), d_ref)
    if N > 50000:
        return
    # score_thr 0.05 of the sweep (multiclass_nms filters before batched_nms, bbox_head.py:60-74)
    m = scores > 0.05
    idx = m.nonzero().squeeze(1)
    _, k_ref2 = oracle.batched_nms(boxes[idx], scores[idx], labels[idx], cfg)
    _, k2 = nb.batched_nms(boxes[idx].cuda(), scores[idx].cuda(), labels[idx].cuda(), cfg)
    assert torch.equal(k2.cpu(), k_ref2)


def test_batched_nms_negative_coordinates_fall_back_to_all_pairs(oracle):
    """Below split_thr mmcv runs one nms over the offset boxes; with a negative coordinate boxes of different classes CAN
    overlap after the offset, so the class-segmented fast path must hand over to the all-pairs form (device status 3)."""
    import nuhtc_b200 as nb
    from nuhtc_b200 import synth
    boxes, scores, labels = synth.nms_boxes(6000, seed=21, density=30.0)
    boxes = boxes - boxes.max() * 0.75          # most coordinates negative: offsets no longer separate the classes
    cfg = dict(type="nms", iou_threshold=0.5)
    _, k_ref = oracle.batched_nms(boxes, scores, labels, cfg)
    _, k = nb.batched_nms(boxes.cuda(), scores.cuda(), labels.cuda(), cfg)
    assert torch.equal(k.cpu(), k_ref)
    # and the segmented path itself (non-negative coordinates) against the same oracle
    boxes2 = boxes - boxes.min() + 1.0
    _, k_ref2 = oracle.batched_nms(boxes2, scores, labels, cfg)
    _, k2 = nb.batched_nms(boxes2.cuda(), scores.cuda(), labels.cuda(), cfg)
    assert torch.equal(k2.cpu(), k_ref2)


@pytest.mark.parametrize("N", [50000, 200000])
def test_large_sweep_properties(N):
    """BASELINE cfg 3 (large half) through size-independent properties: the kept set is conflict-free, every dropped
    box is suppressed by a kept higher-scoring same-class box, order is score-descending, and NMS is idempotent."""
    import nuhtc_b200 as nb
    from nuhtc_b200 import synth
    import torchvision
    boxes, scores, labels = synth.nms_boxes(N, seed=3)
    cfg = dict(type="nms", iou_threshold=0.5)
    b, s, l = boxes.cuda(), scores.cuda(), labels.cuda()
    dets, keep = nb.batched_nms(b, s, l, cfg)
    assert (dets[1:, 4] <= dets[:-1, 4]).all()
    kb, kl = b[keep], l[keep]
    # idempotence
    dets2, keep2 = nb.batched_nms(kb, s[keep], kl, cfg)
    assert keep2.numel() == keep.numel() and torch.equal(keep2, torch.arange(keep.numel(), device="cuda"))
    # every dropped box overlaps (>0.5) a kept same-class box with a higher score: check on a sample
    dropped = torch.ones(N, dtype=torch.bool, device="cuda"); dropped[keep] = False
    didx = dropped.nonzero().squeeze(1)[:: max(1, N // 4000)]
    off = (l.float() * (b.max() + 1))[:, None]
    iou = torchvision.ops.box_iou((b + off)[didx], (b + off)[keep])
    higher = s[keep][None, :] > s[didx][:, None]
    assert ((iou > 0.5) & higher).any(dim=1).all()


def test_class_segmented_nms_equals_all_pairs(oracle):
    """(group, class) sort segments give the same keep lists as the all-pairs test on offset boxes when no coordinate
    is negative; a negative coordinate is reported through status 3."""
    import nuhtc_b200 as nb
    from nuhtc_b200 import synth
    G = 6
    bs, ss, ls, gs = [], [], [], []
    for g in range(G):
        b, s, l = synth.nms_boxes(900 + 31 * g, seed=200 + g)
        bs.append(b); ss.append(s); ls.append(l); gs.append(torch.full((b.shape[0],), g, dtype=torch.int32))
    B, S, Lb, Gp = torch.cat(bs), torch.cat(ss), torch.cat(ls), torch.cat(gs)
    B = B + 20.0   # the generator centres boxes on [0, side]: shift so that no coordinate is negative
    Gp[::17] = -1  # some candidates are not candidates at all
    perm = torch.randperm(B.shape[0], generator=torch.Generator().manual_seed(2))
    B, S, Lb, Gp = B[perm], S[perm], Lb[perm], Gp[perm]
    args = (B.cuda(), S.cuda(), Lb.cuda(), Gp.cuda(), G, 1200, 0.5, 0, "offset")
    k0, s0, c0, st0 = nb.nms_groups(*args)
    k1, s1, c1, st1 = nb.nms_groups(*args, num_classes=5)
    assert int(st0.item()) == 0 and int(st1.item()) == 0
    assert torch.equal(c0, c1)
    for g in range(G):
        a = k0[s0[g]: s0[g] + c0[g]]
        b = k1[s1[g]: s1[g] + c1[g]]
        assert torch.equal(a, b), g
        idx = (Gp == g).nonzero().squeeze(1)
        _, k_ref = oracle.batched_nms(B[idx], S[idx], Lb[idx], dict(type="nms", iou_threshold=0.5))
        assert torch.equal(b.cpu(), idx[k_ref]), g
    # the capacity of the segmented form is per (group, class) list: the tight bound gives the same lists, one less is flagged
    ok = Gp >= 0
    seg_max = int(torch.bincount((Gp[ok].long() * 5 + Lb[ok]), minlength=G * 5).max())
    k2, s2, c2, st2 = nb.nms_groups(B.cuda(), S.cuda(), Lb.cuda(), Gp.cuda(), G, seg_max, 0.5, 0, "offset", num_classes=5)
    assert int(st2.item()) == 0 and torch.equal(c2, c1)
    for g in range(G):
        assert torch.equal(k2[s2[g]: s2[g] + c2[g]], k1[s1[g]: s1[g] + c1[g]])
    _, _, _, st3 = nb.nms_groups(B.cuda(), S.cuda(), Lb.cuda(), Gp.cuda(), G, seg_max - 1, 0.5, 0, "offset", num_classes=5)
    assert int(st3.item()) == 1
    Bn = B.clone()
    Bn[int((Gp >= 0).nonzero()[3])] -= 4000.0
    _, _, _, st = nb.nms_groups(Bn.cuda(), S.cuda(), Lb.cuda(), Gp.cuda(), G, 1200, 0.5, 0, "offset", num_classes=5)
    assert int(st.item()) == 3
