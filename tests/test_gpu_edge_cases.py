"""Empty, degenerate and out-of-frame inputs through every op of the path (what the reference's own tests call the corner
cases): zero RoIs / boxes / masks / nuclei, candidates that are all below the threshold, boxes outside the frame, zero-area
boxes, tiles without detections.  Each result is compared with the oracle (or the definition) on the same input."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_roi_align_degenerate_rois(oracle):
    import nuhtc_b200 as nb
    from nuhtc_b200 import synth
    feats = synth.fpn_levels(2, 64, frame=256, seed=3)
    rois = torch.tensor([[0, 10.0, 10.0, 10.0, 10.0],          # zero area
                         [1, -300.0, -300.0, -200.0, -250.0],   # fully outside (negative)
                         [0, 900.0, 900.0, 1000.0, 990.0],      # fully outside (beyond the frame)
                         [1, -20.0, -20.0, 300.0, 300.0],       # covers more than the frame
                         [0, 100.5, 30.25, 100.5, 90.0],        # zero width
                         [1, 255.0, 255.0, 256.0, 256.0]])      # last pixel
    for P in (7, 14):
        for sr in (0, 2):
            out = nb.roi_align_levels([f.cuda() for f in feats], rois.cuda(), P, [1 / s for s in synth.FPN_STRIDES], sr, mode="route")
            ref = oracle.single_roi_extract(feats, rois, synth.FPN_STRIDES, P, sr)
            assert torch.isfinite(out).all()
            assert (out.cpu() - ref).abs().max().item() <= 1e-5
    empty = nb.roi_align_levels([f.cuda() for f in feats], rois[:0].cuda(), 7, [1 / s for s in synth.FPN_STRIDES], 0, mode="route")
    assert empty.shape == (0, 64, 7, 7)


def test_nms_empty_ignored_and_duplicates(oracle):
    import nuhtc_b200 as nb
    dev = "cuda"
    e = torch.zeros((0, 4), device=dev)
    dets, keep = nb.nms(e, torch.zeros(0, device=dev), 0.5)
    assert dets.shape == (0, 5) and keep.numel() == 0
    dets, keep = nb.batched_nms(e, torch.zeros(0, device=dev), torch.zeros(0, dtype=torch.long, device=dev), dict(type="nms", iou_threshold=0.5))
    assert dets.shape[0] == 0 and keep.numel() == 0
    # every candidate parked in the negative group
    b = torch.rand(50, 4, device=dev) * 10
    b[:, 2:] += b[:, :2]
    k, s, c, st = nb.nms_groups(b, torch.rand(50, device=dev), torch.zeros(50, dtype=torch.long, device=dev),
                                torch.full((50,), -1, dtype=torch.int32, device=dev), 3, 50, 0.5, 0, "offset", num_classes=2)
    assert int(st.item()) == 0 and c.tolist() == [0, 0, 0]
    # one box, and identical boxes with distinct scores: only the best survives
    one = torch.tensor([[1.0, 2.0, 5.0, 9.0]], device=dev)
    _, keep = nb.nms(one, torch.tensor([0.3], device=dev), 0.5)
    assert keep.tolist() == [0]
    same = one.repeat(70, 1)
    sc = torch.linspace(0.1, 0.9, 70, device=dev)
    _, keep = nb.nms(same, sc, 0.5)
    _, kref = oracle.nms(same.cpu(), sc.cpu(), 0.5)
    assert keep.cpu().tolist() == kref.tolist() == [69]
    # zero-area boxes never suppress each other (0/0 is not > thr)
    z = torch.tensor([[3.0, 3.0, 3.0, 3.0]], device=dev).repeat(5, 1)
    zs = torch.tensor([0.5, 0.4, 0.3, 0.2, 0.1], device=dev)
    _, keep = nb.nms(z, zs, 0.5)
    _, kref = oracle.nms(z.cpu(), zs.cpu(), 0.5)
    assert keep.cpu().tolist() == kref.tolist()


def test_paste_pack_masknms_contours_empty_and_outside(oracle):
    import nuhtc_b200 as nb
    dev = "cuda"
    out = nb.paste_masks(torch.zeros((0, 28, 28), device=dev), torch.zeros((0, 4), device=dev), 64, 64, thr=0.5, kind="bin")
    assert out.shape == (0, 64, 64)
    probs = torch.full((4, 28, 28), 0.9)
    boxes = torch.tensor([[-100.0, -100.0, -50.0, -60.0],     # outside: pastes nothing
                          [10.0, 10.0, 10.0, 30.0],           # zero width
                          [5.0, 6.0, 25.0, 30.0],             # regular
                          [50.0, 50.0, 200.0, 200.0]])        # clipped by the frame
    for kind in ("bin", "bits"):
        res = nb.paste_masks(probs.cuda(), boxes.cuda(), 64, 64, thr=0.5, kind=kind, want_stats=True)
        ref = oracle.paste_masks(probs[:, None], boxes, 64, 64) >= 0.5
        area = res[1].cpu()
        assert area.tolist() == ref.sum((1, 2)).tolist()
        if kind == "bin":
            assert torch.equal(res[0].cpu().bool(), ref)
            dense = res[0]
    bits, area, bbox = nb.pack_masks(dense)
    assert bbox[0].tolist() == [0, 0, 0, 0] and int(area[0]) == 0
    scores = torch.tensor([0.9, 0.8, 0.7, 0.6], device=dev)
    keep, ts, tc, st = nb.mask_nms_device(bits, area, bbox, scores, 64, 0.05)
    ref = oracle.mask_nms(dense.cpu().numpy().astype(np.uint8), scores.cpu().numpy(), thr=0.05)
    assert keep[: int(tc[0])].cpu().tolist() == ref.tolist()
    xy, cnt, _ = nb.mask_contours(bits, 64, max_pts=64, bbox=bbox)
    d = dense.cpu().numpy().astype(np.uint8)
    for i in range(4):
        assert np.array_equal(xy[i, : int(cnt[i])].cpu().numpy(), oracle.contour0(d[i]))
    assert int(cnt[0]) == 0
    none = torch.zeros(4, dtype=torch.uint8, device=dev)
    _, cnt2, _ = nb.mask_contours(bits, 64, max_pts=64, bbox=bbox, select=none)
    assert cnt2.tolist() == [0, 0, 0, 0]
    # nothing at all
    eb = torch.zeros((0, 64, 1), dtype=torch.int64, device=dev)
    xy0, c0, _ = nb.mask_contours(eb, 64)
    assert xy0.shape[0] == 0 and c0.numel() == 0
    k0, _, tc0, _ = nb.mask_nms_device(eb, torch.zeros(0, dtype=torch.int32, device=dev), torch.zeros((0, 4), dtype=torch.int32, device=dev),
                                       torch.zeros(0, device=dev), 64, 0.05)
    assert int(tc0[0]) == 0


def test_merge_empty_single_and_identical(oracle):
    import nuhtc_b200 as nb
    dev = "cuda"
    sq = np.array([[0, 0], [0, 4], [4, 4], [4, 0], [0, 0]], dtype=np.float64)
    for polys, scores in (([], []), ([sq], [0.5]), ([sq, sq.copy(), sq + 100.0], [0.2, 0.9, 0.1])):
        xy = np.concatenate(polys) if polys else np.zeros((0, 2))
        voff = np.cumsum([0] + [len(p) for p in polys]).astype(np.int64)
        sc = np.array(scores, dtype=np.float64)
        got = nb.merge_arrays(torch.from_numpy(xy).to(dev), torch.from_numpy(voff).to(dev), torch.from_numpy(sc).to(dev), 0.05)
        ref = oracle.merge_overlap_arrays(xy, voff, sc, 0.05)
        assert got.cpu().tolist() == list(ref)


def test_glue_and_stage_with_empty_tiles(oracle):
    from nuhtc_b200 import det_ops, synth
    from nuhtc_b200.roi_stage import RoIStage, RoIStageConfig
    dev = "cuda"
    assert det_ops.delta2bbox(torch.zeros((0, 5), device=dev), torch.zeros((0, 4), device=dev), max_shape=(64, 64)).shape == (0, 5)
    # tile 1 has proposals but none of its classes passes the score threshold; tile 2 has no proposals at all
    B, n_per = 3, 60
    cfg = RoIStageConfig(extractor="single", max_per_img=20, score_thr=0.05, contour_max_pts=128)
    feats = synth.fpn_levels(B, 64, frame=512, seed=1)
    rois = synth.proposals(2, n_per, "nuclei", frame=512, seed=2)          # tiles 0 and 1 only
    heads = synth.SyntheticHeads(2 * n_per, seed=3)
    for c in heads.cls:
        c[n_per:, :5] = -30.0                                                # tile 1: background wins everywhere
        c[n_per:, 5] = 30.0
    gh = copy.copy(heads).to(dev)
    st = RoIStage(cfg, gh.bbox_heads(), gh.mask_head)
    raw = st.run([f.cuda() for f in feats], rois.cuda(), max_rois_per_tile=n_per)
    raw.check()
    kept = raw.kept_indices()
    assert len(kept) == B and kept[0].numel() > 0 and kept[1].numel() == 0 and kept[2].numel() == 0
    assert int(raw.det_valid.view(B, -1)[1:].sum()) == 0
    res = raw.compact()
    assert (res.det_tile == 0).all()
    from oracle.stage import roi_stage_cpu
    ref = roi_stage_cpu(feats[:], rois, heads.bbox_heads(), heads.mask_head, cfg)
    assert ref[1]["det_boxes"].shape[0] == 0 and ref[2]["det_boxes"].shape[0] == 0
    assert res.det_boxes.shape[0] == ref[0]["det_boxes"].shape[0]
    assert set((kept[0].cpu().numpy()).tolist()) == set(ref[0]["keep"].tolist())
