"""GPU parity of the whole RoI stage (BASELINE cfg 1 stand-in: RoI-head-only harness with random FPN features,
random proposals and seeded head outputs, since mmcv / the Swin backbone cannot be imported here).

The device-resident driver is checked at EVERY op boundary against the oracle fed with the same inputs the GPU op saw
(so a 1-ulp difference of torch's CUDA `exp` in the box decode cannot make the comparison flaky), and end to end
against the reference-style per-image CPU flow."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(extractor, C, B=2, n_per=300, sr=0, max_per_img=100, score_thr=0.05):
    from nuhtc_b200 import synth
    from nuhtc_b200.roi_stage import RoIStageConfig
    cfg = RoIStageConfig(extractor=extractor, bbox_sampling_ratio=sr, max_per_img=max_per_img, score_thr=score_thr,
                         contour_max_pts=256)
    feats = synth.fpn_levels(B, C, frame=512, seed=1)
    rois = synth.proposals(B, n_per, "nuclei" if extractor == "sum" else "routed", frame=512, seed=2)
    heads = synth.SyntheticHeads(B * n_per, seed=3)
    return cfg, feats, rois, heads


@pytest.mark.parametrize("extractor,C,sr", [("single", 256, 0), ("sum", 64, 2)])
def test_stage_op_boundaries(oracle, extractor, C, sr):
    from nuhtc_b200.roi_stage import RoIStage
    cfg, feats, rois, heads = _setup(extractor, C, sr=sr)
    gh = copy.copy(heads).to("cuda")
    st = RoIStage(cfg, gh.bbox_heads(), gh.mask_head)
    st.trace = {}
    raw = st.run([f.cuda() for f in feats], rois.cuda())   # slot form: max_per_img slots per tile, no host sync inside
    res = raw.compact()
    tr = st.trace
    ext = (lambda r, P, s: oracle.single_roi_extract(feats, r, cfg.featmap_strides, P, s)) if extractor == "single" else \
          (lambda r, P, s: oracle.sum_roi_extract(feats[:2], r, cfg.featmap_strides[:2], P, s))
    tol = 1e-5 if extractor == "single" else 2e-5
    # RoIAlign, three cascade stages + mask branch, on the RoIs the GPU stage actually used
    for r, f in zip(tr["bbox_rois"], tr["bbox_feats"]):
        assert (f.cpu() - ext(r.cpu(), 7, sr)).abs().max().item() <= tol
    assert (tr["mask_feats"][0].cpu() - ext(tr["mask_rois"][0].cpu(), 14, 0)).abs().max().item() <= tol
    # decode glue vs the reference formula on CPU
    dec = oracle.delta2bbox(tr["bbox_rois"][0][:, 1:].cpu(), heads.reg[0], stds=cfg.stage_stds[0], max_shape=cfg.img_shape)
    assert (tr["bbox_rois"][1][:, 1:].cpu() - dec).abs().max().item() <= 1e-3
    # multiclass NMS per tile: bit-exact given the same decoded boxes / scores
    boxes, scores = tr["nms_boxes"][0].cpu(), tr["nms_scores"][0].cpu()
    keep, gs, gc = tr["nms_keep"][0].cpu(), tr["nms_start"][0].cpu(), tr["nms_count"][0].cpu()
    tile = rois[:, 0].long()
    C5 = cfg.num_classes
    det_ref = []
    for b in range(feats[0].shape[0]):
        idx = (tile == b).nonzero().squeeze(1)
        sc = torch.cat([scores[idx, :C5], torch.zeros(idx.numel(), 1)], 1)
        dets, labels, cand = oracle.multiclass_nms(boxes[idx], sc, cfg.score_thr, dict(type="nms", iou_threshold=0.5), -1)
        glob = idx[cand // C5] * C5 + cand % C5
        assert torch.equal(keep[gs[b]: gs[b] + gc[b]], glob)
        det_ref.append((dets[: cfg.max_per_img], labels[: cfg.max_per_img]))
    ref_boxes = torch.cat([d[0][:, :4] for d in det_ref])
    assert torch.equal(res.det_boxes.cpu(), ref_boxes)
    assert torch.equal(res.det_labels.cpu(), torch.cat([d[1] for d in det_ref]))
    # paste + threshold
    probs, pb = tr["paste_probs"][0].cpu(), tr["paste_boxes"][0].cpu()
    refp = oracle.paste_masks(probs, pb, 256, 256)
    diff = raw.masks.cpu() != (refp >= 0.5)
    assert ((refp - 0.5).abs()[diff] <= 1e-6).all()
    assert not raw.masks[~raw.det_valid].any()                      # padding slots paste nothing
    # mask NMS per tile on the GPU's own masks (slot indices)
    m = raw.masks.cpu().numpy().astype(np.uint8)
    tid = tr["mnms_tile"][0].cpu().numpy()
    kept = raw.kept_indices()
    for b in range(feats[0].shape[0]):
        sel = np.nonzero(tid == b)[0]
        ref = sel[oracle.mask_nms(m[sel], raw.det_scores.cpu().numpy()[sel], thr=0.05)] if len(sel) else sel
        assert (kept[b].cpu().numpy() == ref).all()
        assert len(ref) > 0
    # mask2inst contours of every slot, on the GPU's own masks (tools/infer_wsi.py:528)
    raw.check()
    cxy, ccnt = raw.contour_xy.cpu().numpy(), raw.contour_count.cpu().numpy()
    kept_all = set(int(i) for k in kept for i in k.cpu().tolist())
    for i in range(m.shape[0]):
        if i in kept_all:
            assert np.array_equal(cxy[i, :ccnt[i]], oracle.contour0(m[i]))
        else:
            assert ccnt[i] == 0                                     # suppressed / filtered / padding slots are not traced


def test_stage_end_to_end_vs_reference_flow(oracle):
    """Whole stage vs the per-image CPU flow (oracle/stage.py).  Detections are matched by their candidate identity;
    boxes agree to float tolerance and the mask-NMS survivors are the same set."""
    from nuhtc_b200.roi_stage import RoIStage
    from oracle.stage import roi_stage_cpu
    cfg, feats, rois, heads = _setup("single", 64, B=2, n_per=250, max_per_img=80)
    gh = copy.copy(heads).to("cuda")
    st = RoIStage(cfg, gh.bbox_heads(), gh.mask_head)
    res = st.run([f.cuda() for f in feats], rois.cuda()).compact()
    ref = roi_stage_cpu(feats, rois, heads.bbox_heads(), heads.mask_head, cfg)
    kept = res.kept_indices()
    tile = res.det_tile.cpu().numpy()
    for b, r in enumerate(ref):
        sel = np.nonzero(tile == b)[0]
        assert len(sel) == r["det_boxes"].shape[0]
        assert (res.det_boxes.cpu()[sel] - r["det_boxes"]).abs().max().item() <= 1e-3
        assert torch.equal(res.det_labels.cpu()[sel], r["det_labels"])
        # masks: the two flows paste through boxes that differ by up to 1e-3 px (GPU vs CPU head GEMMs, three decodes deep), so
        # a pixel may only disagree where the reference's pasted probability is within that displacement's reach of the
        # threshold: 1e-3 px of a ~40 px box = 7e-4 of a 28-px map cell at <= 0.25 probability per cell -> 2e-4; bound 1e-3
        diff = res.masks.cpu()[sel] != r["masks"]
        if diff.any():
            H, W = r["masks"].shape[1:]
            prob = oracle.paste_masks(r["mask_prob"], r["det_boxes"], H, W)
            assert ((prob - 0.5).abs()[diff] <= 1e-3).all()
        assert diff.float().mean().item() < 1e-4
        got = set((kept[b].cpu().numpy() - sel[0]).tolist())
        assert got == set(r["keep"].tolist())


def test_bits_lane_equals_dense_lane():
    from nuhtc_b200.roi_stage import RoIStage
    cfg, feats, rois, heads = _setup("single", 64, B=2, n_per=200)
    heads.to("cuda")
    f = [x.cuda() for x in feats]
    a = RoIStage(cfg, heads.bbox_heads(), heads.mask_head).run(f, rois.cuda())
    a.check()
    cfg2 = copy.copy(cfg)
    cfg2.dense_masks = False
    b = RoIStage(cfg2, heads.bbox_heads(), heads.mask_head).run(f, rois.cuda())
    assert b.masks is None and torch.equal(a.mask_bits, b.mask_bits) and torch.equal(a.mask_area, b.mask_area)
    assert torch.equal(a.tile_count, b.tile_count)
    for x, y in zip(a.kept_indices(), b.kept_indices()):
        assert torch.equal(x, y)


def test_stage_contours_feed_the_merge(oracle):
    """masks -> mask NMS survivors -> mask2inst contours (+ tile origin) -> cross-tile merge, all on the device, against the
    oracle driven the way tools/infer_wsi.py:526-546 + tools/nuclei_merge.py do it (on the GPU stage's own masks)."""
    import nuhtc_b200 as nb
    from nuhtc_b200 import det_ops
    from nuhtc_b200.roi_stage import RoIStage
    cfg, feats, rois, heads = _setup("single", 64, B=4, n_per=250, max_per_img=90)
    gh = copy.copy(heads).to("cuda")
    raw = RoIStage(cfg, gh.bbox_heads(), gh.mask_head).run([f.cuda() for f in feats], rois.cuda())
    raw.check()
    D = raw.det_boxes.shape[0]
    tile_xy = torch.tensor([[0, 0], [192, 0], [0, 192], [192, 192]], dtype=torch.int32, device="cuda")   # 64 px overlap
    origin = tile_xy[raw.det_tile.clamp(min=0).long()].contiguous()
    sel = det_ops.keep_flags(raw.keep, raw.tile_start, raw.tile_count, cfg.max_per_img, D).bool()
    ring, voff, index = nb.rings_for_merge(raw.contour_xy, raw.contour_count, select=sel, origin=origin)
    assert index.numel() > 40
    scores = raw.det_scores[index].double()
    kept = nb.merge_arrays(ring, voff, scores, 0.05).cpu().numpy()
    # reference flow on the CPU
    m = raw.masks.cpu().numpy().astype(np.uint8)
    org = origin.cpu().numpy()
    polys, sc = [], []
    for i in index.cpu().tolist():
        c = oracle.mask2inst(m[i]).reshape(-1, 2) + org[i]
        assert len(c) >= 3
        polys.append(c.astype(np.float64))
        sc.append(float(raw.det_scores[i]))
    xy = np.concatenate(polys)
    vo = np.cumsum([0] + [len(p) for p in polys]).astype(np.int64)
    assert np.array_equal(ring.cpu().numpy(), xy) and np.array_equal(voff.cpu().numpy(), vo)
    ref = oracle.merge_overlap_arrays(xy, vo, np.array(sc, dtype=np.float64), 0.05)
    assert kept.tolist() == list(ref)
    assert 0 < len(ref) < len(polys)          # some nuclei of overlapping tiles were merged away


def test_compact_remaps_per_tile_keep_lists():
    """RoIStageResult.compact(): tile t's kept list sits at keep[tile_start[t] : +tile_count[t]] and tile_start is the scan
    of the ENTRANT counts, so once a tile suppresses a mask the lists are not packed from 0: the remapped lists
    must name the same detections as the raw ones, for every tile of a B > 1 batch."""
    from nuhtc_b200.roi_stage import RoIStage
    cfg, feats, rois, heads = _setup("single", 64, B=3, n_per=250, max_per_img=60)
    gh = copy.copy(heads).to("cuda")
    raw = RoIStage(cfg, gh.bbox_heads(), gh.mask_head).run([f.cuda() for f in feats], rois.cuda())
    res = raw.compact()
    a, b = raw.kept_indices(), res.kept_indices()
    assert len(a) == len(b) == 3 and sum(len(x) for x in a) > 0
    assert any(int(c) < 60 for c in raw.tile_count.cpu())          # some tile did lose masks: the gaps exist
    for ka, kb in zip(a, b):
        assert torch.equal(raw.det_boxes[ka], res.det_boxes[kb]) and torch.equal(raw.det_scores[ka], res.det_scores[kb])
        assert torch.equal(raw.det_tile[ka], res.det_tile[kb])
