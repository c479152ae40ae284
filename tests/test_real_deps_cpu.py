"""When the box happens to have the reference's native dependencies, diff the oracle against them directly (SURVEY.md 8c:
"the test-suite must auto-detect and additionally diff against them").  mmcv-full 1.7.2, pycocotools 2.0.7 and shapely >= 2
are absent from the build container and from the GPU image, where these tests skip; the oracle is otherwise anchored on
torchvision's CPU ops, brute-force pixel counts, exact rational arithmetic and an independent slab-decomposition area."""
import numpy as np
import pytest
import torch

from nuhtc_b200 import synth


def test_mmcv_roi_align_and_nms(oracle):
    ops = pytest.importorskip("mmcv.ops")
    torch.manual_seed(0)
    x = torch.randn(2, 8, 40, 40)
    rois = synth.proposals(2, 200, "routed", frame=160, seed=1)
    for P, sr in ((7, 0), (7, 2), (14, 0)):
        ref = ops.roi_align(x, rois, (P, P), 0.25, sr, "avg", True)
        assert torch.equal(oracle.roi_align(x, rois, P, 0.25, sr), ref)
    boxes, scores, labels = synth.nms_boxes(12000, seed=2)
    for n in (3000, 12000):       # both sides of split_thr
        b, s, l = boxes[:n], scores[:n], labels[:n]
        d_ref, k_ref = ops.batched_nms(b, s, l, dict(type="nms", iou_threshold=0.5))
        d, k = oracle.batched_nms(b, s, l, dict(type="nms", iou_threshold=0.5))
        assert torch.equal(k, k_ref) and torch.equal(d, d_ref)
    d_ref, k_ref = ops.nms(boxes[:5000], scores[:5000], 0.4)
    d, k = oracle.nms(boxes[:5000], scores[:5000], 0.4)
    assert torch.equal(k, k_ref)


def test_pycocotools_rle_iou(oracle):
    mu = pytest.importorskip("pycocotools.mask")
    b, probs, sc = synth.nuclei_masks(120, seed=3)
    masks = (oracle.paste_masks(probs, b, 256, 256) >= 0.5).numpy().astype(np.uint8)
    rles = [mu.encode(np.asfortranarray(m)) for m in masks]
    ref = mu.iou(rles, rles, [0] * len(rles))
    assert np.array_equal(oracle.mask_iou(masks), ref)          # integer pixel counts, one double division
    from nuhtc_b200.mask_nms import rle_encode
    for m, r in zip(masks, rles):
        assert rle_encode(m) == {"size": list(r["size"]), "counts": r["counts"]}      # the compressed byte string itself


def test_shapely_polygon_iou_and_merge(oracle):
    sh = pytest.importorskip("shapely.geometry")
    d = synth.slide_nuclei(6, 6, per_tile=23, seed=4)
    N = len(d["score"])
    rings = [d["xy"][d["voff"][i]:d["voff"][i + 1]] for i in range(N)]
    polys = [sh.Polygon(r) for r in rings]
    checked = 0
    for i in range(N):
        for j in range(i + 1, min(N, i + 60)):
            if not polys[i].intersects(polys[j]):
                continue
            inter = polys[i].intersection(polys[j]).area
            assert abs(oracle.poly_inter_area(rings[i], rings[j]) - inter) <= 1e-9 * max(1.0, polys[i].area)
            checked += 1
    assert checked > 20
    # the greedy loop itself, with shapely's IoU, in score order (nuclei_merge.py:114-150, 'probability')
    order = sorted(range(N), key=lambda i: (-d["score"][i], i))
    dead, kept = set(), []
    for ai, a in enumerate(order):
        if a in dead:
            continue
        kept.append(a)
        for b in order[ai + 1:]:
            if b in dead or not polys[a].intersects(polys[b]):
                continue
            inter = polys[a].intersection(polys[b]).area
            if inter / (polys[a].area + polys[b].area - inter) > 0.05:
                dead.add(b)
    assert oracle.merge_overlap_arrays(d["xy"], d["voff"], d["score"], 0.05).tolist() == kept
