"""GPU: watershed proposals (nuhtc_b200.watershed, csrc/ccl.cu) against the reference-source golden and the oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden", "watershed.npz")


def test_watershed_proposal_matches_reference_golden():
    from nuhtc_b200.watershed import semantic_mask, watershed_proposal
    z = np.load(G)
    sem = torch.from_numpy(z["semantic_pred"]).cuda()
    B = sem.shape[0]
    m = semantic_mask(sem, (256, 256), 0.0)
    assert np.array_equal(m[:, 0].cpu().numpy().astype(np.uint8), z["mask"])     # min |blurred| of the golden input is 3e-5
    plist, ws = watershed_proposal(sem, proposal_list=[torch.from_numpy(z[f"props{i}"]).cuda() for i in range(B)], img_shape=(256, 256, 3),
                                   min_area=10, thres=0)
    for i in range(B):
        assert ws[i].dtype == torch.float32
        assert np.array_equal(ws[i].cpu().numpy(), z[f"ws{i}"])
        assert np.array_equal(plist[i].cpu().numpy(), z[f"plist{i}"])


def _random_masks(B, H, W, seed):
    g = np.random.default_rng(seed)
    m = np.zeros((B, H, W), np.float32)
    ys, xs = np.mgrid[0:H, 0:W]
    for b in range(B):
        for _ in range(int(g.integers(5, 60))):
            cy, cx = g.uniform(0, H), g.uniform(0, W)
            ry, rx = g.uniform(2, H / 6), g.uniform(2, W / 6)
            d = ((ys - cy) / ry) ** 2 + ((xs - cx) / rx) ** 2
            if g.random() < 0.4:
                m[b][(d < 1) & (d > 0.45)] = 1                    # rings: holes, sometimes with islands inside
            else:
                m[b][d < 1] = 1
        noise = g.random((H, W))
        m[b][noise < 0.01] = 1                                     # specks (below min_area) and diagonal contacts
        m[b][noise > 0.995] = 0                                    # pin holes
    return m


@pytest.mark.parametrize("B,H,W,seed", [(4, 256, 256, 0), (3, 512, 512, 1), (5, 96, 160, 2), (1, 17, 33, 3)])
def test_components_vs_oracle(oracle, B, H, W, seed):
    from nuhtc_b200.watershed import mask_components
    m = _random_masks(B, H, W, seed)
    boxes, counts, filled = mask_components(torch.from_numpy(m).cuda(), min_area=10, return_filled=True)
    counts = counts.cpu().numpy()
    n_total = 0
    for b in range(B):
        ref_boxes, ref_filled = oracle.watershed_instances(m[b], 10)
        assert np.array_equal(filled[b].cpu().numpy().astype(bool), ref_filled)
        assert counts[b] == len(ref_boxes)
        assert np.array_equal(boxes[b, : counts[b]].cpu().numpy(), ref_boxes)
        n_total += len(ref_boxes)
    assert n_total > 0 or H * W < 1000       # the tiny frame checks the index arithmetic, its blobs rarely qualify


def test_components_edge_cases(oracle):
    from nuhtc_b200.watershed import mask_components
    z = torch.zeros(2, 64, 64).cuda()
    assert mask_components(z)[1].tolist() == [0, 0]
    assert mask_components(torch.ones(2, 64, 64).cuda())[1].tolist() == [0, 0]        # area >= H*W/4
    d = torch.zeros(1, 32, 32)
    d[0, 2:7, 2:7] = 1
    d[0, 7:12, 7:12] = 1
    boxes, counts = mask_components(d.cuda())
    assert counts.tolist() == [2] and boxes[0, :2].cpu().tolist() == [[2, 2, 7, 7, 1], [7, 7, 12, 12, 1]]
    # a frame-touching ring: its inside is a hole (filled), the outside is not
    r = torch.zeros(1, 48, 48)
    r[0, 0:20, 5:25] = 1
    r[0, 4:16, 9:21] = 0
    boxes, counts, filled = mask_components(r.cuda(), return_filled=True)
    assert counts.tolist() == [1] and int(filled.sum()) == 400
    # more instances than max_boxes: the count says so
    many = torch.zeros(1, 128, 128)
    for i in range(0, 120, 8):
        for j in range(0, 120, 8):
            many[0, i:i + 5, j:j + 5] = 1
    boxes, counts = mask_components(many.cuda(), max_boxes=16)
    assert counts.tolist() == [225]
    ref = oracle.watershed_instances(many[0].numpy(), 10)[0]
    assert np.array_equal(boxes[0].cpu().numpy(), ref[:16])
    with pytest.raises(Exception):
        mask_components(torch.zeros(1, 8, 8))             # CPU tensor: no fallback


def test_head_prepends_watershed_proposals():
    """HybridTaskCascadeRoIHead_Lite(watershed_proposal=True) == the same head fed with the proposals the standalone call returns."""
    import _toy_heads as T
    from nuhtc_b200 import synth
    from nuhtc_b200.htc_roi_head import HybridTaskCascadeRoIHead_Lite, seesaw_activation
    from nuhtc_b200.watershed import watershed_proposal
    Bn, FRAME = 2, 512
    feats = [f.cuda() for f in synth.fpn_levels(Bn, T.C, frame=FRAME, seed=41)]
    props = synth.proposals(Bn, 60, "nuclei", frame=FRAME, seed=42)
    plist = [torch.cat([props[props[:, 0] == b][:, 1:], torch.ones(int((props[:, 0] == b).sum()), 1)], 1).cuda() for b in range(Bn)]
    metas = [dict(img_shape=(FRAME, FRAME, 3), ori_shape=(256, 256, 3), scale_factor=np.array([2., 2., 2., 2.], dtype=np.float32), flip=False)] * Bn
    heads = [T.ToyBBoxHead(i).cuda() for i in range(3)]
    for h in heads:
        h.score_activation = seesaw_activation
    mh, sh = T.ToyMaskHead().cuda(), T.ToySemanticHead().cuda()
    kw = dict(extractor="attention", semantic_head=sh, bbox_roi_layer=dict(type="RoIAlign", output_size=7, sampling_ratio=2))
    with_ws = HybridTaskCascadeRoIHead_Lite(3, heads, mh, T.TEST_CFG, watershed_proposal=True, **kw)
    plain = HybridTaskCascadeRoIHead_Lite(3, heads, mh, T.TEST_CFG, **kw)
    sem_pred, _ = sh(feats)
    plist2, ws = watershed_proposal(sem_pred, proposal_list=plist, img_shape=(FRAME, FRAME), min_area=10, thres=0)
    a = with_ws.simple_test(None, tuple(feats), plist, metas, rescale=True)
    b = plain.simple_test(None, tuple(feats), plist2, metas, rescale=True)
    for (ba, sa), (bb, sb) in zip(a, b):
        assert all(np.array_equal(x, y) for x, y in zip(ba, bb))
