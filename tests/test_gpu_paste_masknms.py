"""GPU parity: mask paste vs the reference's own grid_sample formulation (torch CPU), per-tile mask NMS vs the
pycocotools restatement.  Tolerances from BASELINE.json: probabilities 1e-5 abs; binary masks identical except
at pixels within 1e-6 of the threshold; mask-NMS indices bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _boxes_with_edges(n, frame, seed):
    from nuhtc_b200 import synth
    boxes, probs, scores = synth.nuclei_masks(n, frame=frame, seed=seed)
    boxes[0] = torch.tensor([-10.0, -7.5, 14.2, 9.9])                  # hangs off the top-left corner
    boxes[1] = torch.tensor([frame - 9.3, frame - 20.0, frame + 12.0, frame + 3.0])  # off the bottom-right
    boxes[2] = torch.tensor([30.0, 40.0, 30.0, 60.0])                  # zero width  -> inf/NaN grid, reference gives 0
    boxes[3] = torch.tensor([0.0, 0.0, float(frame), float(frame)])    # whole frame
    boxes[4] = torch.tensor([50.25, 60.75, 51.0, 61.5])                # sub-pixel box
    boxes[5] = torch.tensor([-300.0, -300.0, -200.0, -250.0])          # fully outside
    return boxes, probs, scores


@pytest.mark.parametrize("frame", [256, 100])
def test_paste_prob_and_binary(oracle, frame):
    import nuhtc_b200 as nb
    boxes, probs, _ = _boxes_with_edges(64, frame, seed=frame)
    ref = oracle.paste_masks(probs, boxes, frame, frame)
    out = nb.paste_masks(probs.cuda(), boxes.cuda(), frame, frame, kind="prob").cpu()
    assert out.shape == ref.shape
    assert (out - ref).abs().max().item() <= 1e-5
    thr = 0.5
    b, area, bbox = nb.paste_masks(probs.cuda(), boxes.cuda(), frame, frame, thr=thr, kind="bin", want_stats=True)
    b = b.cpu()
    assert b.dtype == torch.bool
    rb = ref >= thr
    diff = b != rb
    assert ((ref - thr).abs()[diff] <= 1e-6).all()
    assert torch.equal(area.cpu().long(), b.sum((1, 2)))
    for i in range(b.shape[0]):
        ys, xs = np.nonzero(b[i].numpy())
        exp = [0, 0, 0, 0] if len(ys) == 0 else [xs.min(), ys.min(), xs.max() + 1, ys.max() + 1]
        assert bbox[i].cpu().tolist() == exp
    # bit rows carry the same mask
    bits = nb.paste_masks(probs.cuda(), boxes.cuda(), frame, frame, thr=thr, kind="bits").cpu().numpy().view(np.uint64)
    unpacked = ((bits[:, :, :, None] >> np.arange(64, dtype=np.uint64)) & 1).astype(bool).reshape(b.shape[0], frame, -1)[:, :, :frame]
    assert (unpacked == b.numpy()).all()


def test_do_paste_mask_contract(oracle):
    import nuhtc_b200 as nb
    boxes, probs, _ = _boxes_with_edges(32, 256, seed=3)
    boxes = boxes[6:]; probs = probs[6:]
    full, sl = nb._do_paste_mask(probs.cuda(), boxes.cuda(), 256, 256, skip_empty=False)
    assert sl == () and full.shape == (26, 256, 256)
    part, sl = nb._do_paste_mask(probs.cuda(), boxes.cuda(), 256, 256, skip_empty=True)
    assert isinstance(sl[0], slice) and torch.equal(part, full[(slice(None),) + sl])
    # zero threshold: the zero padding passes `>= 0` too (docstring example, fcn_mask_head.py:204-226)
    b0 = nb.paste_masks(probs.cuda(), boxes.cuda(), 256, 256, thr=0.0, kind="bin")
    assert bool(b0.all())


def test_get_seg_masks_matches_reference_flow(oracle):
    import nuhtc_b200 as nb
    boxes, probs, scores = _boxes_with_edges(40, 256, seed=11)
    logits = torch.log(probs / (1 - probs)).clamp(-20, 20)
    det = torch.cat([boxes * 2.0, scores[:, None]], 1)        # network frame (scale_factor 2), rescale=True
    labels = torch.randint(0, 5, (40,))
    sf = np.array([2.0, 2.0, 2.0, 2.0], dtype=np.float32)
    ref = oracle.get_seg_masks(logits.sigmoid(), det, 256, 256, sf, True, 0.5)
    out = nb.get_seg_masks(logits.cuda(), det.cuda(), labels.cuda(), dict(mask_thr_binary=0.5), (256, 256, 3), sf, True, 5)
    assert len(out) == 5 and sum(map(len, out)) == 40
    refp = oracle.paste_masks(logits.sigmoid(), det[:, :4] / 2.0, 256, 256)
    for c in range(5):
        idx = (labels == c).nonzero().squeeze(1).tolist()
        for j, i in enumerate(idx):
            diff = torch.from_numpy(out[c][j]) != ref[i]
            assert ((refp[i] - 0.5).abs()[diff] <= 1e-6).all()


def test_paste_full_size_roundtrip():
    """BASELINE cfg-4 size (16 tiles x 500 masks): dense uint8 output == unpacked bit output, area == popcount."""
    import nuhtc_b200 as nb
    from nuhtc_b200 import synth
    boxes, probs, _ = synth.nuclei_masks(8000, seed=0)
    b, area, bbox = nb.paste_masks(probs.cuda(), boxes.cuda(), 256, 256, thr=0.5, kind="bin", want_stats=True)
    bits, area2, bbox2 = nb.paste_masks(probs.cuda(), boxes.cuda(), 256, 256, thr=0.5, kind="bits", want_stats=True)
    pb, pa, pbb = nb.pack_masks(b)
    assert torch.equal(pb, bits) and torch.equal(pa, area) and torch.equal(area, area2)
    assert torch.equal(pbb, bbox) and torch.equal(bbox, bbox2)
    assert torch.equal(area.long(), b.sum((1, 2)))


@pytest.mark.parametrize("n,frame", [(1, 64), (40, 64), (300, 256), (500, 256)])
def test_mask_nms_exact(oracle, n, frame):
    import nuhtc_b200 as nb
    from nuhtc_b200 import synth
    boxes, probs, scores = synth.nuclei_masks(n, frame=frame, seed=n)
    masks = (oracle.paste_masks(probs, boxes, frame, frame) >= 0.5).numpy().astype(np.uint8)
    if n > 4:
        masks[3] = 0  # an empty mask
    ref = oracle.mask_nms(masks, scores.numpy(), thr=0.05)
    rles, idx = nb.mask_nms(masks, scores.numpy(), thr=0.05)
    assert idx.dtype == np.int64 and (idx == ref).all()
    assert len(rles) == len(idx)
    for r, i in zip(rles, idx):
        from nuhtc_b200.mask_nms import rle_decode
        assert isinstance(r["counts"], bytes) and np.array_equal(rle_decode(r), np.asarray(masks[i]).astype(np.uint8))
    for thr in (0.0, 0.5, 0.9):
        assert (nb.mask_nms(masks, scores.numpy(), thr=thr)[1] == oracle.mask_nms(masks, scores.numpy(), thr=thr)).all()


def test_mask_nms_odd_width_and_tiles(oracle):
    import nuhtc_b200 as nb
    from nuhtc_b200 import synth
    frame = 100  # not a multiple of 64: generic pack path
    T, n = 5, 60
    ms, ss, ts = [], [], []
    for t in range(T):
        boxes, probs, scores = synth.nuclei_masks(n + t, frame=frame, seed=50 + t)
        ms.append((oracle.paste_masks(probs, boxes, frame, frame) >= 0.5).numpy().astype(np.uint8))
        ss.append(scores.numpy()); ts.append(np.full(n + t, t, np.int32))
    M, S, Tt = np.concatenate(ms), np.concatenate(ss), np.concatenate(ts)
    perm = np.random.default_rng(0).permutation(len(S))
    M, S, Tt = M[perm], S[perm], Tt[perm]
    bits, area, bbox = nb.pack_masks(torch.from_numpy(M).cuda())
    assert (area.cpu().numpy() == oracle.mask_area(M)).all()
    keep, tstart, tcount, status = nb.mask_nms_device(bits, area, bbox, torch.from_numpy(S).cuda(), frame, 0.05,
                                                      tile=torch.from_numpy(Tt).cuda(), num_tiles=T, max_tile_size=n + T)
    assert int(status.item()) == 0
    keep, tstart, tcount = keep.cpu().numpy(), tstart.cpu().numpy(), tcount.cpu().numpy()
    for t in range(T):
        idx = np.nonzero(Tt == t)[0]
        ref = idx[oracle.mask_nms(M[idx], S[idx], thr=0.05)]
        assert (keep[tstart[t]: tstart[t] + tcount[t]] == ref).all()


def test_dense_and_bits_from_one_evaluation():
    """nuhtc_paste_masks_dense_bits == the two single-output calls (frames, bit rows, area, tight boxes)."""
    import nuhtc_b200 as nb
    from nuhtc_b200 import synth
    from nuhtc_b200.mask_paste import paste_masks_dense_bits
    boxes, probs, _ = synth.nuclei_masks(300, frame=256, seed=21)
    boxes[0] = torch.tensor([-50.0, -50.0, -20.0, -10.0])      # outside
    boxes[1] = torch.tensor([200.0, 180.0, 300.0, 290.0])      # clipped
    boxes[2] = torch.tensor([0.0, 0.0, 256.0, 256.0])          # whole frame
    for thr in (0.5, 0.0):
        dense, bits, area, bbox = paste_masks_dense_bits(probs.cuda(), boxes.cuda(), 256, 256, thr)
        d1, a1, b1 = nb.paste_masks(probs.cuda(), boxes.cuda(), 256, 256, thr=thr, kind="bin", want_stats=True)
        w1, a2, b2 = nb.paste_masks(probs.cuda(), boxes.cuda(), 256, 256, thr=thr, kind="bits", want_stats=True)
        assert torch.equal(dense, d1) and torch.equal(bits, w1)
        assert torch.equal(area, a1) and torch.equal(area, a2) and torch.equal(bbox, b1) and torch.equal(bbox, b2)
