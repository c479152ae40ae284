"""GPU contour tracing against the oracle's Suzuki-Abe restatement, the cv2 golden and, when importable,
cv2 itself; bit-exact (integer work)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def _run(masks_u8, max_pts=512):
    from nuhtc_b200 import mask_contours, pack_masks
    m = torch.from_numpy(np.ascontiguousarray(masks_u8)).cuda()
    bits, _, bbox = pack_masks(m)
    xy, cnt, _ = mask_contours(bits, m.shape[2], max_pts=max_pts)
    xy2, cnt2, _ = mask_contours(bits, m.shape[2], max_pts=max_pts, bbox=bbox)
    assert torch.equal(cnt, cnt2)
    live = torch.arange(xy.shape[1], device="cuda")[None, :] < cnt[:, None]          # entries beyond the count are not written
    assert torch.equal(xy[live], xy2[live])
    xy, cnt = xy.cpu().numpy(), cnt.cpu().numpy()
    return [xy[i, :cnt[i]] for i in range(len(cnt))]


def test_golden_cv2():
    z = np.load(os.path.join(G, "contours.npz"))
    cm = np.unpackbits(z["masks"], axis=2)[:, :, :96]
    got = _run(cm)
    off = z["offsets"]
    for i in range(len(cm)):
        ref = z["points"][off[i]: off[i + 1]][:-1]      # the golden holds mask2inst output (first point repeated)
        assert np.array_equal(got[i], ref), i


@pytest.mark.parametrize("h,w", [(40, 40), (64, 64), (96, 130), (256, 256)])
def test_random_masks_vs_oracle(oracle, h, w):
    rng = np.random.default_rng(h * 1000 + w)
    n = 64
    cm = np.zeros((n, h, w), np.uint8)
    for i in range(n):
        kind = i % 4
        if kind == 0:
            cm[i] = rng.random((h, w)) < rng.uniform(0.05, 0.9)                  # whole-frame noise: deep nesting, big window
        elif kind == 1:
            y0, x0 = rng.integers(0, h - 20), rng.integers(0, w - 20)
            cm[i, y0:y0 + 20, x0:x0 + 20] = rng.random((20, 20)) < 0.6            # small window anywhere (word straddling)
        elif kind == 2:
            yy, xx = np.mgrid[:h, :w]
            cy, cx, r = rng.uniform(0, h), rng.uniform(0, w), rng.uniform(2, 18)
            cm[i] = ((yy - cy) ** 2 + (xx - cx) ** 2 <= r * r) & ~((yy - cy) ** 2 + (xx - cx) ** 2 <= (r / 3) ** 2)
        # kind 3: empty
    cm[5] = 1                                                                    # full frame
    got = _run(cm, max_pts=2 * h * w)
    for i in range(n):
        assert np.array_equal(got[i], oracle.contour0(cm[i])), (i, got[i][:6], oracle.contour0(cm[i])[:6])


def test_nuclei_masks_from_paste(oracle):
    """The shape the tile loop produces: pasted + thresholded 28x28 mask logits in a 256x256 frame."""
    from nuhtc_b200 import synth, paste_masks
    from nuhtc_b200.contours import mask_contours
    boxes, probs, _ = synth.nuclei_masks(500, frame=256, seed=9)
    bits, area, bbox = paste_masks(probs.cuda(), boxes.cuda(), 256, 256, thr=0.5, kind="bits", want_stats=True)
    dense = paste_masks(probs.cuda(), boxes.cuda(), 256, 256, thr=0.5, kind="bin").to(torch.uint8)
    xy, cnt, _ = mask_contours(bits, 256, max_pts=256)
    xy2, cnt2, _ = mask_contours(bits, 256, max_pts=256, bbox=bbox)          # tight boxes from the paste kernel
    assert torch.equal(cnt, cnt2)
    xy, cnt, dense, xy2 = xy.cpu().numpy(), cnt.cpu().numpy(), dense.cpu().numpy(), xy2.cpu().numpy()
    assert all(np.array_equal(xy[i, :cnt[i]], xy2[i, :cnt[i]]) for i in range(len(cnt)))
    assert cnt.max() > 8
    for i in range(len(cnt)):
        assert np.array_equal(xy[i, :cnt[i]], oracle.contour0(dense[i]))


def test_overflow_is_reported():
    from nuhtc_b200 import NuhtcError
    m = np.zeros((2, 64, 64), np.uint8)
    m[1, ::2, ::2] = 1; m[1, 10:50, 10:50] = np.indices((40, 40)).sum(0) % 2       # long contour
    m[0, 3:9, 3:9] = 1
    with pytest.raises(NuhtcError):
        _run(m, max_pts=4 - 1)
    got = _run(m, max_pts=4096)
    assert len(got[0]) == 4


def test_rings_and_mask2inst(oracle):
    from nuhtc_b200 import mask_contours, pack_masks, rings_for_merge, mask2inst
    rng = np.random.default_rng(3)
    cm = np.zeros((12, 64, 80), np.uint8)
    for i in range(12):
        y0, x0 = rng.integers(0, 40), rng.integers(0, 50)
        cm[i, y0:y0 + rng.integers(2, 20), x0:x0 + rng.integers(2, 25)] = 1
    cm[4] = 0; cm[4, 7, 7] = 1          # single pixel -> closed contour of 2 points -> dropped (infer_wsi.py:530)
    cm[6] = 0                           # empty
    bits, _, _ = pack_masks(torch.from_numpy(cm).cuda())
    xy, cnt, _ = mask_contours(bits, 80)
    origin = torch.from_numpy(rng.integers(0, 5000, (12, 2)).astype(np.int32)).cuda()
    sel = torch.ones(12, dtype=torch.bool, device="cuda"); sel[9] = False
    ring, voff, index = rings_for_merge(xy, cnt, sel, origin)
    assert index.cpu().tolist() == [i for i in range(12) if i not in (4, 6, 9)]
    ring, voff = ring.cpu().numpy(), voff.cpu().numpy()
    for k, i in enumerate(index.cpu().tolist()):
        ref = oracle.mask2inst(cm[i]).reshape(-1, 2) + origin[i].cpu().numpy()
        assert np.array_equal(ring[voff[k]: voff[k + 1]], ref.astype(np.float64))
    assert np.array_equal(mask2inst(cm[0]), oracle.mask2inst(cm[0]))


def test_full_size_contour_properties():
    """8000 nucleus masks in 256x256 frames (the bench's detection count), no oracle: every contour vertex is a set pixel,
    a contour exists iff the mask is non-empty, and for a hole-free single blob Pick's theorem ties the traced polygon to the
    pixel count: 2 * area(polygon) = 2 * pixels - B - 2 with B the lattice points on the polygon's edges."""
    from nuhtc_b200 import synth, paste_masks, mask_contours
    boxes, probs, _ = synth.nuclei_masks(8000, frame=256, seed=4)
    bits, area, bbox = paste_masks(probs.cuda(), boxes.cuda(), 256, 256, thr=0.5, kind="bits", want_stats=True)
    dense = paste_masks(probs.cuda(), boxes.cuda(), 256, 256, thr=0.5, kind="bin")
    xy, cnt, _ = mask_contours(bits, 256, max_pts=256, bbox=bbox)
    xy2, cnt2, _ = mask_contours(bits, 256, max_pts=256)                 # idempotent, with or without the tight boxes
    assert torch.equal(cnt, cnt2)
    assert torch.equal((cnt > 0), (area > 0))
    n, P = xy.shape[0], xy.shape[1]
    live = torch.arange(P, device="cuda")[None, :] < cnt[:, None]
    assert torch.equal(xy[live], xy2[live])
    idx = torch.arange(n, device="cuda")[:, None].expand(n, P)[live]
    pts = xy[live].long()
    assert dense[idx, pts[:, 1], pts[:, 0]].all()                        # vertices sit on set pixels
    # Pick's theorem on the closed polygon
    x = torch.where(live, xy[..., 0], torch.zeros_like(xy[..., 0])).double()
    y = torch.where(live, xy[..., 1], torch.zeros_like(xy[..., 1])).double()
    last = (cnt.long() - 1).clamp(min=0)
    nxt = (torch.arange(P, device="cuda")[None, :] + 1) % P
    xn = torch.where(torch.arange(P, device="cuda")[None, :] == last[:, None], x[:, :1], torch.gather(x, 1, nxt.expand(n, P)))
    yn = torch.where(torch.arange(P, device="cuda")[None, :] == last[:, None], y[:, :1], torch.gather(y, 1, nxt.expand(n, P)))
    cross = torch.where(live, x * yn - xn * y, torch.zeros_like(x)).sum(1)
    edge_pts = torch.where(live, torch.gcd((xn - x).abs().long(), (yn - y).abs().long()), torch.zeros_like(pts.new_zeros(1)).expand(n, P)).sum(1)
    twice_area = cross.abs().round().long()
    ok = twice_area == 2 * area.long() - edge_pts - 2
    big = area > 30
    assert ok[big].float().mean().item() > 0.97                          # blobs without holes / 1-px spurs
