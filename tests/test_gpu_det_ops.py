"""Detection glue kernels (det_glue.cu) against the torch chains they replace (bit-exact: same separately rounded fp32
steps) and the oracle / the reference's own known-answer test."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def test_delta2bbox_golden_and_torch_chain(oracle):
    from nuhtc_b200 import det_ops
    from nuhtc_b200.roi_stage import delta2bbox as torch_chain
    z = np.load(os.path.join(G, "delta2bbox.npz"))
    # the reference's KAT (mmdet tests/test_utils/test_coder.py:27-40) through the golden file
    kat = det_ops.delta2bbox(torch.from_numpy(z["kat_rois"]).cuda(), torch.from_numpy(z["kat_deltas"]).cuda(), max_shape=(32, 32, 3))
    assert torch.allclose(kat.cpu(), torch.from_numpy(z["kat_out"]), atol=1e-4)
    rois, deltas = torch.from_numpy(z["rois"]).cuda(), torch.from_numpy(z["deltas"]).cuda()
    out = det_ops.delta2bbox(rois, deltas, stds=(0.1, 0.1, 0.2, 0.2), max_shape=(512, 512, 3))
    assert (out.cpu() - torch.from_numpy(z["out"])).abs().max().item() <= 1e-3
    # bit-identical to the chain of torch CUDA kernels (mul, add, clamp, exp ... each rounded on its own)
    g = torch.Generator().manual_seed(4)
    K = 5000
    r = torch.rand(K, 4, generator=g) * 400
    r[:, 2:] += r[:, :2]
    r5 = torch.cat([torch.randint(0, 16, (K, 1), generator=g).float(), r], 1).cuda()
    d = (torch.randn(K, 4, generator=g) * 2).cuda()
    d[7] = float("nan"); d[9, 2] = 50.0; d[11, 3] = -50.0
    for stds, ms, div in (((0.1, 0.1, 0.2, 0.2), (512, 512), 1.0), ((0.033, 0.033, 0.067, 0.067), (512, 512), 2.0),
                          ((1., 1., 1., 1.), None, 1.0), ((0.05, 0.05, 0.1, 0.1), (300, 517), 1.7)):
        ref = torch_chain(r5[:, 1:], d, stds, ms)
        if div != 1.0:
            ref = ref / torch.tensor(div, device="cuda")       # tensor / tensor: IEEE division like mmdet's rescale
        got = det_ops.delta2bbox(r5, d, stds=stds, max_shape=ms, divide_by=div)
        assert torch.equal(got[:, 0], r5[:, 0])
        assert torch.equal(got[:, 1:].isnan(), ref.isnan())
        assert torch.equal(got[:, 1:].nan_to_num(), ref.nan_to_num())
        got4 = det_ops.delta2bbox(r5[:, 1:].contiguous(), d, stds=stds, max_shape=ms, divide_by=div)
        assert torch.equal(got4.nan_to_num(), ref.nan_to_num())


def test_candidates_slots_filter_vs_torch():
    from nuhtc_b200 import det_ops, nms_groups
    g = torch.Generator().manual_seed(8)
    B, per, C, M = 4, 300, 5, 1600
    K = B * per
    xy = torch.rand(K, 2, generator=g) * 200
    boxes5 = torch.cat([torch.arange(B).repeat_interleave(per).float()[:, None], xy, xy + 5 + torch.rand(K, 2, generator=g) * 30], 1).cuda()
    scores = torch.softmax(torch.randn(K, C + 1, generator=g) * 2, -1).cuda()
    cb, cs, cl, ct, gr = det_ops.multiclass_candidates(boxes5[:, 1:], scores, boxes5, C, 0.05)
    tile = boxes5[:, 0].to(torch.int32)
    assert torch.equal(cb, boxes5[:, None, 1:].expand(K, C, 4).reshape(-1, 4))
    assert torch.equal(cs, scores[:, :C].reshape(-1))
    assert torch.equal(cl, torch.arange(C, device="cuda").repeat(K))
    assert torch.equal(ct, tile.repeat_interleave(C))
    assert torch.equal(gr, torch.where(cs > 0.05, ct, torch.full_like(ct, -1)))
    keep, gstart, gcount, status = nms_groups(cb, cs, cl, gr, B, per * C, 0.5, 0, "offset", num_classes=C)
    db, ds, dl, dt, dv, dc, mr = det_ops.detection_slots(keep, gstart, gcount, M, cb, cs, cl, ct, 2.0)
    r = torch.arange(M, device="cuda")
    valid = (r[None] < gcount.clamp(max=M)[:, None]).reshape(-1)
    idx = (gstart[:, None] + r[None]).clamp(max=keep.numel() - 1).reshape(-1)
    cand = torch.where(valid, keep[idx], torch.zeros_like(idx))
    far = torch.tensor([[-4096.0, -4096.0, -4095.0, -4095.0]], device="cuda")
    assert valid.any() and (~valid).any()
    assert torch.equal(dv, valid) and torch.equal(dc, cand)
    assert torch.equal(db, torch.where(valid[:, None], cb[cand], far))
    assert torch.equal(ds, torch.where(valid, cs[cand], torch.zeros_like(ds)))
    assert torch.equal(dl, cl[cand])
    assert torch.equal(dt, torch.where(valid, ct[cand], torch.full_like(dt, -1)))
    assert torch.equal(mr, torch.cat([dt.clamp(min=0).float()[:, None], db * 2.0], 1))
    area = torch.randint(0, 40, (B * M,), generator=g, dtype=torch.int32).cuda()
    ids = det_ops.tile_filter(db, area, dt, 12, 256, 256, 10)
    ok = (db[:, 0] >= 12) & (db[:, 1] >= 12) & (db[:, 2] <= 256 - 12) & (db[:, 3] <= 256 - 12) & (area >= 10)
    assert torch.equal(ids, torch.where(ok, dt, torch.full_like(dt, -1)))
    assert (ids >= 0).any() and (ids < 0).any()
