"""CPU: the oracle's restatement of the watershed proposals against the golden made from the reference's own method bodies
(tests/golden/make_golden_watershed.py executes _watershed_proposal / binary_open / _inst_mask_to_bbox of
nuhtc/models/htc_roi_head_cus.py)."""
import os

import numpy as np
import pytest
import torch

G = os.path.join(os.path.dirname(__file__), "golden", "watershed.npz")


def test_oracle_matches_reference_golden(oracle):
    z = np.load(G)
    sem = torch.from_numpy(z["semantic_pred"])
    B = sem.shape[0]
    assert float(z["blurred_absmin"]) > 1e-5            # no pixel of the golden input sits on the threshold
    m = oracle.watershed_semantic_mask(sem, (256, 256), 0.0).numpy()
    assert np.array_equal(m.astype(np.uint8), z["mask"])
    plist, ws = oracle.watershed_proposal(sem, [torch.from_numpy(z[f"props{i}"]) for i in range(B)], (256, 256))
    total = 0
    for i in range(B):
        assert np.array_equal(ws[i].numpy(), z[f"ws{i}"])
        assert np.array_equal(plist[i].numpy(), z[f"plist{i}"])
        filled = oracle.watershed_instances(m[i])[1]
        assert filled.sum() > m[i].sum()                # the golden masks do have holes to fill
        total += len(ws[i])
    assert total >= 40


def test_gaussian_restatement_matches_torchvision(oracle):
    TF = pytest.importorskip("torchvision.transforms.functional")
    x = torch.randn(2, 1, 40, 56)
    assert torch.equal(oracle.gaussian_blur5(x), TF.gaussian_blur(x, kernel_size=5))
    from nuhtc_b200.watershed import gaussian_blur5
    assert torch.equal(gaussian_blur5(x), TF.gaussian_blur(x, kernel_size=5))


def test_instances_edge_cases(oracle):
    e = np.zeros((32, 32), np.uint8)
    assert oracle.watershed_instances(e)[0].shape == (0, 5)
    e[:] = 1                                             # one component of the whole frame: area >= H*W/4 -> dropped
    assert oracle.watershed_instances(e)[0].shape == (0, 5)
    d = np.zeros((32, 32), np.uint8)
    d[2:7, 2:7] = 1
    d[7:12, 7:12] = 1                                    # touches the first one only diagonally: two instances (4-connectivity)
    b = oracle.watershed_instances(d)[0]
    assert b.tolist() == [[2, 2, 7, 7, 1], [7, 7, 12, 12, 1]]


def test_product_torch_half_matches_golden_mask_on_cpu():
    """The device half of the product (upsample, blur, threshold, opening) is plain torch and also runs on CPU tensors: it must
    reproduce the mask the reference's code produced for the golden input.  (The component half has no CPU path.)"""
    from nuhtc_b200.watershed import semantic_mask, mask_components
    z = np.load(G)
    m = semantic_mask(torch.from_numpy(z["semantic_pred"]), (256, 256), 0.0)
    assert np.array_equal(m[:, 0].numpy().astype(np.uint8), z["mask"])
    with pytest.raises(Exception):
        mask_components(m)
