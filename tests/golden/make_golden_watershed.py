#!/usr/bin/env python
"""Golden vectors for the watershed proposals, made by EXECUTING the reference's own method bodies:
`_watershed_proposal`, `binary_erosion`, `binary_dilate`, `binary_open`, `_inst_mask_to_bbox` are AST-extracted from
/root/reference/nuhtc/models/htc_roi_head_cus.py and run on CPU with torch, torchvision (TF.gaussian_blur) and scipy.

skimage is not installed in the build container.  Its `watershed(-distance, markers, mask=mask)` is replaced by a stand-in
that ASSERTS the markers already cover the whole mask and returns them -- with the Euclidean distance as the landscape that
is always the case (distance >= 1 > 0.25 on the mask), and then the flood has nothing to do.  The assertion runs on every
image of the golden set.

    python tests/golden/make_golden_watershed.py     (only in the build container: needs /root/reference)
"""
import ast
import os
import textwrap

import numpy as np
import torch
import torch.nn.functional as F
import torchvision.transforms.functional as TF
from scipy import ndimage as ndi

REF = "/root/reference/nuhtc/models/htc_roi_head_cus.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def watershed(landscape, markers, mask=None):
    assert ((markers > 0) == np.asarray(mask, dtype=bool)).all(), "markers do not cover the mask: a real flood would be needed"
    return markers


def reference_methods():
    tree = ast.parse(open(REF).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "HybridTaskCascadeRoIHead_Cus")
    want = {"_watershed_proposal", "binary_erosion", "binary_dilate", "binary_open", "_inst_mask_to_bbox"}
    src = []
    for n in cls.body:
        if isinstance(n, ast.FunctionDef) and n.name in want:
            n.decorator_list = []
            src.append(textwrap.indent(ast.unparse(n), "    "))
    code = "class Ref:\n" + "\n\n".join(src)
    ns = dict(torch=torch, F=F, TF=TF, ndi=ndi, np=np, watershed=watershed)
    exec(code, ns)
    ref = ns["Ref"]()
    ref.kernel = torch.ones((1, 1, 5, 5))
    return ref


def synthetic_semantic(B=4, h=64, seed=0):
    """Logit maps at stride 4 of a 256 px tile: blobs, rings (holes to fill), blobs that only touch diagonally, specks that the
    opening removes, and a blob on the frame."""
    g = np.random.default_rng(seed)
    ys, xs = np.mgrid[0:h, 0:h].astype(np.float32)
    out = np.full((B, 1, h, h), -3.0, np.float32)
    for b in range(B):
        m = out[b, 0]
        for _ in range(26):
            cy, cx, r = g.uniform(3, h - 3), g.uniform(3, h - 3), g.uniform(1.6, 4.0)
            d = np.hypot(ys - cy, xs - cx)
            m += 6.0 * np.exp(-(d / r) ** 4)
        for _ in range(4):   # rings
            cy, cx, r = g.uniform(10, h - 10), g.uniform(10, h - 10), g.uniform(4.0, 7.0)
            d = np.hypot(ys - cy, xs - cx)
            m += 7.0 * np.exp(-((d - r) / 1.6) ** 2)
        m[0:5, 20:30] += 6.0
        m += g.normal(0, 0.15, size=m.shape).astype(np.float32)
    return torch.from_numpy(out)


def main():
    ref = reference_methods()
    sem = synthetic_semantic()
    img_shape = (256, 256)
    props = [torch.cat([torch.rand(20, 2) * 100, torch.rand(20, 2) * 100 + 120, torch.rand(20, 1)], 1) for _ in range(sem.shape[0])]
    plist, ws = ref._watershed_proposal(sem.clone(), proposal_list=[p.clone() for p in props], img_shape=img_shape, min_area=10, thres=0)
    # the mask the host half starts from (same code as the head of _watershed_proposal)
    m = F.interpolate(sem, size=img_shape, mode="bilinear", align_corners=True)
    m = TF.gaussian_blur(m, kernel_size=5)
    blurred = m.clone()
    m[m > 0] = 1
    m[m <= 0] = 0
    m = ref.binary_open(m, ref.kernel, 2)
    out = dict(semantic_pred=sem.numpy(), mask=m[:, 0].numpy().astype(np.uint8), blurred_absmin=np.float32(blurred.abs().min().item()))
    for i, (w, p, q) in enumerate(zip(ws, props, plist)):
        out[f"ws{i}"] = w.numpy()
        out[f"props{i}"] = p.numpy()
        out[f"plist{i}"] = q.numpy()
        print(f"image {i}: {len(w)} watershed boxes, proposal list {tuple(q.shape)}")
    np.savez_compressed(os.path.join(HERE, "watershed.npz"), **out)
    print("wrote watershed.npz; min |blurred| =", float(out["blurred_absmin"]))


if __name__ == "__main__":
    main()
