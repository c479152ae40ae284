"""Watershed proposals of the HTC RoI heads, device resident.

Mirror of ``HybridTaskCascadeRoIHead_Cus._watershed_proposal(semantic_pred, semantic_dist, proposal_list, img_shape,
min_area, thres, sample_num)`` (/root/reference/nuhtc/models/htc_roi_head_cus.py:283-342) for the test path
(``simple_test`` :2217-2221 calls it with ``min_area=10, thres=0`` and no ``semantic_dist`` / ``sample_num``).

The reference runs the first half on the device with stock torch ops -- bilinear upsampling (align_corners), a 5x5 Gaussian
blur, the threshold and two erosions + two dilations with a 5x5 box as ``conv2d`` -- and these stay the same torch ops here
(same kernels, same bits).  Its second half copies every mask to the host and goes through scipy, skimage and Python loops
over instances; that half is ONE C-ABI call here (``nuhtc_mask_components``, csrc/ccl.cu): hole filling, labelling, areas,
the area filter and the boxes, no host round trip except the per-image box counts.
"""
from __future__ import annotations

import ctypes
import math
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from . import _lib as L

__all__ = ["semantic_mask", "mask_components", "watershed_proposal"]


def _gaussian_kernel2d(ksize: int, dtype, device) -> torch.Tensor:
    """torchvision.transforms.functional.gaussian_blur's kernel for sigma=None (torchvision/transforms/_functional_tensor.py:
    _get_gaussian_kernel1d/2d): sigma = 0.3 * ((k - 1) * 0.5 - 1) + 0.8."""
    sigma = 0.3 * ((ksize - 1) * 0.5 - 1) + 0.8
    half = (ksize - 1) * 0.5
    x = torch.linspace(-half, half, steps=ksize, dtype=dtype, device=device)
    pdf = torch.exp(-0.5 * (x / sigma).pow(2))
    k1 = pdf / pdf.sum()
    return torch.mm(k1[:, None], k1[None, :])


def gaussian_blur5(x: torch.Tensor) -> torch.Tensor:
    """TF.gaussian_blur(x, kernel_size=5) on [B,C,H,W]: reflect padding 2, depthwise conv2d."""
    C = x.shape[1]
    k = _gaussian_kernel2d(5, x.dtype, x.device).expand(C, 1, 5, 5)
    return F.conv2d(F.pad(x, [2, 2, 2, 2], mode="reflect"), k, groups=C)


def semantic_mask(semantic_pred: torch.Tensor, img_shape: Sequence[int], thres: float = 0.0, kernel: Optional[torch.Tensor] = None) -> torch.Tensor:
    """htc_roi_head_cus.py:285-299: upsample, blur, threshold, binary_open(kernel 5x5 ones, 2 iterations) -> [B,1,H,W] of 0/1."""
    m = F.interpolate(semantic_pred, size=tuple(int(v) for v in img_shape[:2]), mode="bilinear", align_corners=True)
    m = gaussian_blur5(m)
    m = (m > thres).to(semantic_pred.dtype)
    if kernel is None:
        kernel = torch.ones((1, 1, 5, 5), dtype=m.dtype, device=m.device)
    pad = kernel.shape[-1] // 2
    ksum = kernel.sum()
    for _ in range(2):   # binary_erosion :239-244
        m = torch.clamp(F.conv2d(m, kernel, padding=pad) - ksum + 1, min=0, max=1)
    for _ in range(2):   # binary_dilate :246-251
        m = torch.clamp(F.conv2d(m, kernel, padding=pad), min=0, max=1)
    return m


def mask_components(mask: torch.Tensor, min_area: int = 10, max_area: Optional[float] = None, max_boxes: int = 4096,
                    return_filled: bool = False):
    """mask [B,H,W] (or [B,1,H,W]) fp32 on CUDA -> (boxes [B,max_boxes,5], counts [B] int32[, filled [B,H,W] uint8])."""
    L.require_cuda(mask, "mask")
    if mask.dim() == 4:
        mask = mask[:, 0]
    mask = mask.contiguous().to(torch.float32)
    B, H, W = mask.shape
    dev = mask.device
    if max_area is None:
        max_area = H * W / 4
    boxes = torch.zeros((B, max_boxes, 5), dtype=torch.float32, device=dev)
    counts = torch.zeros(B, dtype=torch.int32, device=dev)
    filled = torch.empty((B, H, W), dtype=torch.uint8, device=dev) if return_filled else None
    lib = L.lib()
    wsb = lib.nuhtc_mask_components_workspace_bytes(B, H, W)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.nuhtc_mask_components(mask.data_ptr(), B, H, W, int(min_area), int(math.ceil(max_area)), int(max_boxes), boxes.data_ptr(),
                                       counts.data_ptr(), L.ptr(filled), ws.data_ptr(), wsb, L.stream_ptr(dev))
    L.check(rc, "mask_components")
    L.count("components")
    return (boxes, counts, filled) if return_filled else (boxes, counts)


@torch.no_grad()
def watershed_proposal(semantic_pred: torch.Tensor, semantic_dist=None, proposal_list: Optional[List[torch.Tensor]] = None,
                       img_shape=None, min_area: int = 10, thres: float = 0, sample_num=None,
                       max_boxes: int = 4096) -> Tuple[Optional[List[torch.Tensor]], List[torch.Tensor]]:
    """Same arguments and return value as the reference method: (proposal_list with the watershed boxes prepended per
    image, the watershed boxes per image as fp32 [n,5] = (x0, y0, x1+1, y1+1, 1.0))."""
    if semantic_dist is not None:
        raise NotImplementedError("semantic_dist (the optional seg_head's distance map) is not configured by any shipped NuHTC "
                                  "config; with the Euclidean distance the watershed reduces to the components computed here")
    if sample_num is not None:
        raise NotImplementedError("sample_num is the training-time resampling of the boxes (forward_train); the test path passes None")
    m = semantic_mask(semantic_pred, img_shape, thres)
    boxes, counts = mask_components(m, min_area=min_area, max_boxes=max_boxes)
    n = counts.cpu().tolist()          # the one host read: how many boxes each image got
    if max(n, default=0) > max_boxes:
        raise L.NuhtcError(f"watershed_proposal: {max(n)} instances in one image exceed max_boxes={max_boxes}")
    ws = [boxes[i, : n[i]] for i in range(len(n))]
    if proposal_list is not None:
        proposal_list = list(proposal_list)
        for i, w in enumerate(ws):
            if w.shape[0]:
                proposal_list[i] = torch.cat((w.to(proposal_list[i].dtype), proposal_list[i]), dim=0)
    return proposal_list, ws
