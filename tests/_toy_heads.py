"""Seeded toy heads for the RoI-head drop-in tests: small torch modules whose outputs DEPEND on the pooled features (so the
extractor fusion is exercised end to end) and that are built identically by the golden generator
(tests/golden/make_golden_roi_head.py, which drives the REFERENCE simple_test with them) and by the GPU test."""
import torch
import torch.nn as nn
import torch.nn.functional as F

C, NUM_CLASSES = 64, 5
STDS = ((0.1, 0.1, 0.2, 0.2), (0.05, 0.05, 0.1, 0.1), (0.033, 0.033, 0.067, 0.067))


class ToyBBoxHead(nn.Module):
    def __init__(self, stage: int):
        super().__init__()
        g = torch.Generator().manual_seed(100 + stage)
        self.num_classes = NUM_CLASSES
        self.target_stds = STDS[stage]
        self.w_cls = nn.Parameter(torch.randn(NUM_CLASSES + 2, C * 49, generator=g) * 0.05, requires_grad=False)
        b = torch.randn(NUM_CLASSES + 2, generator=g) * 0.5
        b[-2:] = torch.tensor([2.5, -2.5])     # mostly "object": enough detections pass score_thr to make the NMS work
        self.b_cls = nn.Parameter(b, requires_grad=False)
        self.w_reg = nn.Parameter(torch.randn(4, C * 49, generator=g) * 0.004, requires_grad=False)

    def forward(self, feats):
        f = feats.flatten(1)
        return F.linear(f, self.w_cls, self.b_cls), F.linear(f, self.w_reg)


class ToyMaskHead(nn.Module):
    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(200)
        self.num_classes = NUM_CLASSES
        self.class_agnostic = True
        self.w = nn.Parameter(torch.randn(1, C, 1, 1, generator=g) * 0.15, requires_grad=False)
        u = (torch.arange(28, dtype=torch.float32) + 0.5) / 28 * 2 - 1
        self.blob = nn.Parameter((3.0 * (1.0 - (u[:, None] / 0.8) ** 2 - (u[None, :] / 0.8) ** 2))[None, None], requires_grad=False)

    def forward(self, feats, last_feat=None):
        # a nucleus-shaped blob modulated by the pooled features: logits [D,1,28,28]
        return F.interpolate(F.conv2d(feats, self.w), scale_factor=2, mode="nearest") + self.blob


class ToySemanticHead(nn.Module):
    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(300)
        self.w = nn.Parameter(torch.randn(C, C, 1, 1, generator=g) * 0.1, requires_grad=False)
        self.wp = nn.Parameter(torch.randn(1, C, 1, 1, generator=g) * 0.1, requires_grad=False)

    def forward(self, x):
        feat = F.conv2d(x[0], self.w)
        return F.conv2d(feat, self.wp), feat


TEST_CFG = dict(score_thr=0.35, nms=dict(type="nms", iou_threshold=0.5), max_per_img=100, mask_thr_binary=0.5)
