"""Time nuhtc_mask_contours on nucleus-like masks (8000 masks in 256x256 frames, the bench's detection count)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from nuhtc_b200 import synth, paste_masks, mask_contours

boxes, probs, _ = synth.nuclei_masks(8000, frame=256, seed=1)
bits, area, bbox = paste_masks(probs.cuda(), boxes.cuda(), 256, 256, thr=0.5, kind="bits", want_stats=True)
for _ in range(3):
    xy, cnt, st = mask_contours(bits, 256, max_pts=256)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    xy, cnt, st = mask_contours(bits, 256, max_pts=256, check=False, bbox=bbox)
e1.record()
torch.cuda.synchronize()
print(f"mask_contours 8000 masks: {e0.elapsed_time(e1) / 20:.3f} ms/call, mean points {cnt.float().mean().item():.1f}, max {int(cnt.max())}")
