import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import nuhtc_b200 as nb
from nuhtc_b200 import synth
B, C, P, sr = 2, 64, 7, 0
feats = synth.fpn_levels(B, C, frame=256, seed=3)
rois = synth.proposals(B, 200, "routed", frame=256, seed=4)
out = nb.roi_align_levels([f.cuda() for f in feats], rois.cuda(), P, [1 / s for s in synth.FPN_STRIDES], sr, mode="route", finest_scale=56)
torch.cuda.synchronize()
lit = nb.roi_align_levels([f.cuda() for f in feats], rois.cuda(), P, [1 / s for s in synth.FPN_STRIDES], sr, mode="route", finest_scale=56, impl="direct")
print("max diff", (out - lit).abs().max().item())
