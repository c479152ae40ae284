import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import nuhtc_b200 as nb
from nuhtc_b200 import synth
P = int(sys.argv[1]) if len(sys.argv) > 1 else 7
dist = sys.argv[2] if len(sys.argv) > 2 else "nuclei"
n = int(sys.argv[3]) if len(sys.argv) > 3 else 4
B, C = 16, 256
feats = nb.stage_levels([f.cuda() for f in synth.fpn_levels(B, C)])
rois = synth.proposals(B, 1000 if P == 7 else 500, dist).cuda()
o = torch.empty(rois.shape[0], C, P, P, device="cuda")
for _ in range(n):
    nb.roi_align_levels(feats, rois, P, [1 / s for s in synth.FPN_STRIDES], 0, mode="route", out=o)
torch.cuda.synchronize()
