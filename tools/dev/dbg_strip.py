import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import nuhtc_b200 as nb
from nuhtc_b200 import synth
B, C = 16, 256
g = torch.Generator().manual_seed(0)
a = torch.randn(B, C, 128, 128, generator=g).cuda()
for P in (7, 14):
  for dist in ("nuclei", "routed"):
    rois = synth.proposals(B, 1000 if P == 7 else 500, dist).cuda()
    lit = nb.roi_align_levels([a], rois, P, [0.25], 0, impl="direct")
    for trial in range(3):
        fast = nb.roi_align(a, rois, P, 0.25, 0)
        d = (fast - lit).abs().amax(dim=(1, 2, 3))
        bad = (d > 1e-5).nonzero().squeeze(1)
        print(P, dist, trial, "max", d.max().item(), "bad rois", bad.numel(), bad[:8].tolist())
        if bad.numel():
            k = int(bad[0])
            dd = (fast[k] - lit[k]).abs()
            print("   roi", rois[k].tolist(), "bad channels", (dd.amax(dim=(1, 2)) > 1e-5).nonzero().squeeze(1)[:10].tolist(),
                  "bad bins", (dd.amax(dim=0) > 1e-5).nonzero()[:10].tolist(), "max", dd.max().item())
