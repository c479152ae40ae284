import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import nuhtc_b200 as nb
from nuhtc_b200 import synth
P = int(sys.argv[1]); nstreams = int(sys.argv[2]); reps = int(sys.argv[3])
B, C = 16, 256
feats = [f.cuda() for f in synth.fpn_levels(B, C)]
st = nb.stage_levels(feats)
rois = [synth.proposals(B, 1000 if P == 7 else 500, "nuclei", seed=s).cuda() for s in range(nstreams)]
scales = [1 / s for s in synth.FPN_STRIDES]
streams = [torch.cuda.Stream() for _ in range(nstreams)]
outs = [torch.empty(r.shape[0], C, P, P, device="cuda") for r in rois]
torch.cuda.synchronize()
t = time.time()
for it in range(reps):
    for i, s in enumerate(streams):
        with torch.cuda.stream(s):
            nb.roi_align_levels(st, rois[i], P, scales, 0, mode="route", out=outs[i])
    if it % 10 == 9:
        torch.cuda.synchronize()
        print("P", P, "streams", nstreams, "iter", it, "ok %.1f ms" % ((time.time() - t) * 1e3), flush=True)
torch.cuda.synchronize()
print("done", flush=True)
