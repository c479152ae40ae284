#!/usr/bin/env python
"""How much of the L2->SM window traffic of the 7x7 RoIAlign could be shared between RoIs?  (CPU analysis, no GPU.)

For a proposal distribution, every RoI's level-0 window (the cells its bilinear taps touch) is computed like the kernel
does, the RoIs of a tile are grouped greedily under a shared box of BOX x BOX cells, and the cells of the groups' union
windows are compared with the sum of the individual windows (what the RoI-stationary kernel reads today).  DESIGN.md quotes
the 'nuclei' numbers; 'clustered' models proposals crowding around nuclei."""
import sys

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from nuhtc_b200 import synth  # noqa: E402


def main():
    B, n_per, scale, cells = 16, 1000, 0.25, 128
    for dist in ("nuclei", "clustered"):
        rois = synth.proposals(B, n_per, dist, frame=512, seed=0).numpy()
        b, x1, y1, x2, y2 = rois.T
        X0 = np.floor(np.maximum(x1 * scale - 0.5, 0)).astype(int)
        X1 = np.minimum(np.floor(x2 * scale - 0.5).astype(int) + 1, cells - 1)
        Y0 = np.floor(np.maximum(y1 * scale - 0.5, 0)).astype(int)
        Y1 = np.minimum(np.floor(y2 * scale - 0.5).astype(int) + 1, cells - 1)
        area = (X1 - X0 + 1) * (Y1 - Y0 + 1)
        print(f"{dist}: mean window {area.mean():.0f} cells ({area.mean() * 1.024:.0f} KB at C=256)")
        for box in (12, 16, 20, 24):
            tot_union, groups = 0, 0
            for t in range(B):
                idx = np.nonzero(b == t)[0]
                order = idx[np.lexsort((X0[idx], Y0[idx] // 6))]
                xs0, ys0, xs1, ys1 = X0[order], Y0[order], X1[order], Y1[order]
                used = np.zeros(len(order), bool)
                for i in range(len(order)):
                    if used[i]:
                        continue
                    fit = (~used) & (xs0 >= xs0[i]) & (ys0 >= ys0[i]) & (xs1 < xs0[i] + box) & (ys1 < ys0[i] + box)
                    fit[i] = True
                    used |= fit
                    groups += 1
                    tot_union += (xs1[fit].max() - xs0[fit].min() + 1) * (ys1[fit].max() - ys0[fit].min() + 1)
            print(f"  shared box {box:2d}x{box:2d} cells ({box * box * 1.024:4.0f} KB): {len(area) / groups:5.2f} RoIs per group, "
                  f"union / sum of windows = {tot_union / area.sum():.2f}")


if __name__ == "__main__":
    main()
