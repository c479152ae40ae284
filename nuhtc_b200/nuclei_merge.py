"""Cross-tile nucleus merge: drop-in for /root/reference/tools/nuclei_merge.py.

``merge_overlap(cells, overlap_threshold, merge_strategy, uniform_classification)`` keeps the reference's
signature and return value (a DataFrame of the kept rows in score order, whose row index is the
``nuclei_id`` written at nuclei_merge.py:201); the work runs in csrc/merge.cu.  ``main()`` keeps the CLI
flags of nuclei_merge.py:221-230.
"""
from __future__ import annotations

import json
import os
import time
from argparse import ArgumentParser
from typing import Sequence, Tuple

import numpy as np
import torch

from . import _lib as L

__all__ = ["merge_arrays", "merge_overlap", "features_to_arrays", "main"]

_STRATEGY = {"probability": 0, "area": 1}


def merge_arrays(xy: torch.Tensor, voff: torch.Tensor, score: torch.Tensor, overlap_threshold: float = 0.01,
                 merge_strategy: str = "probability", max_pairs: int = 0) -> torch.Tensor:
    """xy [sumV,2] fp64, voff [N+1] int64, score [N] fp64 (CUDA) -> kept ORIGINAL indices (int64, CUDA) ordered by
    score rank: element r is the nucleus that receives nuclei_id r."""
    if merge_strategy not in _STRATEGY:
        raise ValueError(f"Invalid merge strategy: {merge_strategy}. Use 'probability' or 'area'.")
    L.require_cuda(xy, "xy")
    dev = xy.device
    xy = xy.to(torch.float64).contiguous()
    voff = voff.to(dev, torch.int64).contiguous()
    score = score.to(dev, torch.float64).contiguous()
    N = score.numel()
    if N == 0:
        return torch.empty(0, dtype=torch.int64, device=dev)
    sumV = xy.shape[0]
    lib = L.lib()
    keep = torch.empty(N, dtype=torch.int64, device=dev)
    nkeep = torch.zeros(1, dtype=torch.int64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    cap = int(max_pairs) if max_pairs > 0 else 8 * N + 1024
    for _ in range(8):
        wsb = lib.nuhtc_merge_workspace_bytes(N, sumV, cap)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = lib.nuhtc_merge(xy.data_ptr(), voff.data_ptr(), score.data_ptr(), N, sumV, float(overlap_threshold),
                                 _STRATEGY[merge_strategy], cap, keep.data_ptr(), nkeep.data_ptr(), status.data_ptr(),
                                 ws.data_ptr(), wsb, L.stream_ptr(dev))
        if rc == -4:  # NUHTC_EOVERFLOW: num_keep carries the pair count to retry with
            cap = int(nkeep.item()) + 1024
            del ws
            continue
        L.check(rc, "merge")
        L.count("merge")
        break
    else:
        raise L.NuhtcError("merge: candidate-pair capacity could not be satisfied")
    return keep[: int(nkeep.item())]


def features_to_arrays(features: Sequence[dict]) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Flat GeoJSON Feature list (the format infer_wsi.py writes, :550-566) -> (xy [sumV,2], voff [N+1], score [N]).
    The polygon is the exterior ring ``geometry.coordinates[0]`` (nuclei_merge.py:40); score is
    ``properties.get('score', 0)`` (:80)."""
    rings = [np.asarray(f["geometry"]["coordinates"][0], dtype=np.float64).reshape(-1, 2) for f in features]
    voff = np.zeros(len(rings) + 1, dtype=np.int64)
    if rings:
        voff[1:] = np.cumsum([r.shape[0] for r in rings])
        xy = np.concatenate(rings, axis=0)
    else:
        xy = np.zeros((0, 2), dtype=np.float64)
    score = np.asarray([f.get("properties", {}).get("score", 0) for f in features], dtype=np.float64)
    return xy, voff, score


def merge_overlap(cleaned_edge_cells, overlap_threshold=0.01, merge_strategy="probability", uniform_classification=False,
                  device="cuda"):
    """Reference signature.  ``cleaned_edge_cells``: list of GeoJSON Feature dicts (or a DataFrame with
    'geometry' and 'properties' columns).  Returns a pandas DataFrame of the kept rows, score-descending, index
    0..k-1 (+ a 'score' column), like the reference's ``merged_cells.sort_index()``."""
    import pandas as pd

    start_time = time.time()
    frame = pd.DataFrame(cleaned_edge_cells)
    feats = frame.to_dict("records")
    xy, voff, score = features_to_arrays(feats)
    frame["score"] = score if len(frame) else []
    keep = merge_arrays(torch.from_numpy(xy).to(device), torch.from_numpy(voff).to(device), torch.from_numpy(score).to(device),
                        overlap_threshold, merge_strategy)
    keep = keep.cpu().numpy()
    out = frame.iloc[keep].reset_index(drop=True).copy()
    print(f"Cell overlap removal elapsed time: {time.time() - start_time:.4f} seconds")
    return out


def parse_args(argv=None):
    parser = ArgumentParser()
    parser.add_argument("--geojson", help="geojson file name")
    parser.add_argument("--output_name", default=None, type=str, help="output geojson file name")
    parser.add_argument("--overlap_threshold", type=float, default=0.01, help="area overlap percentage threshold to be removed")
    parser.add_argument("--merge_strategy", default="probability",
                        help="whether to keep the cell with highest probability or largest area, specify 'probability' or 'area'")
    parser.add_argument("--uniform_classification", action="store_true",
                        help="whether to classify all cells uniformly and represent as the same color (yellow)")
    return parser.parse_args(argv)


def main(argv=None):
    args = parse_args(argv)
    datadir = os.path.dirname(args.geojson)
    name = os.path.basename(args.geojson).split(".geojson")[0]
    with open(os.path.join(datadir, f"{name}.geojson"), "r") as f:
        data = json.load(f)
    merged = merge_overlap(data, args.overlap_threshold, args.merge_strategy, args.uniform_classification)
    out_path = os.path.join(datadir, f"{args.output_name}.geojson" if args.output_name else f"{name}_merged.geojson")
    features = []
    for idx, row in merged.iterrows():
        props = row["properties"]
        props["nuclei_id"] = idx
        if args.uniform_classification:
            props["classification"]["name"] = "uniform"
            props["classification"]["color"] = [255, 255, 0]
        features.append({"type": "Feature", "geometry": row["geometry"], "properties": props})
    with open(out_path, "w") as f:
        json.dump(features, f)
    print("Merge and save complete.")


if __name__ == "__main__":
    main()
