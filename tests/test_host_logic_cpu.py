"""CPU: host-side logic of the package (no CUDA calls)."""
import numpy as np
import torch

from nuhtc_b200 import synth
from nuhtc_b200.mask_nms import rle_encode
from nuhtc_b200.nuclei_merge import features_to_arrays, parse_args


def test_rle_encode_matches_oracle_runs(oracle):
    import ctypes
    rng = np.random.default_rng(0)
    for _ in range(10):
        m = (rng.random((17, 23)) > 0.6).astype(np.uint8)
        cnts = np.zeros(m.size + 1, dtype=np.uint32)
        k = oracle.lib().oracle_rle_encode(m.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), 17, 23,
                                           cnts.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)))
        assert rle_encode(m)["counts"] == cnts[:k].tolist()
    assert rle_encode(np.zeros((4, 4), np.uint8))["counts"] == [16]
    assert rle_encode(np.ones((4, 4), np.uint8))["counts"] == [0, 16]


def test_features_to_arrays_and_cli():
    feats = [{"geometry": {"type": "Polygon", "coordinates": [[[0, 0], [4, 0], [4, 4], [0, 0]]]}, "properties": {"score": 0.5}},
             {"geometry": {"type": "Polygon", "coordinates": [[[1, 1], [2, 1], [2, 2], [1, 2], [1, 1]]]}, "properties": {}}]
    xy, voff, score = features_to_arrays(feats)
    assert voff.tolist() == [0, 4, 9] and xy.shape == (9, 2) and score.tolist() == [0.5, 0.0]
    a = parse_args(["--geojson", "x.geojson", "--overlap_threshold", "0.05", "--merge_strategy", "area"])
    assert a.overlap_threshold == 0.05 and a.merge_strategy == "area" and not a.uniform_classification


def test_synth_shapes_and_levels(oracle):
    r = synth.proposals(2, 100, "nuclei")
    assert r.shape == (200, 5) and oracle.map_roi_levels(r, 4).max() == 0
    r = synth.proposals(4, 500, "routed")
    assert len(torch.unique(oracle.map_roi_levels(r, 4))) == 4
    b, s, l = synth.nms_boxes(2000)
    assert len(torch.unique(s)) == 2000
    d = synth.slide_nuclei(3, 2, per_tile=5)
    assert d["voff"][-1] == d["xy"].shape[0] and len(np.unique(d["score"])) == len(d["score"])
    assert (np.diff(d["voff"]) >= 3).all()
