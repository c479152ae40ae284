"""CPU: host-side logic of the package (no CUDA calls)."""
import os

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
        assert rle_encode(m, compressed=False)["counts"] == cnts[:k].tolist()
    assert rle_encode(np.zeros((4, 4), np.uint8), compressed=False)["counts"] == [16]
    assert rle_encode(np.ones((4, 4), np.uint8), compressed=False)["counts"] == [0, 16]


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


def test_tile_features_wire_format_and_sidecar(tmp_path):
    """QuPath Feature dicts of tools/infer_wsi.py:541-585 and the binary sidecar round trip (host logic, no GPU)."""
    import json
    from nuhtc_b200.contours import read_sidecar, tile_features, write_sidecar
    ring_xy = np.array([[10, 20], [10, 24], [15, 24], [15, 20], [10, 20], [100, 7], [103, 9], [101, 12], [100, 7]], dtype=np.float64)
    voff = np.array([0, 5, 9], dtype=np.int64)
    boxes = np.array([[10, 20, 16, 25], [100, 7, 104, 13]], dtype=np.float64)
    labels, scores = np.array([2, 0]), np.array([0.91, 0.42], dtype=np.float32)
    classes = ("a", "b", "c")
    colors = ([255, 0, 0], [0, 255, 0], [0, 0, 255])
    geo, pts = tile_features(ring_xy, voff, boxes, labels, scores, classes, colors)
    json.dumps(geo), json.dumps(pts)                                      # plain python types only
    assert geo[0]["geometry"] == {"type": "Polygon", "coordinates": [[[10, 20], [10, 24], [15, 24], [15, 20], [10, 20]]]}
    assert geo[0]["properties"] == {"objectType": "annotation", "label": 2, "score": float(np.float32(0.91)),
                                    "classification": {"name": "c", "color": [0, 0, 255]}, "isLocked": False}
    assert pts[1]["geometry"] == {"type": "Point", "coordinates": [102.0, 10.0]}
    # the flat feature list is what nuclei_merge reads back
    from nuhtc_b200.nuclei_merge import features_to_arrays
    xy2, voff2, sc2 = features_to_arrays(geo)
    assert np.array_equal(xy2, ring_xy) and np.array_equal(voff2, voff) and np.allclose(sc2, scores)
    path = str(tmp_path / "nuclei.npz")
    write_sidecar(path, ring_xy, voff, scores, labels, boxes)
    xy3, voff3, sc3, lab3, bb3 = read_sidecar(path)
    assert np.array_equal(xy3, ring_xy) and np.array_equal(voff3, voff) and np.array_equal(lab3, labels)
    assert np.allclose(sc3, scores) and np.allclose(bb3, boxes)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs first) prints ONE JSON line with the contract keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-tiles", "1", "--proposals", "60"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "wsi_tiles_per_sec_roi_stage_plus_merge" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_compressed_rle_string_round_trip():
    """The `counts` byte string of pycocotools (maskApi.c rleToString / rleFrString): 6-bit groups as ASCII 48..111, counts from
    the fourth on stored as differences to the count two places back (negative differences included)."""
    from nuhtc_b200.mask_nms import rle_counts_to_string, rle_decode, rle_string_to_counts
    assert rle_counts_to_string([5]) == b"5"                       # 5 -> one group, no continuation
    assert rle_counts_to_string([31]) == b"o0"                     # 31 = 0b11111: bit 4 set -> continuation flag + an empty group
    assert rle_counts_to_string([0, 16]) == b"0`0"                 # 16 -> group 16|32 then 0
    for counts in ([10, 3, 200, 1, 7, 100000, 2], [0, 1, 1, 1, 1, 1], [7, 300, 2, 5, 1000, 1, 3, 900]):
        s = rle_counts_to_string(counts)
        assert all(48 <= c <= 111 for c in s)
        assert rle_string_to_counts(s) == counts
    g = np.random.default_rng(0)
    for t in range(100):
        h, w = int(g.integers(1, 70)), int(g.integers(1, 70))
        m = (g.random((h, w)) < g.random()).astype(np.uint8)
        r = rle_encode(m)
        assert isinstance(r["counts"], bytes) and r["size"] == [h, w]
        assert np.array_equal(rle_decode(r), m)
