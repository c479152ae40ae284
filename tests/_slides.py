"""Golden merge cases (tests/golden/merge_large.json): the seeded synthetic slide minus the nuclei whose decisive IoU sits
exactly on the threshold (outside the parity contract; see tests/golden/make_merge_golden.py)."""
import hashlib
import json
import os

import numpy as np

G = os.path.join(os.path.dirname(__file__), "golden")


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def golden():
    return json.load(open(os.path.join(G, "merge_large.json")))


def load_case(i):
    """-> (case dict, slide dict).  Fails loudly if the generator no longer reproduces the hashed inputs."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_merge_golden", os.path.join(G, "make_merge_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    z = golden()
    c = z["cases"][i]
    d = mod.case_inputs(dict(tiles=c["tiles"], seed=c["seed"]), removed=c["removed_out_of_contract"])
    assert sha(d["xy"], d["voff"], d["score"]) == c["inputs_sha256"], "synthetic slide generator drifted"
    return c, d, z["overlap_threshold"]
