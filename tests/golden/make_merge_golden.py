"""Golden kept-id hashes of the cross-tile merge at BASELINE cfg-5 sizes ("merge kept IDs for 10k/100k/1M").

Made with the CPU oracle (oracle.cpu.merge_overlap_arrays, the restatement of /root/reference/tools/nuclei_merge.py:62-174)
on the seeded synthetic slides of nuhtc_b200.synth.slide_nuclei.  The oracle itself is cross-checked at these sizes by
tests/test_oracle_cpu.py (slab-decomposition area, O(N^2) candidate search at 10k).  The input arrays are hashed too, so a
drift of the generator is told apart from a wrong merge.

Integer vertices make IoU a rational that can equal the threshold 1/20 EXACTLY (88 decisive pairs among 1.7M nuclei), and
what GEOS's double arithmetic decides there is not knowable without GEOS: |IoU - thr| < 1e-9 is outside the parity contract
(SURVEY.md H3).  The lower-scored nucleus of each such pair is therefore dropped (oracle.cpu.drop_threshold_ties); the dropped
original indices are stored so the GPU test rebuilds the identical input without re-deriving them.  Likewise the
lower-scored copy of two nuclei with IDENTICAL rings is dropped: the reference keys a dict by the shapely polygon
(nuclei_merge.py:101-103), the copies collide and it keeps the lower-scored one -- excluded from the contract like score ties.

    python tests/golden/make_merge_golden.py        # rewrites tests/golden/merge_large.json
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import numpy as np  # noqa: E402

CASES = [dict(tiles=(20, 20), seed=20), dict(tiles=(66, 66), seed=66), dict(tiles=(208, 208), seed=208)]


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def remove_nuclei(d, removed):
    """the slide without the nuclei `removed` (original indices)"""
    alive = np.ones(len(d["score"]), dtype=bool)
    alive[np.asarray(removed, dtype=np.int64)] = False
    idx = np.nonzero(alive)[0]
    cnt = np.diff(d["voff"])[idx]
    nv = np.zeros(len(idx) + 1, dtype=np.int64)
    nv[1:] = np.cumsum(cnt)
    sel = np.repeat(d["voff"][:-1][idx], cnt) + (np.arange(nv[-1]) - np.repeat(nv[:-1], cnt))
    out = dict(d)
    out.update(xy=d["xy"][sel], voff=nv, score=d["score"][idx], tile_id=d["tile_id"][idx])
    return out


def case_inputs(c, removed=None):
    from nuhtc_b200 import synth
    d = synth.slide_nuclei(c["tiles"][0], c["tiles"][1], per_tile=23, seed=c["seed"])
    return d if removed is None else remove_nuclei(d, removed)


def main():
    from oracle import cpu as O
    out = []
    for c in CASES:
        raw = case_inputs(c)
        d, removed = O.drop_threshold_ties(raw, 0.05)
        chk = remove_nuclei(raw, removed)
        assert all(np.array_equal(d[k], chk[k]) for k in ("xy", "voff", "score"))
        row = dict(tiles=list(c["tiles"]), seed=c["seed"], nuclei=int(len(d["score"])), vertices=int(len(d["xy"])),
                   removed_out_of_contract=[int(i) for i in removed], inputs_sha256=sha(d["xy"], d["voff"], d["score"]))
        for strat in ("probability", "area"):
            k, margin, ties = O.merge_overlap_arrays_check(d["xy"], d["voff"], d["score"], 0.05, strat, naive=False, with_margin=True)
            assert len(ties) == 0 and margin > 1e-9
            row[strat] = dict(kept=int(len(k)), kept_ids_sha256=sha(k.astype(np.int64)), min_decisive_margin=margin)
        out.append(row)
        print(row)
    json.dump(dict(overlap_threshold=0.05, per_tile=23, cases=out), open(os.path.join(HERE, "merge_large.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
