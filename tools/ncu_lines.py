#!/usr/bin/env python
"""Instructions executed and stall samples of one kernel of an .ncu-rep aggregated by CUDA source line (needs -lineinfo and
--import-source on at capture time):   python tools/ncu_lines.py x.ncu-rep [top]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    cur, hdr, agg = None, None, []
    for r in csv.reader(out.splitlines()):
        if len(r) == 2 and r[0] in ("File Path", "File Name"):
            cur = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[0].isdigit():
            d = dict(zip(hdr, r))
            try:
                inst, samp = int(d["Instructions Executed"]), int(d["# Samples"])
            except (KeyError, ValueError):
                continue
            if inst or samp:
                agg.append((inst, samp, cur, int(r[0]), r[1].strip()[:100]))
    tot, ts = sum(a[0] for a in agg) or 1, sum(a[1] for a in agg) or 1
    print(f"warp instructions {tot}, samples {ts}")
    for a in sorted(agg, reverse=True)[:top]:
        print(f"{100 * a[0] / tot:5.1f}% inst {100 * a[1] / ts:5.1f}% samp  {a[2]}:{a[3]:<5d} {a[4]}")


if __name__ == "__main__":
    main()
