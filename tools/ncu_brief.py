#!/usr/bin/env python
"""Brief of one kernel of an .ncu-rep (read here, no GPU): duration, DRAM / L2 / shared traffic, issue and pipe use, stall
reasons, opcode mix and the hottest SASS lines.   python tools/ncu_brief.py gpurun_out/x.ncu-rep [kernel-index]"""
import csv
import io
import subprocess
import sys
from collections import Counter


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = page(rep, "raw")
    hdr, vals = rows[0], rows[2 + idx]
    get = {h: v for h, v in zip(hdr, vals)}
    print("kernel:", get.get("Kernel Name", "?")[:100])
    for k in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
              "lts__t_sectors_srcunit_tex_op_write.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
              "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active",
              "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
              "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
              "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
              "launch__registers_per_thread", "sm__cycles_elapsed.avg"):
        if k in get:
            print(f"  {k:75s} {get[k]}")
    st = {}
    for h, v in get.items():
        if "smsp__pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
            try:
                st[h.replace("smsp__pcsamp_warps_issue_stalled_", "")] = float(v.replace(",", ""))
            except ValueError:
                pass
    tot = sum(st.values()) or 1
    print("  stalls:", ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in sorted(st.items(), key=lambda x: -x[1])[:9]))
    src = page(rep, "source")
    if len(src) > 2 and "Instructions Executed" in src[1]:
        h = src[1]
        ia, isamp, isrc = h.index("Instructions Executed"), h.index("# Samples"), h.index("Source")
        data = [r for r in src[2:] if len(r) > max(ia, isamp, isrc) and r[ia].strip().isdigit() and r[isamp].strip().isdigit()]   # a report with several kernels repeats the header rows
        ti = sum(int(r[ia]) for r in data) or 1
        ts = sum(int(r[isamp]) for r in data) or 1
        c, cs = Counter(), Counter()
        for r in data:
            t = r[isrc].split()
            op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
            c[op] += int(r[ia])
            cs[op] += int(r[isamp])
        print(f"  SASS lines {len(data)}, warp instructions {ti}, samples {ts}")
        print("  opcodes:", ", ".join(f"{op} {100 * n / ti:.1f}%/{100 * cs[op] / ts:.1f}%s" for op, n in c.most_common(12)))
        top = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:12]
        for i in sorted(top):
            print(f"   {i:5d} inst {int(data[i][ia]):10d} samp {int(data[i][isamp]):6d}  {data[i][isrc].strip()[:80]}")


if __name__ == "__main__":
    main()
