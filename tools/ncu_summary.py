#!/usr/bin/env python
"""Summarise an .ncu-rep (captured on the B200 box with `ncu --set full --import-source on`) into a text file for
profiles/: headline metrics per launch, SASS opcode mix and warp-stall reasons of the first kernel.
usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/NAME.txt"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "lts__t_bytes.sum"]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main(path):
    rows = list(csv.reader(ncu(["-i", path, "--page", "raw", "--csv"]).splitlines()))
    hdr, units = rows[0], rows[1]
    print("# ncu summary of", path)
    for r in rows[2:]:
        print("\n## launch:", r[hdr.index("Kernel Name")][:90])
        for k in KEYS:
            if k in hdr:
                print("%-62s %16s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
        try:
            rd = float(r[hdr.index("dram__bytes_read.sum")]); wr = float(r[hdr.index("dram__bytes_write.sum")])
            u = units[hdr.index("dram__bytes_read.sum")]
            print("%-62s %16.3f %s  (read+write = roofline.traffic)" % ("dram traffic per launch", rd + wr, u))
        except Exception:
            pass
    rows = list(csv.reader(ncu(["-i", path, "--page", "source", "--csv", "--print-source", "sass"]).splitlines()))
    hdr = None; kern = 0
    ops = collections.Counter(); samples = collections.Counter(); stalls = collections.Counter(); hot = []
    for r in rows:
        if r and r[0] == "Kernel Name":
            kern += 1; continue
        if r and r[0] == "Address":
            hdr = r; continue
        if kern != 1 or hdr is None or len(r) < len(hdr) - 5:
            continue
        src = r[1].strip()
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
        op = ".".join((m.group(2) if m else src).split(".")[:2])
        ie = int(r[hdr.index("Instructions Executed")] or 0); ss = int(r[hdr.index("# Samples")] or 0)
        ops[op] += ie; samples[op] += ss
        hot.append((ss, src[:80]))
        for i, h in enumerate(hdr):
            if h.startswith("stall_") and "Not Issued" not in h:
                try:
                    stalls[h] += int(r[i] or 0)
                except ValueError:
                    pass
    tot = sum(ops.values()) or 1; ts = sum(samples.values()) or 1
    print("\n## SASS opcode mix of the first launch (warp instructions executed: %d; stall samples: %d)" % (tot, ts))
    for op, c in ops.most_common(28):
        print("%-22s %12d %6.1f %%   stall samples %5.1f %%" % (op, c, 100.0 * c / tot, 100.0 * samples[op] / ts))
    print("\n## warp stall reasons (all samples)")
    s = sum(stalls.values()) or 1
    for k, v in stalls.most_common(10):
        print("%-28s %8d %6.1f %%" % (k, v, 100.0 * v / s))
    print("\n## hottest instructions by stall samples")
    hot.sort(key=lambda t: -t[0])
    for ss, src in hot[:20]:
        print("%6d  %s" % (ss, src))


if __name__ == "__main__":
    main(sys.argv[1])
