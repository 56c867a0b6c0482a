#!/usr/bin/env python
"""Aggregate an ncu SASS source page (--page source --csv) per CUDA source line, using nvdisasm -g line annotations of the same
cubin (instructions match by order).  usage: ncu_by_line.py <ncu_sass.csv> <nvdisasm.txt> <kernel mangled name> [top]"""
import csv
import re
import sys
from collections import defaultdict


def main():
    sass_csv, dis, kern = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    rows = list(csv.reader(open(sass_csv)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]     # one block per profiled launch: keep the first
    rows = rows[starts[0]:(starts[1] if len(starts) > 1 else len(rows))]
    hdr = rows[1]
    ix, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    insts = [(r[1].strip(), int(r[ix] or 0), int(r[isamp] or 0)) for r in rows[2:] if len(r) > ix]
    lines, cur, inside = [], None, False
    for ln in open(dis):
        if ln.startswith(".text."):
            inside = (kern in ln)
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", ln):
            lines.append(cur)
    n = min(len(lines), len(insts))
    if abs(len(lines) - len(insts)) > 2:
        print("warning: %d disassembled vs %d profiled instructions" % (len(lines), len(insts)))
    agg, samp = defaultdict(int), defaultdict(int)
    for k in range(n):
        agg[lines[k]] += insts[k][1]
        samp[lines[k]] += insts[k][2]
    tot, tots = sum(agg.values()) or 1, sum(samp.values()) or 1
    print("total warp instructions %d, samples %d" % (tot, tots))
    for key, v in sorted(agg.items(), key=lambda x: -x[1])[:top]:
        print("%-22s %12d  %5.1f%% inst  %5.1f%% samples" % ("%s:%d" % key if key else "?", v, 100.0 * v / tot, 100.0 * samp[key] / tots))


if __name__ == "__main__":
    main()
