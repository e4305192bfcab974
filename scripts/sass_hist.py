#!/usr/bin/env python
"""Opcode histogram (weighted by executed warp instructions) of one kernel from an ncu report.
usage: python scripts/sass_hist.py report.ncu-rep kernel_regex [top_n]"""
import csv, collections, re, subprocess, sys, io

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern, "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None
ops, samp, tot = collections.Counter(), collections.Counter(), 0
nstatic = 0
done_first = False
for r in rows:
    if r and r[0] == "Address":
        if hdr is not None:
            break  # first matching launch only
        hdr = r
        ie, src, smp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= max(ie, src, smp):
        continue
    s = re.sub(r"^@!?U?P\d+\s+", "", r[src].strip())
    if not s:
        continue
    op = s.split()[0].split(".")[0]
    n = int(r[ie] or 0)
    ops[op] += n
    tot += n
    samp[op] += int(r[smp] or 0)
    nstatic += 1
print("executed warp-instructions", tot, "static SASS lines", nstatic)
for op, n in ops.most_common(top):
    print(f"{op:10s} {n:12d} {100 * n / tot:5.1f}%  stall-samples {samp[op]}")
