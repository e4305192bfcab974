#!/usr/bin/env python
"""Source-line profile of one kernel: joins the per-SASS-instruction counters of an ncu report with
the line table (and inline chains) nvdisasm prints for the same build of libvag_b200.so.
usage: python scripts/line_profile.py report.ncu-rep kernel_regex [lib.so] [top_n] [--chain]
Aggregates executed warp-instructions and stall samples by innermost source line (default) or, with
--chain, by the outermost non-kernel frame (which top-level call the time belongs to)."""
import csv, collections, io, os, re, subprocess, sys, tempfile

args = [a for a in sys.argv[1:] if not a.startswith("--")]
chain_mode = "--chain" in sys.argv
in_file = None  # --file=NAME: attribute every instruction to its innermost frame inside that source file
for a in sys.argv[1:]:
    if a.startswith("--file="):
        in_file = a.split("=", 1)[1]
rep, kern = args[0], args[1]
lib = args[2] if len(args) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vegasafterglow_b200", "libvag_b200.so")
top = int(args[3]) if len(args) > 3 else 40

raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern, "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, sass, kname = None, [], None
for r in rows:
    if r and r[0] == "Kernel Name" and kname is None:
        kname = r[1]
    if r and r[0] == "Address":
        if hdr is not None:
            break
        hdr = r
        ia, ie, isrc, ismp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= max(ie, isrc, ismp):
        continue
    sass.append((int(r[ia], 16), r[isrc].strip(), int(r[ie] or 0), int(r[ismp] or 0)))
base = sass[0][0]

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
# locate the function whose demangled name matches: use the mangled-name fragment of the kernel regex
lines = dis.splitlines()
func_re = re.compile(r"^\s*\.section\s+\.text\.(\S*%s\S*)," % re.escape(kern.split("<")[0]))
start = None
cands = [i for i, l in enumerate(lines) if func_re.match(l)]
# choose the candidate whose instruction count matches the ncu listing
line_re = re.compile(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?')
inst_re = re.compile(r"^\s*/\*([0-9a-f]{4,})\*/\s+(.*?);")
best = None
for c in cands:
    offs, cur = {}, []
    pending = []
    for l in lines[c + 1:]:
        if l.startswith("//---------------------") and offs:
            break
        m = line_re.search(l)
        if m:
            pending.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = inst_re.match(l)
        if m:
            if pending:
                cur = pending
                pending = []
            offs[int(m.group(1), 16)] = cur
    if best is None or abs(len(offs) - len(sass)) < abs(len(best) - len(sass)):
        best = offs
offs = best
print(f"kernel {kname}: {len(sass)} SASS instructions in the report, {len(offs)} in {os.path.basename(lib)}")
agg, smp = collections.Counter(), collections.Counter()
tot = tots = 0
for addr, src, n, s in sass:
    chain = offs.get(addr - base, [])
    if not chain:
        key = "?"
    elif in_file:
        hit = [f for f in chain if f[0] == in_file]
        key = "%s:%d" % hit[0] if hit else "(outside %s)" % in_file
    elif chain_mode:
        # frames listed innermost first; pick the outermost frame below the kernel body
        key = "%s:%d" % chain[-2] if len(chain) >= 2 else "%s:%d" % chain[-1]
    else:
        key = "%s:%d" % chain[0]
    agg[key] += n
    smp[key] += s
    tot += n
    tots += s
print(f"total executed warp-instructions {tot}, stall samples {tots}")
for key, n in sorted(agg.items(), key=lambda kv: -smp[kv[0]])[:top]:
    print(f"{key:34s} inst {n:12d} {100 * n / tot:5.1f}%   samples {smp[key]:8d} {100 * smp[key] / max(tots, 1):5.1f}%")
