"""Stall samples of one kernel of an ncu report, aggregated per source line (read here, no GPU).

    python tools/ncu_lines.py gpurun_out/prof_x.ncu-rep k_decode_fused decode [top]

The SASS page of the report carries the samples per instruction; the line table comes from the cubin of the
in-tree library (`nvdisasm -g`), so the library must be the build the report was taken from.
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, kern, unit = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "mjpeg423-video-decoder-software_b200", "libmjpeg423_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", unit, lib], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.startswith(unit + ".")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=tmp, capture_output=True, text=True).stdout
fn, line, amap = None, None, {}
for l in sass.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        fn = m.group(1)
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        line = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m and fn and kern in fn:
        amap[int(m.group(1), 16)] = line
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
hdr = rows[h]
A, S, IE, SRC = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
cols = ["stall_long_sb", "stall_wait", "stall_short_sb", "stall_math", "stall_not_selected", "stall_selected",
        "stall_branch_resolving", "stall_no_inst", "stall_mio", "stall_lg", "stall_dispatch"]
ci = [hdr.index(c) for c in cols]
agg = collections.defaultdict(lambda: [0, 0] + [0] * len(cols))
base, tot, totie, recs = None, 0, 0, []
for r in rows[h + 1:]:
    if len(r) <= S or not r[A]:
        continue
    if r[A] == "Address":          # the regex matched another launch: keep the first one
        break
    a = int(r[A], 16) if r[A].startswith("0x") else int(r[A])
    base = a if base is None else base
    s, ie = int(r[S] or 0), int(r[IE] or 0)
    g = agg[amap.get(a - base)]
    g[0] += s
    g[1] += ie
    for k, i in enumerate(ci):
        g[2 + k] += int(r[i] or 0)
    tot += s
    totie += ie
    recs.append((a - base, s, ie, amap.get(a - base), r[SRC]))
print("samples", tot, "warp instructions", totie)
tc = [sum(g[2 + k] for g in agg.values()) for k in range(len(cols))]
print("stall totals %:", {c: round(100 * v / tot, 1) for c, v in zip(cols, tc)})
print("line | samples % | instructions % |", cols)
for k, g in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print(k, round(100 * g[0] / tot, 1), round(100 * g[1] / totie, 1), g[2:])
print("--- hottest instructions")
for i in sorted(range(len(recs)), key=lambda i: -recs[i][1])[:12]:
    off, s, ie, ln, src = recs[i]
    print(hex(off), round(100 * s / tot, 1), ie, ln, src[:90])
