#!/usr/bin/env python
"""Join an `ncu --page source --csv` SASS export with nvdisasm -g line info and print the
source lines that collect the most warp-stall samples / executed instructions.

    tools/ncu_lines.py <ncu sass csv> <nvdisasm -g -c output> <kernel name substring> [top]
"""
import csv
import re
import sys
from collections import defaultdict

csvf, sassf, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# nvdisasm: per function, instruction offset -> line
lines = {}
cur = None
infn = False
for ln in open(sassf):
    if ln.startswith(".text.") or ".section\t.text." in ln or re.match(r"\s*\.section\s+\.text\.", ln):
        infn = kname in ln
    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m and infn:
        lines[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(csvf)))
hdr = rows[1]
ia, isamp, iinst = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
base = None
agg = defaultdict(lambda: [0, 0, defaultdict(int)])
for r in rows[2:]:
    if len(r) < len(hdr):
        r = r + ["0"] * (len(hdr) - len(r))
    a = int(r[ia], 16)
    if base is None:
        base = a
    key = lines.get(a - base, ("?", 0))
    agg[key][0] += int(r[isamp] or 0)
    agg[key][1] += int(r[iinst] or 0)
    for i in stall_cols:
        v = int(r[i] or 0)
        if v:
            agg[key][2][hdr[i][6:]] += v
tot = sum(v[0] for v in agg.values())
toti = sum(v[1] for v in agg.values())
print("total samples %d, warp instructions %d" % (tot, toti))
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = sorted(v[2].items(), key=lambda kv: -kv[1])[:3]
    print("%-16s %5d  %5.1f%% samples  %5.1f%% inst  %s" % (key[0], key[1], 100.0 * v[0] / tot, 100.0 * v[1] / toti,
                                                         " ".join("%s=%d" % s for s in st)))
