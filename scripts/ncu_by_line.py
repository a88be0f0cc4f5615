#!/usr/bin/env python
"""Join an ncu source-page CSV (SASS level, `ncu -i rep --page source --csv`) with nvdisasm -gi line info of the cubin,
and print stall samples / executed instructions aggregated per source line and per file.
usage: ncu_by_line.py <src.csv> <nvdisasm -gi output> <kernel substring> [top]"""
import collections
import csv
import re
import sys

src_csv, sass, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# --- address -> (file, line) from nvdisasm
addr2line = {}
infn = False
cur = ("?", 0)
stack = []
for ln in open(sass, errors="replace"):
    if ln.startswith("//---") and ".text." in ln:
        infn = kern in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        stack.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
    if m:
        if stack:
            # innermost frame first, outermost (kernel body) last: key = innermost, tagged with the outermost line (= role)
            cur = (f"{stack[0][0]}", stack[0][1], stack[-1][1])
            stack = []
        addr2line[int(m.group(1), 16)] = (cur, m.group(2))
rows = list(csv.reader(open(src_csv)))
h = rows[1]
ai, si, ii = h.index("Address"), h.index("# Samples"), h.index("Instructions Executed")
stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
base = None
per_line = collections.defaultdict(lambda: [0, 0, collections.Counter()])
per_file = collections.defaultdict(lambda: [0, 0])
tot_s = tot_i = 0
for r in rows[2:]:
    if len(r) <= max(ai, si, ii):
        continue
    a = int(r[ai], 16) if r[ai].startswith("0x") else int(r[ai])
    if base is None:
        base = a
    off = a - base
    key, ins = addr2line.get(off, (("?", 0, 0), "?"))
    s = int(r[si] or 0); n = int(r[ii] or 0)
    per_line[key][0] += s; per_line[key][1] += n
    for c in stall_cols:
        v = int(r[c] or 0)
        if v:
            per_line[key][2][h[c]] += v
    per_file[f"outer-line {key[2]}"][0] += s; per_file[f"outer-line {key[2]}"][1] += n
    tot_s += s; tot_i += n
print(f"total samples {tot_s}, warp-instructions {tot_i}")
for f, (s, n) in sorted(per_file.items(), key=lambda kv: -kv[1][0]):
    print(f"  {f:28s} samples {100*s/tot_s:5.1f}%  instr {100*n/max(1,tot_i):5.1f}%")
print("top lines by samples:")
for key, (s, n, st) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
    tops = ", ".join(f"{k[6:]}={v}" for k, v in st.most_common(3))
    print(f"  {key[0]:20s}:{key[1]:4d} @{key[2]:3d}  samples {100*s/tot_s:5.1f}%  instr {100*n/max(1,tot_i):5.1f}%   {tops}")
