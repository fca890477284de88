#!/usr/bin/env python3
"""Summarise an `ncu --page source --csv` dump: stall reasons, hottest SASS lines, opcode mix.
usage: ncu -i X.ncu-rep --page source --csv > src.csv ; python profiles/ncu_source_summary.py src.csv"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if r and r[0] == "Address")
data = [r for r in rows if len(r) == len(hdr) and r[0].startswith("0x")]
ia, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[isamp]) for r in data)
totex = sum(int(r[iex]) for r in data)
print("samples", tot, "warp-instructions", totex, "sass lines", len(data))
st = {hdr[i]: sum(int(r[i]) for r in data) for i in stall_cols}
print("stalls:", [(k, v, round(100 * v / max(tot, 1), 1)) for k, v in sorted(st.items(), key=lambda x: -x[1])[:8]])
for r in sorted(data, key=lambda r: -int(r[isamp]))[: int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    s = {hdr[i]: int(r[i]) for i in stall_cols if int(r[i]) > 0}
    print(r[isamp].rjust(6), r[iex].rjust(8), r[ia][:64].ljust(64), sorted(s.items(), key=lambda x: -x[1])[:2])
c = Counter()
for r in data:
    parts = r[ia].split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    c[op.split(".")[0]] += int(r[iex])
print("opcode mix:", c.most_common(16))
