#!/usr/bin/env python
"""Hot spots of an ncu source page (SASS view): python scripts/ncu_hot.py rep.ncu-rep [kernel-regex] [top]
Groups the sampled stall reasons by SASS instruction and prints the top entries plus per-opcode totals."""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
# first line is the kernel name row
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
S = ix["# Samples"]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = 0
ops = collections.Counter()
opsn = collections.Counter()
recs = []
for r in rows[1:]:
    if len(r) < len(hdr):
        continue
    try:
        n = int(r[S])
    except ValueError:
        continue
    tot += n
    src = r[ix["Source"]]
    op = src.split()[0] if src.split() else "?"
    if op.startswith("@"):
        op = src.split()[1]
    ops[op] += n
    opsn[op] += int(r[ix["Instructions Executed"]] or 0)
    st = {c: int(r[ix[c]] or 0) for c in stall_cols}
    recs.append((n, r[ix["Address"]], src, st))
print("total samples", tot)
print("by opcode (samples, share, executed):")
for op, n in ops.most_common(25):
    print(f"  {op:24s} {n:8d} {100*n/tot:5.1f}%  exec={opsn[op]}")
print("top instructions:")
for n, a, src, st in sorted(recs, key=lambda x: -x[0])[:top]:
    s = ", ".join(f"{k[6:]}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v)
    print(f"  {n:6d} {100*n/tot:4.1f}% {a[-5:]} {src[:70]:70s} {s}")
