"""Summarise an `ncu --page source --csv` export: opcode mix, stall samples, hottest SASS lines."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
ops, samples, tot = collections.Counter(), collections.Counter(), 0
data = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[0] == "Address":
        continue
    src = r[idx["Source"]]
    try:
        n = int(r[idx["Instructions Executed"]] or 0)
        s = int(r[idx["# Samples"]] or 0)
    except ValueError:
        continue
    toks = src.split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0]
    ops[op] += n
    samples[op] += s
    tot += n
    data.append((r[idx["Address"]], src, n, s))
print("total warp instructions", tot, " SASS lines", len(data))
stot = sum(samples.values())
for op, n in ops.most_common(22):
    print(f"{op:10s} {n:10d} {n / tot:6.1%}   stall samples {samples[op]:6d} {samples[op] / max(stot,1):6.1%}")
print("hottest lines by stall samples:")
for a, src, n, s in sorted(data, key=lambda d: -d[3])[:25]:
    print(f"  {a} n={n:8d} samples={s:5d}  {src[:100]}")
