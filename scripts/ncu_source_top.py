"""Aggregates `ncu --page source --csv` (SASS view): stall samples by reason over the whole kernel, and the address ranges
(loops found from backward branches) that collect the most samples.  usage: python scripts/ncu_source_top.py source.csv"""
import csv, sys, re
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix["# Samples"]]) for r in data)
print("samples", tot)
agg = {h: sum(int(r[ix[h]]) for r in data) for h in stall}
for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]:
    print(f"  {h:24s} {v:8d} {100 * v / tot:5.1f} %")
base = int(data[0][ix["Address"]], 16)
addr = [int(r[ix["Address"]], 16) - base for r in data]
samp = [int(r[ix["# Samples"]]) for r in data]
src = [r[ix["Source"]].strip() for r in data]
# windows of 64 instructions
W = 64
win = [(sum(samp[i:i + W]), i) for i in range(0, len(data), W)]
print("hottest 64-instruction windows:")
for s, i in sorted(win, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 16]:
    r = data[i:i + W]
    top = sorted(((sum(int(x[ix[h]]) for x in r), h) for h in stall), reverse=True)[:4]
    ops = {}
    for x in r:
        op = x[ix["Source"]].split()[0] if x[ix["Source"]].split() else "?"
        if op.startswith("@"): op = x[ix["Source"]].split()[1]
        ops[op.split(".")[0]] = ops.get(op.split(".")[0], 0) + 1
    print(f"  {addr[i]:#08x}: {s:7d} ({100 * s / tot:4.1f} %)  " + ", ".join(f"{h[6:]} {v}" for v, h in top) + "   " + " ".join(f"{k}:{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:5]))
