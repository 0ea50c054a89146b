"""In-order issue model of one warp running a SASS loop: cycles per iteration when the warp is alone on its SM partition
(dependent-issue latencies: FP64 9, shared-memory load 30, everything else 5; the FP64 pipe takes one warp instruction per
2 cycles).  With W warps per partition the loop cannot run faster than max(alone / W, 2 * #FP64) cycles per iteration and
warp.  usage: python scripts/sass_sim.py loop.sass [iterations]"""
import re, sys
LAT = {"D": 9, "LDS": 30}
def regs(tok):
    m = re.match(r'-?\|?~?(U?R)(\d+)', tok)
    return (m.group(1), int(m.group(2))) if m and not tok.startswith(("RZ", "URZ")) else None
ins = []
for ln in open(sys.argv[1]):
    m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?(\S+)\s*(.*?);', ln)
    if not m: continue
    op, rest = m.group(1), m.group(2)
    toks = [t.strip() for t in re.split(r',\s*', rest)] if rest else []
    toks = [re.sub(r'\.reuse|\.H[01]_H[01]', '', t) for t in toks]
    wide = 2 if op[0] == "D" and op[:4] in ("DMUL", "DFMA", "DADD", "DSET", "DMNM") else 1
    dst, src = [], []
    if op.startswith(("ST", "BRA", "BAR", "RED", "ATOM", "EXIT", "WARPSYNC", "NOP")):
        srct = toks
    else:
        srct = toks[1:]
        r = regs(toks[0]) if toks else None
        if r:
            n = 4 if ".128" in op else 2 if (".64" in op or wide == 2) else 1
            dst = [(r[0], r[1] + i) for i in range(n)]
    for t in srct:
        for a in re.findall(r'U?R\d+', t):
            r = regs(a)
            if r:
                src += [(r[0], r[1] + i) for i in range(wide if not t.startswith("[") else 1)]
    kind = "D" if wide == 2 else "LDS" if op.startswith("LDS") else "X"
    ins.append((op, dst, src, kind))
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
ready, t, last_d, marks = {}, 0, -10, []
for it in range(iters):
    for op, dst, src, kind in ins:
        t = max([t + 1] + [ready.get(s, 0) for s in src])
        if kind == "D":
            t = max(t, last_d + 2); last_d = t
        for d in dst: ready[d] = t + LAT.get(kind, 5)
    marks.append(t)
nd = sum(k == "D" for *_, k in ins)
per = (marks[-1] - marks[1]) / (iters - 2)
print(f"{len(ins)} instructions, {nd} FP64, {sum(k == 'LDS' for *_, k in ins)} LDS: one warp alone {per:.0f} cycles per iteration; FP64 pipe floor {2 * nd}")
