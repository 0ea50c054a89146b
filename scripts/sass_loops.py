"""Lists the loops (backward branches) of a kernel's SASS with their size: the hot loop has to fit the instruction cache.
usage: cuobjdump -sass -fun <mangled> obj.o > k.sass; python scripts/sass_loops.py k.sass"""
import re, sys
instrs = []
for ln in open(sys.argv[1]):
    m = re.match(r'\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);', ln)
    if m: instrs.append((int(m.group(1), 16), m.group(2).strip()))
print(len(instrs), "instructions,", len(instrs) * 16 // 1024, "KB")
loops = []
for a, t in instrs:
    m = re.search(r'BRA\S*\s+(0x[0-9a-f]+)', t)
    if m and int(m.group(1), 16) <= a: loops.append((a - int(m.group(1), 16), int(m.group(1), 16), a))
loops.sort(reverse=True)
for n, a0, a1 in loops[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    seg = [t for (a, t) in instrs if a0 <= a <= a1]
    cnt = lambda s: sum(s in t for t in seg)
    print(f"loop {a0:#07x}-{a1:#07x}: {n // 16:6d} instr ({n / 1024:6.1f} KB) DFMA={cnt('DFMA'):4d} DMUL={cnt('DMUL'):3d} LDS={cnt('LDS'):4d} BAR={cnt('BAR.'):3d} LDG={cnt('LDG'):3d} STG={cnt('STG'):3d} SHFL={cnt('SHFL'):3d}")
