"""Decodes the scheduling control bits of SASS (sm_70+ 128-bit encoding) from `cuobjdump -sass` output: stall count,
write / read scoreboard, wait mask.  usage: python scripts/sass_ctrl.py file.sass 0xSTART 0xEND"""
import re, sys
lines = open(sys.argv[1]).read().split("\n")
a0, a1 = int(sys.argv[2], 16), int(sys.argv[3], 16)
i = 0
while i < len(lines):
    m = re.match(r'\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);\s*/\* (0x[0-9a-f]+) \*/', lines[i])
    if m and i + 1 < len(lines):
        m2 = re.match(r'\s*/\* (0x[0-9a-f]+) \*/', lines[i + 1])
        if m2:
            addr = int(m.group(1), 16)
            if a0 <= addr <= a1:
                hi = int(m2.group(1), 16)
                ctrl = hi >> 41
                stall = ctrl & 0xf; yld = (ctrl >> 4) & 1; wr = (ctrl >> 5) & 7; rd = (ctrl >> 8) & 7; wait = (ctrl >> 11) & 0x3f
                print(f"{addr:#07x} {m.group(2)[:44]:44s} stall={stall:2d} {'Y' if yld else ' '} wr={'-' if wr == 7 else wr} rd={'-' if rd == 7 else rd} wait={''.join(str(b) for b in range(6) if wait >> b & 1) or '-'}")
            i += 2
            continue
    i += 1
