"""Decodes the scheduling control bits of `cuobjdump -sass` output (two 64-bit words per instruction): stall count,
yield, write / read scoreboard ids and the wait mask.  usage: python tools/sass_ctl.py file.sass [regex]"""
import re
import sys

lines = open(sys.argv[1]).read().split("\n")
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
i = 0
n = 0
while i < len(lines):
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* 0x([0-9a-f]{16}) \*/", lines[i])
    if m and i + 1 < len(lines):
        m2 = re.match(r"\s*/\* 0x([0-9a-f]{16}) \*/", lines[i + 1])
        if m2:
            hi = int(m2.group(1), 16)
            ctl = hi >> 41                       # bits [105:128) of the 128-bit word
            stall = ctl & 0xF
            yld = (ctl >> 4) & 1
            wbar = (ctl >> 5) & 7
            rbar = (ctl >> 8) & 7
            wait = (ctl >> 11) & 0x3F
            txt = f"{n:5d} {m.group(1)} st={stall:2d} y={yld} w={wbar if wbar != 7 else '-'} r={rbar if rbar != 7 else '-'} wait={wait:06b} {m.group(2).strip()}"
            if pat is None or pat.search(m.group(2)):
                print(txt)
            n += 1
            i += 2
            continue
    i += 1
