#!/usr/bin/env python
"""Decode the scheduling control bits (stall count, yield, barriers) of a SASS listing (sm_70+ 128-bit encoding).
usage: cuobjdump -sass -fun <mangled> file.o | python tools/sass_ctrl.py 0x42f0 0x4770"""
import re, sys
lo, hi = int(sys.argv[1], 16), int(sys.argv[2], 16)
lines = sys.stdin.read().splitlines()
i = 0
tot_stall = 0
n = 0
while i < len(lines):
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", lines[i])
    if m and i + 1 < len(lines):
        m2 = re.match(r"\s+/\* (0x[0-9a-f]+) \*/", lines[i + 1])
        a = int(m.group(1), 16)
        if m2 and lo <= a <= hi:
            w1 = int(m2.group(1), 16)
            ctrl = w1 >> 41
            stall = ctrl & 0xf
            yld = (ctrl >> 4) & 1
            wb = (ctrl >> 5) & 7
            rb = (ctrl >> 8) & 7
            wait = (ctrl >> 11) & 0x3f
            tot_stall += max(stall, 1)
            n += 1
            print(f"{a:5x} st={stall:2d} y={yld} wb={wb if wb != 7 else '-'} rb={rb if rb != 7 else '-'} wait={wait:06b}  {m.group(2)[:80]}")
        i += 2
    else:
        i += 1
print("instructions", n, "sum of stall counts", tot_stall)
