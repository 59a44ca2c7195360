#!/usr/bin/env python
"""List backward branches (loops) of a kernel's SASS with their instruction counts and opcode mix.
usage: cuobjdump -sass -fun <mangled> file.o | python tools/sass_loops.py"""
import re, sys, collections
ins = []
for ln in sys.stdin:
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        ins.append((int(m.group(1), 16), m.group(2)))
addr = {a: i for i, (a, _) in enumerate(ins)}
print("total instructions", len(ins))
for i, (a, t) in enumerate(ins):
    m = re.search(r"\bBRA(?:\.\w+)*\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < a and tgt in addr:
            body = ins[addr[tgt]:i + 1]
            ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", x[1]).split()[0].split(".")[0] for x in body)
            print(f"loop {tgt:#x}..{a:#x}: {len(body)} instrs ({len(body)*16/1024:.1f} KB)", dict(ops.most_common(12)))
