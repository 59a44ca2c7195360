#!/usr/bin/env python
"""Aggregate the per-instruction page of an ncu report (--page source --csv) into code regions:
samples and executed instructions per region (regions split at a list of SASS-offset boundaries).
usage: python tools/ncu_hotspots.py src.csv [bucket_instrs]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, isamp, iexec = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = [(int(r[ia], 16), r[isrc].strip(), int(r[isamp] or 0), int(r[iexec] or 0)) for r in rows[2:] if len(r) > iexec]
base = data[0][0]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 250
tot_s = sum(d[2] for d in data); tot_e = sum(d[3] for d in data)
print(f"instructions {len(data)}, samples {tot_s}, executed warp-instr {tot_e}")
for k in range(0, len(data), B):
    blk = data[k:k + B]
    s = sum(d[2] for d in blk); e = sum(d[3] for d in blk)
    ops = {}
    for d in blk:
        op = d[1].split()[0] if not d[1].startswith("@") else d[1].split()[1]
        op = op.split(".")[0]
        ops[op] = ops.get(op, 0) + d[2]
    top = ", ".join(f"{o}:{v}" for o, v in sorted(ops.items(), key=lambda kv: -kv[1])[:4])
    print(f"  [{(blk[0][0]-base):#07x}..{(blk[-1][0]-base):#07x}] samples {100*s/tot_s:5.1f}%  exec {100*e/tot_e:5.1f}%  exec/instr {e/len(blk)/1184:8.1f}  | {top}")
