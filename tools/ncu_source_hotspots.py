#!/usr/bin/env python
"""Per-instruction warp-stall samples of one kernel from `ncu -i X.ncu-rep --page source --csv [--kernel-name regex:..]`:
share of the samples per block of SASS instructions (with the execution count, which tells the loops apart) and the
most-stalled instructions with their stall reasons.  python tools/ncu_source_hotspots.py source.csv [block] [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
kern = []
for r in rows:
    if r and r[0] == "Kernel Name":
        kern.append([r[1], []])
    elif r and r[0].startswith("0x"):
        kern[-1][1].append(r)
name, data = kern[0]
ia, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
names = ["stall_wait", "stall_math", "stall_long_sb", "stall_short_sb", "stall_selected", "stall_not_selected", "stall_no_inst",
         "stall_branch_resolving", "stall_dispatch", "stall_mio", "stall_lg"]
cols = {n: hdr.index(n) for n in names}
iconf = hdr.index("L1 Wavefronts Shared Excessive")
tot = sum(int(r[isamp]) for r in data)
print(f"kernel {name}\ntotal samples {tot}, {len(data)} SASS instructions")
segs = [(i, int(r[isamp]), int(r[iex]), r[ia].strip()[:70], {k[6:]: int(r[v]) for k, v in cols.items() if int(r[v])}, int(r[iconf] or 0))
        for i, r in enumerate(data)]
step = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print("first instr | samples | share | max executions in block | excess shared wavefronts")
for b in range(0, len(segs), step):
    blk = segs[b:b + step]
    ss = sum(x[1] for x in blk)
    if ss:
        print(f"{b:5d} {ss:6d} {100 * ss / tot:5.1f}% {max(x[2] for x in blk):9d} {sum(x[5] for x in blk):8d}")
print("--- most-stalled instructions: index, samples, executions, SASS, stall reasons")
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
for s in sorted(sorted(segs, key=lambda x: -x[1])[:top], key=lambda x: x[0]):
    print(s[0], s[1], s[2], s[3], s[4])
