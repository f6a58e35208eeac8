#!/usr/bin/env python
"""Summarise the SASS page of an .ncu-rep: instructions executed per opcode class and the hot
address ranges (basic blocks by equal execution count).  Usage: ncu_sass.py report.ncu-rep [kernel-regex]"""
import csv, subprocess, sys, re, collections
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = None
ins = []
for r in rows:
    if r and r[0] == "Address":
        hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        ins.append((d["Address"], d["Source"].strip(), int(d["Instructions Executed"] or 0), int(d["# Samples"] or 0),
                    int(d["Thread Instructions Executed"] or 0)))
tot = sum(i[2] for i in ins); smp = sum(i[3] for i in ins)
print("instructions:", len(ins), "executed (warp):", tot, "samples:", smp)
cls = collections.Counter(); cs = collections.Counter()
for a, s, n, sm, t in ins:
    op = s.split()[0] if not s.startswith("@") else s.split()[1]
    op = op.split(".")[0]
    cls[op] += n; cs[op] += sm
print("-- by opcode")
for op, n in cls.most_common(28):
    print("  %-10s %6.2f%%  samples %5.2f%%" % (op, 100 * n / tot, 100 * cs[op] / max(smp, 1)))
# blocks: consecutive instructions with the same exec count
print("-- hot blocks (consecutive SASS with equal exec count)")
blocks = []
start = 0
for k in range(1, len(ins) + 1):
    if k == len(ins) or ins[k][2] != ins[start][2]:
        n = ins[start][2]
        blocks.append((n * (k - start), start, k, n, sum(i[3] for i in ins[start:k]), sum(i[4] for i in ins[start:k])))
        start = k
blocks.sort(reverse=True)
for w, s, e, n, sm, t in blocks[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    ops = collections.Counter((x[1].split()[0] if not x[1].startswith("@") else x[1].split()[1]).split(".")[0] for x in ins[s:e])
    print("  %5.2f%% of instr, %5.2f%% samples, [%d..%d) %d instr x %d, lanes %.1f : %s" % (
        100 * w / tot, 100 * sm / max(smp, 1), s, e, e - s, n, t / max(w, 1), dict(ops.most_common(8))))
