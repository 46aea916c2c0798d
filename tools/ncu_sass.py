#!/usr/bin/env python
"""Per-SASS-instruction view of one kernel of an ncu report (--import-source on): opcode mix by executed warp instructions,
the instructions with the most stall samples, and execution counts of selected opcodes.
usage: ncu_sass.py report.ncu-rep kernel_regex [launch_index] [--top N] [--grep OPCODE]"""
import csv, io, subprocess, sys, collections
args = [a for a in sys.argv[1:] if not a.startswith("--")]
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
grep = sys.argv[sys.argv.index("--grep") + 1] if "--grep" in sys.argv else None
args = [a for a in args if a not in (str(top), grep)]
rep, kre = args[0], args[1]
which = int(args[2]) if len(args) > 2 else 0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
launch, hdr, rows = -1, None, []
for r in csv.reader(io.StringIO(txt)):
    if not r: continue
    if r[0] == "Kernel Name": launch += 1; continue
    if r[0] == "Address": hdr = r; continue
    if launch == which and hdr: rows.append(r)
iS, iI, iSrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
tot_s = sum(int(r[iS]) for r in rows); tot_i = sum(int(r[iI]) for r in rows)
print("instructions executed %d, stall samples %d, %d SASS lines" % (tot_i, tot_s, len(rows)))
mix = collections.Counter(); smp = collections.Counter()
for r in rows:
    t = r[iSrc].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    mix[op] += int(r[iI]); smp[op] += int(r[iS])
print("opcode mix (executed warp instructions, stall samples):")
for op, n in mix.most_common(22): print("  %-10s %5.1f%% ins %5.1f%% smp" % (op, 100.0 * n / tot_i, 100.0 * smp[op] / max(tot_s, 1)))
print("top stall samples:")
for idx, r in sorted(enumerate(rows), key=lambda kv: -int(kv[1][iS]))[:top]:
    print("  line %4d %5.2f%% smp  exec %9s  %s" % (idx, 100.0 * int(r[iS]) / max(tot_s, 1), r[iI], r[iSrc].strip()[:90]))
if grep:
    print("executions of", grep)
    for idx, r in enumerate(rows):
        if grep in r[iSrc]: print("  line %4d exec %9s smp %5s  %s" % (idx, r[iI], r[iS], r[iSrc].strip()[:90]))
