#!/usr/bin/env python
"""Aggregate `ncu --page source --csv --print-source cuda,sass` output by CUDA source line.
usage: ncu_lines.py report.ncu-rep kernel_regex [launch_index] [--sort samples|instr]   (default: stall samples)"""
import csv, subprocess, sys, io, collections, os
args = [a for a in sys.argv[1:] if not a.startswith("--")]
sort_by = 1 if ("--sort" in sys.argv and sys.argv[sys.argv.index("--sort") + 1].startswith("instr")) else 0
sys.argv = [sys.argv[0]] + [a for i, a in enumerate(sys.argv[1:], 1) if not (a == "--sort" or sys.argv[i - 1] == "--sort")]
rep, kre = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
# sections start with "Kernel Name"? find by rows whose first cell == 'Kernel Name'
seen = set(); launch = -1; fname = None; hdr = None; agg = collections.OrderedDict(); tot_s = tot_i = 0
for r in rows:
    if not r: continue
    if r[0] == "File Path":
        fname = os.path.basename(r[1])
        if fname in seen or launch < 0: launch += 1; seen = set()
        seen.add(fname); continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if launch != which or hdr is None: continue
    if r[0] != "":      # a CUDA line summary row
        try:
            s = int(r[hdr.index("# Samples")]); i = int(r[hdr.index("Instructions Executed")])
        except ValueError:
            continue
        key = (fname, int(r[0]), r[1].strip()[:90])
        a = agg.setdefault(key, [0, 0, 0])
        a[0] += s; a[1] += i
        try: a[2] += int(r[hdr.index("stall_barrier")])
        except Exception: pass
        tot_s += s; tot_i += i
print("total samples", tot_s, "instr", tot_i)
for (f, ln, src), (s, i, b) in sorted(agg.items(), key=lambda kv: -kv[1][sort_by])[:45]:
    print("%5.1f%% smp %5.1f%% ins  %s:%d  %s" % (100.0 * s / max(tot_s, 1), 100.0 * i / max(tot_i, 1), f, ln, src))
