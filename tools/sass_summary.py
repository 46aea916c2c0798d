"""Per-kernel SASS summary of liblanedet.so: instruction count, registers, and the mnemonics that prove what the kernel uses
(UTMALDG / UTMASTG = tensor-map TMA load / store, UBLKCP = bulk copy, SYNCS = mbarrier, LDL / STL = local-memory spills, ...).

    python tools/sass_summary.py                      # table -> stdout (commit it as profiles/sass_summary.txt)
    python tools/sass_summary.py --dump k_mask        # the SASS of every kernel whose name contains k_mask -> stdout
"""
import argparse
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "advanced-lane-detection_b200", "csrc", "liblanedet.so")
KEYS = ["UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "LDL", "STL", "LDS", "STS", "LDG", "STG", "IDP", "SHFL", "BAR", "ATOMS", "VOTE"]


def demangle(names):
    out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.split("\n")
    return [re.sub(r"\(.*", "", o).replace("lanedet::", "").replace("void ", "") for o in out[:len(names)]]


def functions(lib):
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    cur, funcs = None, collections.OrderedDict()
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m and cur:
            funcs[cur].append((m.group(1), m.group(2).strip()))
    return funcs


def registers(lib):
    res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
    regs, cur = {}, None
    for line in res.split("\n"):
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+).*?SHARED:(\d+)", line)
        if m and cur:
            regs[cur] = (int(m.group(1)), int(m.group(2)))
    return regs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=LIB)
    ap.add_argument("--dump", default=None)
    a = ap.parse_args()
    funcs = functions(a.lib)
    names = list(funcs)
    pretty = dict(zip(names, demangle(names)))
    if a.dump:
        for n in names:
            if a.dump in pretty[n]:
                print("// ----", pretty[n])
                for addr, ins in funcs[n]:
                    print(addr, ins)
        return
    regs = registers(a.lib)
    print("%-34s %6s %5s %7s " % ("kernel", "instr", "regs", "smem") + " ".join("%7s" % k for k in KEYS))
    for n in sorted(names, key=lambda s: pretty[s]):
        cnt = collections.Counter()
        for _, ins in funcs[n]:
            op = ins.split()[1] if ins.startswith("@") else ins.split()[0]
            for k in KEYS:
                if op.startswith(k):
                    cnt[k] += 1
        r = regs.get(n, (0, 0))
        print("%-34s %6d %5d %7d " % (pretty[n][:34], len(funcs[n]), r[0], r[1]) + " ".join("%7d" % cnt[k] for k in KEYS))


if __name__ == "__main__":
    main()
