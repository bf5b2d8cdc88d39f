"""Per-kernel SASS evidence of the shipped libgnna_b200.so (runs without a GPU: cuobjdump + c++filt).

For every kernel of the library: registers, stack bytes (spills), static shared memory, instruction count and how often the mnemonics occur that say what the kernel is built
from (B200_PROFILING.md, "What proves a Blackwell-native kernel"): UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st,
UTCBAR = tcgen05.commit, UBLKCP / UTMALDG = TMA bulk copies, SYNCS = mbarrier, LDGSTS = cp.async, LDG.E.128 = 128-bit
gathers, REDG.E.ADD.F32x4 = red.global.add.v4.f32, FHADD.BF16 = mixed-precision bf16 accumulate, HMMA = legacy mma.sync.

    python tools/sass_summary.py [path/to/lib.so]      > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gnnadvisor_osdi21_b200", "libgnna_b200.so")
CUOBJDUMP = "/usr/local/cuda/bin/cuobjdump"

PATTERNS = [
    ("tcgen05.mma", r"\bUTC[A-Z]*MMA"), ("tcgen05.ld/st", r"\b(LDTM|STTM)"), ("tcgen05.commit", r"\bUTCBAR"),
    ("tmem alloc", r"\bUTCATOMSWS"), ("TMA bulk", r"\b(UBLKCP|UTMALDG|UTMASTG)"), ("mbarrier", r"\bSYNCS"),
    ("cp.async", r"\bLDGSTS"), ("LDG.128", r"\bLDG\.E\.(?:[A-Z0-9_]+\.)*128"), ("red.v4", r"\bREDG\.E\.ADD\.F32x4"),
    ("bf16 add", r"\bFHADD"), ("mma.sync", r"\bHMMA"),
]


def short(name):
    name = re.sub(r"\(.*$", "", name)                      # drop the parameter list
    return name.replace("gnna::", "").replace("(anonymous namespace)::", "")


def main():
    sass = subprocess.run([CUOBJDUMP, "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = []
            continue
        if cur is not None and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            kernels[cur].append(line)
    res = {}
    usage = subprocess.run([CUOBJDUMP, "-res-usage", LIB], capture_output=True, text=True, check=True).stdout.splitlines()
    for i, line in enumerate(usage):
        m = re.match(r"\s*Function (\S+):", line)
        if m and i + 1 < len(usage):
            f = dict(kv.split(":") for kv in usage[i + 1].split() if ":" in kv and "[" not in kv)
            res[m.group(1)] = (int(f.get("REG", 0)), int(f.get("STACK", 0)), int(f.get("SHARED", 0)))
    names = list(kernels)
    dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    print("SASS summary of %s (sm_100a, %d kernels); columns = occurrences per kernel" % (os.path.relpath(LIB, ROOT), len(names)))
    head = ["regs", "stack B", "static smem", "instr"] + [p[0] for p in PATTERNS]
    print("%-78s %s" % ("kernel", " ".join("%13s" % h for h in head)))
    totals = collections.Counter()
    rows = []
    for mangled, pretty in zip(names, dem):
        body = kernels[mangled]
        counts = list(res.get(mangled, (0, 0, 0))) + [len(body)] + [sum(1 for ln in body if re.search(rx, ln)) for _, rx in PATTERNS]
        for h, c in zip(head, counts):
            totals[h] = max(totals[h], c) if h in ("regs", "stack B", "static smem") else totals[h] + c
        rows.append((short(pretty), counts))
    for name, counts in sorted(rows):
        print("%-78s %s" % (name[:78], " ".join("%13d" % c for c in counts)))
    print("%-78s %s" % ("TOTAL (max for regs / stack / smem)", " ".join("%13d" % totals[h] for h in head)))
    tc = sorted({n.split("<")[0] for n, c in rows if c[4] > 0})
    tma = sorted({n.split("<")[0] for n, c in rows if c[8] > 0})
    print("\nkernels with a stack frame (spills or local arrays): %s" % (", ".join(sorted({n.split('<')[0] for n, c in rows if c[1] > 0})) or "none"))
    print("\nkernels issuing tcgen05.mma: %s" % ", ".join(tc))
    print("kernels using TMA bulk copies: %s" % ", ".join(tma))
    print("kernels with legacy mma.sync (HMMA): %s" % (", ".join(sorted({n.split('<')[0] for n, c in rows if c[-1] > 0})) or "none"))


if __name__ == "__main__":
    main()
