#!/usr/bin/env python
"""Per-kernel SASS / resource summary of libfcapp_cuda.so (no GPU needed): registers, stack, static shared memory from
`cuobjdump --dump-resource-usage`, and the counts of the sm_100a instructions that show what the kernel is built from --
UBLKCP (1-D bulk copy = TMA engine), SYNCS (mbarrier), BAR, ATOM/RED, DFMA/DADD/DMUL (FP64 pipe), LDG/STG, LDS/STS -- from
`cuobjdump -sass`.  Writes one line per kernel.

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "freecappuccino_b200", "libfcapp_cuda.so")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", SO], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            usage[cur] = dict(re.findall(r"(REG|STACK|SHARED|LOCAL):(\d+)", line))
            cur = None
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    counts = collections.defaultdict(collections.Counter)
    arch = collections.Counter()
    cur = None
    keys = ("UBLKCP", "UTMALDG", "SYNCS", "BAR", "ATOM", "RED", "DFMA", "DADD", "DMUL", "LDG", "STG", "LDS", "STS", "MEMBAR",
            "FENCE", "CCTL", "TCGEN", "HMMA", "UTC")
    for line in sass.splitlines():
        m = re.search(r"arch = (sm_\w+)", line)
        if m:
            arch[m.group(1)] += 1
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            op = m.group(1)
            for k in keys:
                if op.startswith(k):
                    counts[cur][k] += 1
            counts[cur]["_all"] += 1
    names = sorted(usage)
    dm = demangle(names)
    print(f"# {os.path.relpath(SO, ROOT)}: {dict(arch)} code objects; {len(names)} kernels")
    tot = collections.Counter()
    for n in names:
        for k, v in counts[n].items():
            tot[k] += v
    print("# whole library: " + ", ".join(f"{k} {tot[k]}" for k in keys if tot[k]))
    print("# kernel | registers stack shared | instructions | " + " ".join(keys[:13]))
    for n in names:
        short = re.sub(r"\(anonymous namespace\)::", "", dm.get(n, n))
        short = re.sub(r"\(.*", "", short)[:90]
        u, c = usage[n], counts[n]
        print(f"{short:90s} | {u.get('REG', '?'):>3} {u.get('STACK', '0'):>4} {u.get('SHARED', '0'):>6} | {c['_all']:>6} | "
              + " ".join(f"{c[k]:>4}" for k in keys[:13]))


if __name__ == "__main__":
    main()
