#!/usr/bin/env python
"""Registers / stack / shared memory (cuobjdump --dump-resource-usage) and SASS instruction counts of the hot
kernels of libspartacus_b200.so.  Usage: python profiles/resource_usage.py > profiles/r02_resource_usage.txt"""
import collections
import re
import subprocess

SO = "spartacus_surface_b200/csrc/libspartacus_b200.so"
HOT = r"k_fast_(layer|sweeps)_\w+<3, (2|4)|k_partition_layers<3, (2|4)|k_stage|k_rec_(up|down)_\w+<3, 2, true|k_fused_\w+<3, 2, true|k_column_keys"
OPS = ("DFMA", "DMUL", "DADD", "MUFU", "LDG", "STG", "LDS", "STS", "LDL", "STL", "SHFL", "BAR")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    lines = subprocess.run(["cuobjdump", "--dump-resource-usage", SO], capture_output=True, text=True).stdout.splitlines()
    raw = {}
    for i, l in enumerate(lines):
        m = re.match(r"\s*Function (\S+):", l)
        if m and i + 1 < len(lines):
            raw[m.group(1)] = lines[i + 1].strip()
    counts = collections.defaultdict(collections.Counter)
    cur = None
    for l in subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout.splitlines():
        m = re.match(r"\s*Function : (\S+)", l)
        if m:
            cur = m.group(1)
            continue
        if cur and re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
            counts[cur]["instr"] += 1
            for op in OPS:
                if re.search(r"\s" + op + r"[\s.]", l):
                    counts[cur][op] += 1
    names = demangle(list(raw))
    print("cuobjdump --dump-resource-usage / -sass of", SO, "(sm_100a), hot kernels at nreg = 3")
    print("(REG registers per thread, STACK bytes of local memory per thread; SASS instruction counts: instr = all)\n")
    for mangled in sorted(raw, key=lambda k: names[k]):
        n = names[mangled]
        if not re.search(HOT, n):
            continue
        c = counts[mangled]
        print(n)
        print("    " + raw[mangled])
        print("    SASS: " + " ".join(f"{k}={c.get(k, 0)}" for k in ("instr",) + OPS))


if __name__ == "__main__":
    main()
