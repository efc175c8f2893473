#!/usr/bin/env python
"""Evidence table for profiles/: per kernel of libplaid_b200.so the tensor-core / TMA instruction counts in the SASS
(cuobjdump) and the register / spill figures ptxas printed at build time (csrc/build/*.o.log).
usage: python scripts/sass_table.py > profiles/rNN_sass_table.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "reranking_multimodal_retrievers_b200", "libplaid_b200.so")
BUILD = os.path.join(ROOT, "reranking_multimodal_retrievers_b200", "csrc", "build")
OPS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "SYNCS", "LDL", "STL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur:
            m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
            if m:
                op = m.group(1)
                counts[cur]["total"] += 1
                if op in OPS:
                    counts[cur][op] += 1
    regs = {}
    for f in sorted(os.listdir(BUILD)):
        if not f.endswith(".o.log"):
            continue
        log = open(os.path.join(BUILD, f)).read()
        for blk in re.split(r"ptxas info    : Compiling entry function '", log)[1:]:
            name = blk.split("'")[0]
            sp = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", blk)
            us = re.search(r"Used (\d+) registers", blk)
            regs[name] = (us.group(1) if us else "?", sp.group(2) if sp else "?", sp.group(3) if sp else "?")
    dm = demangle(list(counts))
    print("# SASS evidence: tcgen05 / TMEM / TMA instructions, registers and spills per kernel (sm_100a)\n")
    print("`cuobjdump -sass libplaid_b200.so` + ptxas `-v` output of the build; UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st,")
    print("UTMALDG = cp.async.bulk.tensor (TMA load), UBLKCP = cp.async.bulk (bulk copy, here shared -> global), LDGSTS = cp.async,")
    print("SYNCS = mbarrier ops, LDL / STL = local-memory (spill) accesses.\n")
    print("| kernel | SASS instr | " + " | ".join(OPS) + " | regs | spill st / ld [B] |")
    print("|---|---|" + "---|" * len(OPS) + "---|---|")
    for name, c in counts.items():
        r = regs.get(name, ("?", "?", "?"))
        short = dm[name].replace("plaid::", "").split("(")[0].replace("void ", "")
        print(f"| `{short}` | {c['total']} | " + " | ".join(str(c[o]) if c[o] else "" for o in OPS) + f" | {r[0]} | {r[1]} / {r[2]} |")


if __name__ == "__main__":
    main()
