#!/usr/bin/env python
"""Counts the Blackwell tensor-core / TMA / tensor-memory opcodes in the built libjrr.so and writes
profiles/sass_opcodes.txt (cuobjdump -sass; runs without a GPU)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "joint-regressor-refinement_b200", "libjrr.so")
OPS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCCP", "UTCATOM", "SYNCS", "HMMA", "FFMA2"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per_fn = collections.OrderedDict()
    fn = None
    archs = set(re.findall(r"arch = (sm_\w+)", sass))
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            per_fn[fn] = collections.Counter()
            continue
        if fn is None:
            continue
        for op in OPS:
            if re.search(r"\b" + op + r"\b|\b" + op + r"\.", line):
                per_fn[fn][op] += 1
    total = collections.Counter()
    for c in per_fn.values():
        total.update(c)
    out = [f"cuobjdump -sass {os.path.relpath(LIB, ROOT)}   (arch: {', '.join(sorted(archs))})", "",
           "total opcode counts:"] + [f"  {op:8s} {total[op]}" for op in OPS if total[op]] + ["", "per kernel (tensor-core / TMA / TMEM users only):"]
    for fn, c in per_fn.items():
        if any(c[o] for o in ("UTCHMMA", "UTMALDG", "LDTM", "STTM", "HMMA")):
            out.append(f"  {fn}: " + ", ".join(f"{o} x{c[o]}" for o in OPS if c[o]))
    txt = "\n".join(out) + "\n"
    open(os.path.join(ROOT, "profiles", "sass_opcodes.txt"), "w").write(txt)
    sys.stdout.write(txt)


if __name__ == "__main__":
    main()
