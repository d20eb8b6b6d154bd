#!/usr/bin/env python
"""Static SASS census of one kernel of libsphk.so: instructions per source line and per source file.

    python tools/sass_static.py <lib.so> <kernel-substring> [--lines FILE:LO-HI ...] [--top 40]

No GPU needed: `cuobjdump -xelf` + `nvdisasm -g -c` (the library is built with -lineinfo).  It counts
instructions as they sit in the binary (every instruction once, loops not weighted), which is what one
needs to compare two formulations of the same straight-line code (the per-pair batch body) before
spending GPU time on them."""
import collections
import os
import re
import subprocess
import sys
import tempfile


def disassemble(lib):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True, check=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    return subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True, check=True).stdout


def census(dis, pat):
    per_line, per_file, ops = collections.Counter(), collections.Counter(), collections.Counter()
    inside, line = False, ("?", 0)
    total = 0
    for l in dis.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
        if m:
            inside = pat in m.group(1)
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", l)
        if m:
            per_line[line] += 1
            per_file[line[0]] += 1
            ops[m.group(1).split(".")[0]] += 1
            total += 1
    return total, per_line, per_file, ops


def main():
    lib, pat = sys.argv[1:3]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    ranges = []
    if "--lines" in sys.argv:
        for a in sys.argv[sys.argv.index("--lines") + 1:]:
            if a.startswith("--"):
                break
            f, r = a.split(":")
            lo, hi = r.split("-")
            ranges.append((f, int(lo), int(hi)))
    total, per_line, per_file, ops = census(disassemble(lib), pat)
    print("kernel *%s*: %d SASS instructions" % (pat, total))
    print("per file:", dict(per_file.most_common()))
    print("opcodes :", dict(ops.most_common(24)))
    for f, lo, hi in ranges:
        n = sum(c for (ff, ln), c in per_line.items() if ff == f and lo <= ln <= hi)
        print("  %s:%d-%d  %d instructions" % (f, lo, hi, n))
    for (f, ln), c in per_line.most_common(top):
        print("  %-22s %5d  %d" % (f, ln, c))


if __name__ == "__main__":
    main()
