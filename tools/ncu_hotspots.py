#!/usr/bin/env python
"""Per-source-line instruction counts of one kernel from an ncu report.

    python tools/ncu_hotspots.py <report.ncu-rep> <lib.so> <kernel-substring> [--top 40]

ncu's CSV source page is SASS-level; the line table comes from `nvdisasm -g` on the cubin embedded
in the shared library (same build), joined on the instruction address."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, lib, pat = sys.argv[1:4]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(out)):
        if row and row[0] == "Kernel Name":
            cur = {"name": row[1], "rows": [], "hdr": None}
            blocks.append(cur)
        elif cur is not None and row and row[0] == "Address":
            cur["hdr"] = row
        elif cur is not None and cur["hdr"] and len(row) == len(cur["hdr"]):
            cur["rows"].append(row)
    base_name = pat.split("<")[0]
    blk = [b for b in blocks if base_name in b["name"]][0]          # (ncu prints template arguments as "(int)5, (int)4")
    if "<" not in pat and "<" in blk["name"]:
        # the report names the instance ("k<(int)5, (int)32, (int)2>(...)"): count only that instance's SASS, not every
        # template instance of the kernel that happens to share instruction offsets
        args = re.findall(r"\(int\)(\d+)|\(bool\)(\w+)", blk["name"].split("(const")[0].split(">(")[0])
        ints = [a if a else ("1" if b == "true" else "0") for a, b in args]
        if ints and all(a for a, _ in args):
            pat = "%s<%s>" % (pat, ",".join(ints))
    hdr = blk["hdr"]
    ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    base = int(blk["rows"][0][ia], 16)
    counts = {int(r[ia], 16) - base: (int(r[ii] or 0), int(r[isamp] or 0), r[hdr.index("Source")]) for r in blk["rows"]}
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    mangled = None
    per_line, per_fn = collections.Counter(), collections.Counter()
    samp_line = collections.Counter()
    line, inside, inl = ("?", 0), False, ""
    key = re.sub(r"[^A-Za-z0-9_]", "", pat.split("<")[0].split("::")[-1])
    for l in dis.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
        if m:
            # integer template arguments are mangled as I Li<n>E ... E right after the name
            frag = key + ("I" + "".join("Li%sE" % t for t in re.findall(r"\d+", pat.split("<")[1])) if "<" in pat else "")
            inside = frag in m.group(1)
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
        if m:
            line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            addr = int(m.group(1), 16)
            if addr in counts:
                per_line[line] += counts[addr][0]
                samp_line[line] += counts[addr][1]
    total = sum(per_line.values())
    print("kernel %s: %d warp instructions attributed" % (blk["name"][:70], total))
    print("%-28s %12s %7s %8s" % ("file:line", "warp-inst", "share", "samples"))
    for (f, n), c in per_line.most_common(top):
        print("%-28s %12d %6.1f%% %8d" % ("%s:%d" % (f, n), c, 100.0 * c / max(total, 1), samp_line[(f, n)]))
    per_file = collections.Counter()
    for (f, n), c in per_line.items():
        per_file[f] += c
    print({k: "%.1f%%" % (100.0 * v / max(total, 1)) for k, v in per_file.items()})


if __name__ == "__main__":
    main()
