#!/usr/bin/env python
"""Per-source-line instruction and stall-sample totals of one profiled kernel.

ncu's CSV source page is per SASS instruction; this joins it (by instruction order) with the line
table `nvdisasm -g` prints for the same function of the object file the kernel was built from.

    python tools/ncu_lines.py gpurun_out/x.ncu-rep umeregrobust_b200/csrc/build/moments.o 'VecAccILi8EEELb0' [top]
"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(obj, pattern):
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, check=True, capture_output=True)
        out = ""
        for cubin in glob.glob(os.path.join(d, "*.cubin")):
            out += subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    res, cur, inside = [], ("?", 0), False
    for ln in out.splitlines():
        if ln.startswith("\t.section\t.text.") or ln.startswith("\t.section\t.nv"):
            inside = (".text." in ln) and re.search(pattern, ln) is not None
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            res.append((int(m.group(1), 16), cur, m.group(2).strip()))
    return res


def main():
    rep, obj, pattern = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr_at = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_at]
    body = [r for r in rows[hdr_at + 1:] if len(r) == len(hdr)]
    ci, cs, cinst = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
    sass = sass_lines(obj, pattern)
    if len(sass) != len(body):
        print("warning: %d SASS instructions in the object, %d in the report" % (len(sass), len(body)))
    inst = collections.Counter()
    samp = collections.Counter()
    stall = collections.defaultdict(collections.Counter)
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot_i = tot_s = 0
    for (off, line, text), r in zip(sass, body):
        n, s = int(r[ci] or 0), int(r[cs] or 0)
        inst[line] += n
        samp[line] += s
        tot_i += n
        tot_s += s
        for i, h in stall_cols:
            v = int(r[i] or 0)
            if v:
                stall[line][h[6:]] += v
    print("total warp instructions %d, samples %d" % (tot_i, tot_s))
    print("%-22s %14s %6s %8s %6s  top stalls" % ("line", "warp-inst", "%", "samples", "%"))
    for line, n in sorted(inst.items(), key=lambda kv: -samp[kv[0]])[:top]:
        st = ", ".join("%s %d" % kv for kv in stall[line].most_common(3))
        print("%-22s %14d %6.2f %8d %6.2f  %s" % ("%s:%d" % line, n, 100.0 * n / max(tot_i, 1), samp[line],
                                                    100.0 * samp[line] / max(tot_s, 1), st))


if __name__ == "__main__":
    main()
