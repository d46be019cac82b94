#!/usr/bin/env python
"""Executed warp instructions per SOURCE LINE of one kernel of an ncu capture, with the FP64 share of each line.

    python profiles/sass_by_line.py gpurun_out/r02_cfg3.ncu-rep intensity_kernel intensity_kernel 'ILi2ELb0ELb1' [top]

ncu's own source page cannot resolve the files of a capture taken on the GPU box (other paths), so the per-instruction
counts of its SASS page are joined, by instruction offset, with the line table `nvdisasm -g` prints for the same
function of the in-tree library (the library must be the build that was profiled).  Arguments: the report, a regex
for the kernel name in the report, the source file stem whose cubin holds the kernel, a substring of the mangled
function name, and how many lines to print."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")


def sass_counts(rep, kernel_re, launch=0):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name",
                          "regex:" + kernel_re], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    h = heads[launch]
    end = heads[launch + 1] - 1 if launch + 1 < len(heads) else len(rows)
    col = {n: i for i, n in enumerate(rows[h])}
    recs = []
    for r in rows[h + 1:end]:
        if len(r) < len(col):
            continue
        try:
            recs.append((int(r[0], 16), r[col["Source"]].strip(), int(r[col["Instructions Executed"]]),
                         int(r[col["# Samples"]] or 0)))
        except ValueError:
            continue
    base = recs[0][0]
    return [(a - base, s, n, smp) for a, s, n, smp in recs], rows[h - 1][1] if h else ""


def line_table(stem, func_sub):
    lib = os.path.abspath(os.environ.get("XRSD_B200_LIB") or os.path.join(ROOT, "xrsdkit_b200", "libxrsd_b200.so"))      # the build that was profiled
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", stem, lib], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
    table, on, cur = {}, False, None
    for ln in txt:
        if ln.startswith(".text."):
            on = func_sub in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            table[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return table


def opcode(s):
    t = s.split()
    if t and t[0].startswith("@"):
        t = t[1:]
    return t[0].split(".")[0] if t else "?"


def main():
    rep, kre, stem, fsub = sys.argv[1:5]
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    recs, name = sass_counts(rep, kre)
    table = line_table(stem, fsub)
    per = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
    tot = tot64 = 0
    miss = 0
    for off, src, n, smp in recs:
        ent = table.get(off)
        if ent is None or opcode(ent[1]) != opcode(src):
            miss += 1
            key = ("?", 0)
        else:
            key = ent[0]
        op = opcode(src)
        p = per[key]
        p[0] += n
        p[2] += smp
        p[3][op] += n
        if op in FP64:
            p[1] += n
            tot64 += n
        tot += n
    print("# %s\n# %d warp instructions executed, %.1f %% FP64 (DFMA/DMUL/DADD/DSETP); %d SASS rows not matched" %
          (name[:110], tot, 100.0 * tot64 / tot, miss))
    print("# %-22s %6s %6s %6s  %s" % ("file:line", "inst%", "fp64%", "smp%", "top non-FP64 opcodes"))
    tsmp = sum(p[2] for p in per.values()) or 1
    for key, p in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        others = ", ".join("%s %.1f" % (o, 100.0 * c / tot) for o, c in p[3].most_common(8) if o not in FP64)
        print("%-24s %6.2f %6.1f %6.2f  %s" % ("%s:%d" % key, 100.0 * p[0] / tot, 100.0 * p[1] / max(1, p[0]), 100.0 * p[2] / tsmp, others))


if __name__ == "__main__":
    main()
