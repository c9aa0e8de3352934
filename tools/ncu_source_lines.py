#!/usr/bin/env python
"""Per-source-line view of one kernel of an ncu capture (development tool).

ncu's source page lists SASS instructions with their executed counts and stall samples; nvdisasm -gi lists the same
instructions with the (inlined) source lines they came from.  This joins the two and sums instructions / samples per
line of the row function, so that "where do the 1878 instructions of a row go" has an answer:

    ncu -i capture.ncu-rep --page source --csv > src.csv
    cuobjdump -xelf all isochrones_b200/_build/iso_lnpost.o          # -> iso_lnpost.sm_100a.cubin (same build!)
    python tools/ncu_source_lines.py src.csv iso_lnpost.sm_100a.cubin '_Z17iso_lnpost_kernelILi1ELb0ELi1ELb1EEv15IsoLnpostParams' \
        --warps 31250 --file iso_lnpost_row.cuh

--warps: warps that executed the kernel body once (rows / 32) — turns executed counts into instructions per row.
"""
import argparse
import collections
import csv
import os
import re
import subprocess


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("source_csv")
    ap.add_argument("cubin")
    ap.add_argument("kernel", help="mangled kernel name")
    ap.add_argument("--warps", type=float, default=31250.0)
    ap.add_argument("--file", default="iso_lnpost_row.cuh", help="attribute to the outermost frame in this file")
    args = ap.parse_args()

    dis = subprocess.run(["nvdisasm", "-gi", args.cubin], capture_output=True, text=True, check=True).stdout.split("\n")
    start = next(i for i, l in enumerate(dis) if l.startswith(".text." + args.kernel + ":"))
    end = next((i for i in range(start + 1, len(dis)) if dis[i].startswith(".text.")), len(dis))
    ins, block, last = [], [], None
    for line in dis[start:end]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            block.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?);", line):
            if block:
                last, block = block, []
            ins.append(last)
    rows = list(csv.reader(open(args.source_csv)))
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    if len(ins) != len(data):
        raise SystemExit("the cubin is not the build that was profiled: %d vs %d instructions" % (len(ins), len(data)))
    src_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "isochrones_b200", "csrc")
    text = open(os.path.join(src_dir, args.file)).read().split("\n")
    agg = collections.defaultdict(lambda: [0.0] * 6)
    for frames, r in zip(ins, data):
        key = next((f for f in reversed(frames or []) if f[0] == args.file), (frames or [("?", 0)])[-1])
        a = agg[key]
        a[0] += float(r[ix["Instructions Executed"]]) / args.warps
        a[1] += float(r[ix["# Samples"]])
        a[2] += 1
        for j, k in enumerate(("stall_long_sb", "stall_short_sb", "stall_wait")):
            a[3 + j] += float(r[ix[k]])
    print("instructions per row %.1f, stall samples %d" % (sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())))
    print("file:line  instr/row  samples (long-scoreboard short-scoreboard wait)  static | source")
    for k, v in sorted(agg.items()):
        if v[0] >= 2 or v[1] > 20:
            t = text[k[1] - 1].strip()[:70] if k[0] == args.file and k[1] <= len(text) else ""
            print("%-20s %4d  %6.1f  %5d (L %4d S %4d W %4d)  %4d | %s" % (k[0], k[1], v[0], v[1], v[3], v[4], v[5], v[2], t))


if __name__ == "__main__":
    main()
