#!/usr/bin/env python
"""Hottest SASS instructions of a kernel in an ncu report, each counted once, with the source lines
(inline stack) that contain it:  python tools/ncu_sass.py report.ncu-rep kernel_regex [top_n]"""
import collections
import csv
import io
import subprocess
import sys


def page(rep, kern, mode):
    return subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", mode, "-k", "regex:" + kern],
                          capture_output=True, text=True).stdout


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    where = collections.defaultdict(list)
    cur, hdr, line = None, None, None
    for row in csv.reader(io.StringIO(page(rep, kern, "cuda,sass"))):
        if not row:
            continue
        if row[0] == "File Path":
            cur, hdr = row[1].split("/")[-1], None
        elif row[0] == "Line No":
            hdr = row
        elif hdr and row[0] not in ("", "Function Name"):
            line = "%s:%s" % (cur, row[0])
        elif hdr and row[0] == "" and len(row) > 2 and row[2].startswith("0x"):
            where[row[2]].append(line)
    rows, hdr = [], None
    for row in csv.reader(io.StringIO(page(rep, kern, "sass"))):
        if row and row[0] == "Address":
            hdr = row
        elif hdr and row and row[0].startswith("0x"):
            rows.append(row)
    si = hdr.index("# Samples")
    stall = [(i, c[6:]) for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
    total = sum(int(r[si] or 0) for r in rows)
    reasons = collections.Counter()
    for r in rows:
        for i, c in stall:
            reasons[c] += int(r[i] or 0)
    print("total samples", total, " by reason:", " ".join("%s=%.1f%%" % (c, 100.0 * v / total) for c, v in reasons.most_common(8)))
    rows.sort(key=lambda r: -int(r[si] or 0))
    for r in rows[:top]:
        s = int(r[si] or 0)
        st = sorted(((int(r[i] or 0), c) for i, c in stall), reverse=True)[:2]
        print("%5.1f%% %-44s %s | %s" % (100.0 * s / total, r[1].strip()[:44], " ".join("%s=%d" % (c, v) for v, c in st if v),
                                       " < ".join(dict.fromkeys(where.get(r[0], ["?"])))))


if __name__ == "__main__":
    main()
