#!/usr/bin/env python
"""Per-source-line summary of an ncu report's source page (needs -lineinfo + --import-source on):
python tools/ncu_lines.py report.ncu-rep kernel_regex [top_n]  -> samples, instructions and the
dominant stall reasons of the hottest CUDA source lines, per file."""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kern],
                         capture_output=True, text=True).stdout
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(out)):
        if not row:
            continue
        if row[0] in ("File Name", "File Path"):
            cur = {"file": row[1], "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and row[0] == "Line No":
            cur["hdr"] = row
        elif cur is not None and cur["hdr"] is not None and row[0] not in ("", "Function Name"):
            cur["rows"].append(row)
    total = 0
    lines = []
    for b in blocks:
        h = b["hdr"]
        if "# Samples" not in h:
            continue
        si, ii = h.index("# Samples"), h.index("Instructions Executed")
        stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
        for r in b["rows"]:
            try:
                s = int(r[si])
            except (ValueError, IndexError):
                continue
            total += s
            if s:
                stalls = sorted(((int(r[i] or 0), c[6:]) for i, c in stall_cols), reverse=True)[:3]
                lines.append((s, b["file"].split("/")[-1], r[0], int(r[ii] or 0), stalls, r[1].strip()[:90]))
    lines.sort(reverse=True)
    print("total samples", total)
    for s, f, ln, inst, stalls, src in lines[:top]:
        print("%5.1f%% %s:%s inst=%d %s | %s" % (100.0 * s / max(total, 1), f, ln, inst,
                                                  " ".join("%s=%d" % (c, v) for v, c in stalls if v), src))


if __name__ == "__main__":
    main()
