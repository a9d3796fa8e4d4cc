#!/usr/bin/env python
"""Executed warp instructions and stall samples per CUDA source line of a captured kernel
(`ncu -i rep --page source --print-source cuda --csv`; needs -lineinfo and --import-source on), grouped into the phases of the
solver by line ranges of csrc/dexr_kernels.cuh given on the command line.

  python tools/ncu_lines.py rep.ncu-rep [top]
"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    # the report holds one block per source file: "File Path" rows, then a header row, then lines
    cur, hdr, col, out = None, None, None, []
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1]
            hdr = None
            continue
        if r and r[0] in ("#", "Line #", "Line"):
            hdr = r
            col = {h: i for i, h in enumerate(hdr)}
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        try:
            n = float(r[col["Instructions Executed"]] or 0)
            s = float(r[col["# Samples"]] or 0)
        except (ValueError, KeyError):
            continue
        out.append((n, s, cur, r[0], r[col.get("Source", 1)].strip()[:110]))
    tot = sum(o[0] for o in out) or 1
    ts = sum(o[1] for o in out) or 1
    print(f"total warp instructions {tot:.4g}, samples {ts:.0f}")
    for n, s, f, ln, src in sorted(out, reverse=True)[:top]:
        print(f"{100 * n / tot:6.2f}% inst {100 * s / ts:6.2f}% smp  {str(f).split('/')[-1]}:{ln}  {src}")


if __name__ == "__main__":
    main()
