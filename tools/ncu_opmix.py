#!/usr/bin/env python
"""Instruction mix of a captured kernel from the ncu source page (`ncu -i rep --page source --csv`): executed warp instructions
and stall samples aggregated by SASS opcode, plus the shared-memory wavefront excess (bank conflicts) by instruction.

  python tools/ncu_opmix.py gpurun_out/job/prof.ncu-rep [top]
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    ex, st, exc, wav = defaultdict(float), defaultdict(float), [], 0.0
    total = samples = 0.0
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        src = r[col["Source"]].strip()
        toks = src.split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        op = op.split(".")[0]
        n = float(r[col["Instructions Executed"]] or 0)
        s = float(r[col["# Samples"]] or 0)
        ex[op] += n
        st[op] += s
        total += n
        samples += s
        e = float(r[col["L1 Wavefronts Shared Excessive"]] or 0)
        w = float(r[col["L1 Wavefronts Shared"]] or 0)
        wav += w
        if e > 0:
            exc.append((e, w, src))
    print(f"warp instructions {total:.4g}, stall samples {samples:.0f}")
    print(f"{'opcode':10s} {'inst %':>7s} {'samples %':>9s}")
    for op, n in sorted(ex.items(), key=lambda kv: -kv[1])[:top]:
        print(f"{op:10s} {100 * n / total:7.2f} {100 * st[op] / max(samples, 1):9.2f}")
    print(f"\nshared wavefronts {wav:.4g}; excess (bank conflicts) by instruction:")
    for e, w, src in sorted(exc, reverse=True)[:12]:
        print(f"  {e:10.0f} of {w:10.0f}  {src}")


if __name__ == "__main__":
    main()
