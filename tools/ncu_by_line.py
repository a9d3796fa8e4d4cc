#!/usr/bin/env python
"""DYNAMIC profile per source line: join the per-instruction counters of an ncu capture (`--page source --csv`: executed warp
instructions, stall samples per SASS address) with the line table of the SAME binary (`nvdisasm -gi -c`), so that instructions
inlined from csrc/dexr_kernels.cuh are attributed to their own lines (ncu's CUDA view only resolves the outermost file).

  cuobjdump -xelf all dex_retargeting_b200/libdexr.so && nvdisasm -gi -c dexr.*.cubin > all.txt     (the profiled build!)
  python tools/ncu_by_line.py prof.ncu-rep all.txt 'dexr_sequences_kernelILi16ELi0ELi8' [top]
"""
import collections
import csv
import io
import re
import subprocess
import sys


def line_table(path, kernel):
    lines = open(path).read().splitlines()
    start = next(i for i, ln in enumerate(lines) if ln.startswith(".text.") and kernel in ln)
    end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith(".text.")), len(lines))
    table, cur, fresh = {}, [], True
    for ln in lines[start:end]:
        m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
        if m:
            if fresh:
                cur, fresh = [], False
            cur.append((m.group(1).split("/")[-1], int(m.group(2))))
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+", ln)
        if m:
            fresh = True
            # innermost location first in the marker list: attribute to the innermost dexr_kernels.cuh line when there is one
            loc = next((c for c in cur if c[0] == "dexr_kernels.cuh"), cur[0] if cur else ("?", 0))
            table[int(m.group(1), 16)] = loc
    return table


def main():
    rep, dis, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
    table = line_table(dis, kernel)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    base = None
    ex, smp = collections.Counter(), collections.Counter()
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        addr = int(r[col["Address"]], 16)
        base = addr if base is None else base
        loc = table.get(addr - base, ("?", 0))
        ex[loc] += float(r[col["Instructions Executed"]] or 0)
        smp[loc] += float(r[col["# Samples"]] or 0)
    tot, ts = sum(ex.values()), sum(smp.values())
    print(f"total warp instructions {tot:.4g}, stall samples {ts:.0f}; by source line (inst %, samples %):")
    for loc, n in sorted(ex.items(), key=lambda kv: -kv[1])[:top]:
        print(f"{100 * n / tot:6.2f} {100 * smp[loc] / ts:6.2f}  {loc[0]}:{loc[1]}")
    return ex, smp


if __name__ == "__main__":
    main()
