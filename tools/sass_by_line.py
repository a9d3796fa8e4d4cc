#!/usr/bin/env python
"""Static SASS instruction count per source line of one kernel (inlining resolved to the line inside Solver::solve).

  cuobjdump -xelf all dex_retargeting_b200/libdexr.so          # -> dexr.sm_100a.cubin
  nvdisasm -gi -c dexr.sm_100a.cubin > all.txt
  python tools/sass_by_line.py all.txt 'dexr_frames_kernelILi16ELi4ELi15' 146

The third argument is the line of dexr.cu that calls solve() for that kernel (frames: 146, sequences: 224 at the
time of writing).  Instructions from the first WARPSYNC.COLLECTIVE onwards (the out-of-line non-converged shuffle
paths the compiler parks at the end of the function) are left out.  Multiply by trip counts by hand: this is code
size / loop-body size, not a profile.
"""
import collections
import re
import sys


def main(path, kernel, call_line):
    lines = open(path).read().splitlines()
    start = next(i for i, ln in enumerate(lines) if ln.startswith(".text.") and kernel in ln)
    end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith(".text.")), len(lines))
    stack, fresh = [], True
    by, ops, total = collections.Counter(), collections.defaultdict(collections.Counter), 0
    for ln in lines[start:end]:
        m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
        if m:
            if fresh:
                stack, fresh = [], False
            stack.append((m.group(1).split("/")[-1], int(m.group(2)), m.group(3).split("/")[-1] if m.group(3) else None, m.group(4)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", ln)
        if not m:
            continue
        fresh = True
        if m.group(2) == "WARPSYNC":
            break
        total += 1
        key = next((("solve", l) for f, l, inf, inl in stack if inl == call_line and f == "dexr_kernels.cuh"), None)
        if key is None:
            key = stack[-1][:2] if stack else ("?", 0)
        by[key] += 1
        ops[key][m.group(2)] += 1
    print("main-body instructions:", total)
    for k, v in sorted(by.items(), key=lambda kv: (kv[0][0] != "solve", kv[0][1])):
        if v >= 6:
            print(f"{k[0]}:{k[1]:<5d} {v:5d}  {dict(ops[k].most_common(4))}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3])
