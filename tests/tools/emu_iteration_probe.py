#!/usr/bin/env python
"""Iteration / rejection counts of the REAL solver source (host emulation, tests/emu) on the bench workloads, default build
against the compile-time experiments.  CPU only, slow (a fraction of a second per frame): small samples.

  python tests/tools/emu_iteration_probe.py offline/shadow_hand_right 48
"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
for p in (ROOT, ROOT / "tests", ROOT / "tools"):
    sys.path.insert(0, str(p))
import emu_host  # noqa: E402
from bench_configs import synth  # noqa: E402
from helpers import build_product  # noqa: E402

VARIANTS = [()]


def main(key, B, tol=None):
    seq = build_product(key)
    opt = seq.optimizer
    if tol:
        opt.step_tol = float(tol)
    kp, x0, fixed, _ = synth(seq, B, 100)
    base = None
    for d in VARIANTS:
        t0 = time.time()
        proj = np.zeros((B, opt._objective_spec().len_proj), np.uint8) if opt.retargeting_type == "DEXPILOT" else None
        lib = emu_host.load(tuple(d))
        r0 = lib.emu_rounds()
        q, st, cost = emu_host.solve_frames(opt, x0, keypoints=kp, fixed_qpos=fixed, projected=proj, defines=d)
        rounds = (lib.emu_rounds() - r0) / B  # warp collectives (shuffles, ballots, __syncwarp) per frame: a latency proxy
        it, rej = st & 0xffff, (st >> 16) & 0xff
        base = q if base is None else base
        print(f"{'+'.join(x.replace('DEXR_EXP_', '').lower() for x in d) or 'default':32s} iterations {it.mean():.3f} (max {it.max()}) "
              f"extra trial solves {rej.mean():.3f}  collectives/frame {rounds:.0f}  flagged {(st >> 24 != 0).sum()}  max |dq| vs default {np.abs(q - base).max():.2e}  "
              f"[{time.time() - t0:.1f}s]", flush=True)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), *sys.argv[3:])
