#!/usr/bin/env python
"""Scan a BASELINE workload through the HOST EMULATION of the solver source on all host cores: iteration statistics, flagged
frames (max iterations) and, optionally, parity against the oracle fixture -- the CPU-side loop for solver-behaviour changes
(the emulated solver takes the same decisions as the kernel: same iteration counts as the B200 on 65 536 frames).

  python tests/tools/emu_scan.py leapdp 65536 [DEXR_EXP_...]     # workloads: metric, cold, real, shadow, shadowship, leapdp
  python tests/tools/emu_scan.py streams 64                       # 64 of the config-4 streams x 300 frames
"""
import multiprocessing as mp
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
for p in (ROOT, ROOT / "tests", ROOT / "tools"):
    sys.path.insert(0, str(p))
import emu_host  # noqa: E402
import workloads as W  # noqa: E402

CASES = {"metric": (W.METRIC_KEY, W.METRIC_SEED, {}, "metric"), "cold": (W.METRIC_KEY, W.METRIC_SEED, dict(sigma=0.5), "metric_cold"),
         "shadow": (W.SHADOW_POS_KEY, W.SHADOW_SEED, dict(narrow_dummy=True), "shadow_narrow"),
         "shadowship": (W.SHADOW_POS_KEY, W.SHADOW_SEED, dict(narrow_dummy=False), "shadow_ship"),
         "leapdp": (W.LEAP_DEXPILOT_KEY, W.SHADOW_SEED, {}, "leap_frames"), "real": (W.METRIC_KEY, 0, {}, "metric_real")}


def _frames(args):
    key, kp, x0, fixed, defines = args
    seq = W.build(key)
    proj = np.zeros((kp.shape[0], seq.optimizer._objective_spec().len_proj), np.uint8) if seq.optimizer.retargeting_type == "DEXPILOT" else None
    return emu_host.solve_frames(seq.optimizer, x0, keypoints=kp, fixed_qpos=fixed, projected=proj, defines=defines)


def _streams(args):
    key, kp, defines = args
    seq = W.build(key)
    got, status, _ = emu_host.solve_sequences(seq, kp, defines=defines)
    return got, status


def main():
    what, n = sys.argv[1], int(sys.argv[2])
    defines = tuple(a for a in sys.argv[3:] if a.startswith("DEXR_"))
    emu_host.load(defines)  # build once, before forking
    t0 = time.time()
    procs = 8
    with mp.get_context("fork").Pool(procs) as pool:
        if what == "streams":
            kp = W.streams(2048, 300)[:n]
            parts = pool.map(_streams, [(W.LEAP_DEXPILOT_KEY, kp[i::procs], defines) for i in range(procs)])
            status = np.zeros((n, 300), np.int32)
            q = np.zeros((n, 300, 16), np.float32)
            for i, (g, s) in enumerate(parts):
                status[i::procs], q[i::procs] = s, g
            tag = "leap_streams"
        else:
            key, seed, kw, tag = CASES[what]
            seq = W.build(key)
            gen_n = 65536
            if what == "real":
                kp, x0 = W.real_frames(seq, gen_n); fixed = None
            else:
                kp, x0, fixed, _ = W.frames(seq, gen_n, seed, **kw)
            chunks = np.array_split(np.arange(n), procs * 4)
            parts = pool.map(_frames, [(key, kp[c], x0[c], None if fixed is None else fixed[c], defines) for c in chunks])
            q = np.concatenate([p[0] for p in parts]); status = np.concatenate([p[1] for p in parts])
    it = status & 0xffff
    rej = (status >> 16) & 0x7f
    bad = np.argwhere((status >> 24) != 0)
    print(f"{what} n={n} defines={defines}: iterations mean {it.mean():.3f} p99 {np.percentile(it, 99):.0f} max {it.max()}  rejects mean {rej.mean():.3f}  "
          f"flagged {len(bad)}  ({time.time() - t0:.0f} s)")
    print("flagged:", bad[:40].tolist())
    hist = np.bincount(it.reshape(-1), minlength=65)
    print("iteration histogram (>=12):", {i: int(c) for i, c in enumerate(hist) if c and i >= 12})
    if tag is not None:
        import parity as P

        nf = int(P.fixture()[f"{tag}/n"]) // (300 if what == "streams" else 1)
        print("parity:", P.compare(tag, q[:nf], None, status[:nf]))
    out = Path("/tmp") / f"emu_scan_{what}_{'_'.join(d[9:].lower() for d in defines) or 'default'}.npz"
    np.savez_compressed(out, q=q, status=status)
    print("saved", out)


if __name__ == "__main__":
    main()
