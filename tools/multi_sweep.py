#!/usr/bin/env python
"""Mixed-robot launch (dexr_solve_frames_multi): persistent kernel vs fork-join launches vs one launch after the other, for
six robots x n frames each -- calibrates the size at which the library switches (DEXR_MULTI_MODE forces a mode; it is read
once per process, so every mode runs in its own subprocess)."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

if len(sys.argv) > 1 and sys.argv[1] == "--child":
    import torch

    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tools"))
    import workloads as W
    from dex_retargeting_b200.optimizer import retarget_batch_mixed

    dev = torch.device("cuda", 0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {}
    for n in (64, 256, 1024, 2048, 4096, 16384):
        jobs = []
        for i, key in enumerate(W.MIXED_KEYS):
            seq = W.build(key, device=0)
            kp, x0, f, _ = W.frames(seq, n, W.MIXED_SEED + i)
            jobs.append((seq.optimizer, dict(keypoints=torch.from_numpy(kp).to(dev), last_qpos=torch.from_numpy(x0).to(dev),
                                            fixed_qpos=torch.from_numpy(f).to(dev) if f is not None else None,
                                            out=torch.empty((n, seq.optimizer.opt_dof), dtype=torch.float32, device=dev))))

        def mixed():
            retarget_batch_mixed(jobs)

        def mixed_spread():
            os.environ["DEXR_MULTI_SLOTS"] = "spread"
            retarget_batch_mixed(jobs)
            os.environ.pop("DEXR_MULTI_SLOTS")

        def sequential():
            for o, kw in jobs:
                o.retarget_batch(**kw)

        res = {}
        for name, fn in (("mixed", mixed), ("mixed_spread", mixed_spread), ("sequential", sequential)):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            ms = 0.0
            for _ in range(10):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record(); b.synchronize()
                ms += a.elapsed_time(b)
            res[name] = ms / 10
        out[n] = res
    print(json.dumps(out))
else:
    rows = {}
    for mode in ("persistent", "streams"):
        env = dict(os.environ)
        if mode != "auto":
            env["DEXR_MULTI_MODE"] = mode
        r = subprocess.run([sys.executable, __file__, "--child"], env=env, capture_output=True, text=True)
        try:
            rows[mode] = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception:
            print(mode, "failed:", r.stderr[-500:])
    print("frames/robot   persistent   fork-join packed (default)   fork-join spread   one-after-the-other   (ms per six-robot step)")
    for n in ("64", "256", "1024", "2048", "4096", "16384"):
        p, s = (rows.get(m, {}).get(n, {}) for m in ("persistent", "streams"))
        print(f"{int(n):12d}   {p.get('mixed', float('nan')):10.4f} {s.get('mixed', float('nan')):28.4f} {s.get('mixed_spread', float('nan')):18.4f} {s.get('sequential', float('nan')):21.4f}")
