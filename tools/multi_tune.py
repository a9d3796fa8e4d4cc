#!/usr/bin/env python
"""Mixed-robot call, joint CTA sizing: sweep the two tunables of dexr_solve_frames_multi (DEXR_MULTI_WAVES, DEXR_MULTI_SPREAD_AT,
both read per call) against the lone-launch sizing (DEXR_MULTI_SLOTS=spread) for six robots x n frames."""
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import workloads as W  # noqa: E402
from dex_retargeting_b200.optimizer import retarget_batch_mixed  # noqa: E402

dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
SETTINGS = [("spread", dict(DEXR_MULTI_SLOTS="spread"))] + \
           [(f"waves={w}", dict(DEXR_MULTI_WAVES=str(w), DEXR_MULTI_SPREAD_AT="100000")) for w in (1, 2, 4, 8, 16)] + \
           [("lpt+spread", dict(DEXR_MULTI_SPREAD_AT="1"))]
print("frames/robot  " + "  ".join(f"{n:>11s}" for n, _ in SETTINGS) + "   (ms per six-robot step)")
for n in (512, 1024, 2048, 4096, 8192, 16384):
    jobs = []
    for i, key in enumerate(W.MIXED_KEYS):
        seq = W.build(key, device=0)
        kp, x0, f, _ = W.frames(seq, n, W.MIXED_SEED + i)
        jobs.append((seq.optimizer, dict(keypoints=torch.from_numpy(kp).to(dev), last_qpos=torch.from_numpy(x0).to(dev),
                                        fixed_qpos=torch.from_numpy(f).to(dev) if f is not None else None,
                                        out=torch.empty((n, seq.optimizer.opt_dof), dtype=torch.float32, device=dev))))
    row = []
    for name, env in SETTINGS:
        os.environ.update(env)
        for _ in range(3):
            retarget_batch_mixed(jobs)
        torch.cuda.synchronize()
        ms = 0.0
        for _ in range(10):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); retarget_batch_mixed(jobs); b.record(); b.synchronize()
            ms += a.elapsed_time(b)
        for k in env:
            os.environ.pop(k)
        row.append(ms / 10)
    print(f"{n:12d}  " + "  ".join(f"{v:11.4f}" for v in row), flush=True)
