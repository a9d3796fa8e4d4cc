#!/usr/bin/env python
"""Position retargeting, initial damping 1e-3 (default since round 2) against 1e-2: config 3 (Shadow position, 65 536 frames) on one B200."""
import os
import sys

import torch

from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import workloads as W  # noqa: E402

dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
NARROW = (True, False) if "--both" in sys.argv else (True,)
cache = {}
for lam in ("default", "1e-2"):
    if lam == "default":
        os.environ.pop("DEXR_LAMBDA0", None)
    else:
        os.environ["DEXR_LAMBDA0"] = lam
    seq = W.build(W.SHADOW_POS_KEY, device=0)
    for narrow in NARROW:
        if narrow not in cache:
            cache[narrow] = W.frames(seq, 65536, W.SHADOW_SEED, narrow_dummy=narrow)
        kp, x0, f, _ = cache[narrow]
        k, x = torch.from_numpy(kp).to(dev), torch.from_numpy(x0).to(dev)
        st = torch.zeros((65536,), dtype=torch.int32, device=dev)
        out = torch.empty((65536, seq.optimizer.opt_dof), dtype=torch.float32, device=dev)
        for _ in range(2):
            seq.optimizer.retarget_batch(keypoints=k, last_qpos=x, out=out, status_out=st)
        torch.cuda.synchronize()
        ms = 0.0
        for _ in range(5):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); seq.optimizer.retarget_batch(keypoints=k, last_qpos=x, out=out, status_out=st); b.record(); b.synchronize()
            ms += a.elapsed_time(b)
        s = st.cpu().numpy()
        print(f"lambda0 {seq.optimizer.lambda0:g} dummy {'narrowed' if narrow else 'shipped '}: {ms / 5:.4f} ms  iterations {(s & 0xffff).mean():.3f}  "
              f"rejected trials {((s >> 16) & 0x7f).mean():.3f}  flagged {int(((s >> 24) != 0).sum())}", flush=True)
