#!/usr/bin/env python
"""One launch of each non-headline kernel instantiation on its BASELINE workload, for `ncu -k regex:dexr_` captures:
Shadow position (arrow, 65536 frames), LEAP DexPilot independent frames (dense 16-lane, 65536), LEAP DexPilot streams
(256 x 300 = one GPU's shard of config 4, then 2048 x 60)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import workloads as W  # noqa: E402

which = sys.argv[1:] or ["shadow", "leapdp", "streams", "streams2048"]
dev = torch.device("cuda", 0)
if "shadow" in which:
    seq = W.build(W.SHADOW_POS_KEY, device=0)
    kp, x0, fixed, _ = W.frames(seq, 65536, W.SHADOW_SEED, narrow_dummy=True)
    for _ in range(2):
        seq.optimizer.retarget_batch(keypoints=torch.from_numpy(kp).to(dev), last_qpos=torch.from_numpy(x0).to(dev))
    torch.cuda.synchronize()
if "leapdp" in which:
    seq = W.build(W.LEAP_DEXPILOT_KEY, device=0)
    kp, x0, fixed, _ = W.frames(seq, 65536, W.SHADOW_SEED)
    for _ in range(2):
        seq.optimizer.retarget_batch(keypoints=torch.from_numpy(kp).to(dev), last_qpos=torch.from_numpy(x0).to(dev))
    torch.cuda.synchronize()
if "streams" in which:
    seq = W.build(W.LEAP_DEXPILOT_KEY, device=0)
    tk = torch.from_numpy(W.streams(2048, 300)[:256]).to(dev)
    seq.retarget_sequences(tk)
    seq.retarget_sequences(tk)
    torch.cuda.synchronize()
if "streams2048" in which:
    seq = W.build(W.LEAP_DEXPILOT_KEY, device=0)
    tk = torch.from_numpy(W.streams(2048, 300)).to(dev)
    seq.retarget_sequences(tk)
    seq.retarget_sequences(tk)
    torch.cuda.synchronize()
