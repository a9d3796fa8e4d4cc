#!/usr/bin/env python
"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck / synccheck):
frames kernel (G=16 dense, G=16 block, G=32, mimic), sequences kernel, preprocessing kernel, ragged batch sizes."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from helpers import build_product, keypoint_trajectory  # noqa: E402
from dex_retargeting_b200.preprocess import preprocess_keypoints  # noqa: E402

dev = torch.device("cuda", 0)
kp = keypoint_trajectory()
for key, B in (("teleop/allegro_hand_right", 37), ("teleop/leap_hand_right_dexpilot", 64), ("offline/shadow_hand_right", 21),
               ("teleop/schunk_svh_hand_right", 33), ("offline/panda_gripper", 9)):
    seq = build_product(key)
    opt = seq.optimizer
    k = torch.from_numpy(kp[:B * 3:3].copy()).to(dev)
    x0 = torch.from_numpy(np.tile(seq.joint_limits.mean(1).astype(np.float32), (B, 1))).to(dev)
    proj = torch.zeros((B, opt._objective_spec().len_proj), dtype=torch.uint8, device=dev) if opt.retargeting_type == "DEXPILOT" else None
    q = opt.retarget_batch(keypoints=k, last_qpos=x0, projected=proj, clip_init=True)
    torch.cuda.synchronize()
    print(key, "frames ok", float(q.abs().sum()))
for key in ("teleop/allegro_hand_right", "teleop/shadow_hand_right_dexpilot"):
    seq = build_product(key)
    tk = torch.from_numpy(np.stack([kp[0:16], kp[100:116], kp[300:316]]).copy()).to(dev)
    out, st = seq.retarget_sequences(tk)
    torch.cuda.synchronize()
    print(key, "sequences ok", float(out.abs().sum()))
raw = torch.from_numpy(kp[:300].copy()).to(dev) + 0.1
out = preprocess_keypoints(raw)
torch.cuda.synchronize()
print("preprocess ok", float(out.abs().sum()))
