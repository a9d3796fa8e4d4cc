#!/usr/bin/env python
"""Single-frame latency of the reference-style call SeqRetargeting.retarget() (numpy in / numpy out, B = 1)."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from helpers import build_oracle, build_product, keypoint_trajectory  # noqa: E402

kp = keypoint_trajectory()
for key in ("teleop/allegro_hand_right", "teleop/shadow_hand_right", "teleop/leap_hand_right_dexpilot", "offline/shadow_hand_right"):
    seq = build_product(key)
    o = build_oracle(key)
    refs = [o.ref_from_keypoints(k) for k in kp[:300]]
    for r in refs[:20]:
        seq.retarget(r)
    t0 = time.perf_counter()
    for r in refs[20:]:
        seq.retarget(r)
    dt = (time.perf_counter() - t0) / 280
    print(f"{key}: {dt * 1e6:.0f} us per retarget() call -> {1 / dt:.0f} Hz")
