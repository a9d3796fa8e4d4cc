#!/usr/bin/env python
"""GPU box: the reference's test protocol (tests/test_optimizer.py) on one config; prints task errors, iterations, flags."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from helpers import build_oracle, build_product  # noqa: E402
from oracle.solvers import generate_problem  # noqa: E402

dev = torch.device("cuda", 0)
np.set_printoptions(linewidth=220, precision=4, suppress=True)
key = sys.argv[1] if len(sys.argv) > 1 else "offline/shadow_hand_right"
kind = "position" if "offline" in key else "vector"
ov = dict(normal_delta=0) if kind == "position" else dict(low_pass_alpha=0, scaling_factor=1.0, normal_delta=0)
seq, o = build_product(key, ov), build_oracle(key, ov)
opt = seq.optimizer
np.random.seed(1)
n = 100
refs, fixed, x0 = [], [], []
for _ in range(n):
    q, init, target = generate_problem(o)
    refs.append(target.astype(np.float32)); fixed.append(q[o.idx_pin2fixed]); x0.append(init[o.idx_pin2target])
refs, fixed, x0 = np.array(refs), np.array(fixed).reshape(n, -1), np.array(x0, dtype=np.float32)
status = torch.zeros(n, dtype=torch.int32, device=dev)
cost = torch.zeros(n, dtype=torch.float32, device=dev)
prm = {}
q = opt.retarget_batch(torch.from_numpy(refs).to(dev), torch.from_numpy(fixed.astype(np.float32)).to(dev) if fixed.shape[1] else None,
                       torch.from_numpy(x0).to(dev), status_out=status, cost_out=cost, clip_init=(kind == "position"))
torch.cuda.synchronize()
q, cost, st = q.cpu().numpy(), cost.cpu().numpy(), status.cpu().numpy()
errs = np.array([o.make_objective(refs[i], fixed[i], x0[i], update_state=False).task_error(q[i].astype(float)) for i in range(n)])
print(key, "mean err", errs.mean(), "median", np.median(errs), "max", errs.max(), "n>1e-2", (errs > 1e-2).sum(),
      "iters mean", (st & 0xffff).mean(), "max", (st & 0xffff).max(), "flagged", (st >> 24 != 0).sum())
bad = np.argsort(-errs)[:5]
for i in bad:
    print(i, "err", errs[i], "cost", cost[i], "iters", st[i] & 0xffff, "rej", (st[i] >> 16) & 0xff, "flag", st[i] >> 24)
    print("   x0", x0[i][:8]); print("   q ", q[i][:8])
