#!/usr/bin/env python
"""GPU box: how often does the CUDA path end in the same basin as SLSQP (reference run + oracle mode B)?  Summary lines."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from helpers import GOLDEN, build_oracle, build_product, synth_problems  # noqa: E402
from oracle.solvers import solve_converged  # noqa: E402

dev = torch.device("cuda", 0)


def solve(seq, o, refs, fixed, x0):
    opt = seq.optimizer
    B = refs.shape[0]
    status = torch.zeros(B, dtype=torch.int32, device=dev)
    cost = torch.zeros(B, dtype=torch.float32, device=dev)
    proj = torch.zeros((B, opt._objective_spec().len_proj), dtype=torch.uint8, device=dev) if o.type == "dexpilot" else None
    q = opt.retarget_batch(torch.from_numpy(refs).to(dev), torch.from_numpy(fixed).to(dev) if fixed.shape[1] else None,
                           torch.from_numpy(x0).to(dev), status_out=status, cost_out=cost, projected=proj)
    torch.cuda.synchronize()
    return q.cpu().numpy(), cost.cpu().numpy().astype(np.float64), status.cpu().numpy()


V = np.load(GOLDEN / "reference_vectors.npz")
tot_w = tot_b = tot_n = 0
for case in sorted({k.split("/")[0] for k in V.files if k.endswith("/retarget_cost")}):
    key = str(V[f"{case}/key"])
    seq, o = build_product(key), build_oracle(key)
    q, cost, st = solve(seq, o, V[f"{case}/ref_value"], V[f"{case}/fixed_qpos"], V[f"{case}/last_qpos"])
    rc = V[f"{case}/retarget_cost"]
    worse = cost > rc * (1 + 2e-5) + 1e-7
    better = cost < rc * (1 - 1e-2)
    print(f"ref {case}: worse-than-reference {worse.sum()}/{len(rc)} much-better {better.sum()} iters {np.mean(st & 0xffff):.1f} rej {np.mean((st >> 16) & 0xff):.1f}")
    tot_w += worse.sum(); tot_b += better.sum(); tot_n += len(rc)
print(f"TOTAL reference vectors: worse {tot_w} better {tot_b} of {tot_n}")
# cold starts in the style of the reference's test protocol (init = q* + 0.5 N) and pinch-perturbed dexpilot targets
for key, noise in [("teleop/allegro_hand_right", 0.5), ("teleop/shadow_hand_right", 0.5), ("teleop/leap_hand_right_dexpilot", 0.5),
                   ("teleop/ability_hand_right_dexpilot", 0.5), ("offline/shadow_hand_right", 0.3), ("teleop/schunk_svh_hand_right", 0.5)]:
    ov = {"scaling_factor": 1.0} if "offline" not in key else {}
    seq, o = build_product(key, ov), build_oracle(key, ov)
    rng = np.random.RandomState(7)
    n = 64
    refs, fixed, x0, _ = synth_problems(o, n, rng, init_noise=noise, target_noise=0.0)
    q, cost, st = solve(seq, o, refs, fixed, x0)
    fb = np.empty(n); dq = np.empty(n)
    for i in range(n):
        if o.type == "dexpilot":
            o.projected[:] = False
        xb, kkt, fb[i] = solve_converged(o, refs[i], fixed[i], x0[i], update_state=False)
        dq[i] = np.abs(xb - q[i]).max()
    print(f"cold {key} noise {noise}: same-basin {(dq < 1e-4).mean():.2f} gpu-worse {(cost > fb * (1 + 1e-4) + 1e-6).sum()} gpu-better {(cost < fb * (1 - 1e-4) - 1e-6).sum()} "
          f"of {n}; mean F gpu {cost.mean():.5f} B {fb.mean():.5f}; iters {np.mean(st & 0xffff):.1f} max {np.max(st & 0xffff)} rej {np.mean((st >> 16) & 0xff):.1f} flagged {(st >> 24 != 0).sum()}")
