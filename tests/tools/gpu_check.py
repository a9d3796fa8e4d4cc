#!/usr/bin/env python
"""Bring-up script for the GPU box: solve a few configs on the GPU, compare with the oracle, print stats."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from helpers import build_oracle, build_product, keypoint_trajectory, synth_problems  # noqa: E402
from oracle.solvers import solve_converged  # noqa: E402

KEYS = sys.argv[1:] or ["teleop/allegro_hand_right", "offline/shadow_hand_right", "teleop/leap_hand_right_dexpilot",
                        "teleop/schunk_svh_hand_right", "teleop/shadow_hand_right_dexpilot", "offline/schunk_svh_hand_right",
                        "teleop/panda_gripper", "teleop/ability_hand_right_dexpilot"]


def main():
    dev = torch.device("cuda", 0)
    print(torch.cuda.get_device_name(0))
    for key in KEYS:
        ov = {"scaling_factor": 1.0} if "offline" not in key else {}
        seq = build_product(key, ov)
        opt = seq.optimizer
        o = build_oracle(key, ov)
        rng = np.random.RandomState(1)
        n = 48
        refs, fixed, x0, _ = synth_problems(o, n, rng, init_noise=0.05, target_noise=0.01)
        XB = []
        for i in range(n):
            if o.type == "dexpilot":
                o.projected[:] = False
            xb, kkt, Fb = solve_converged(o, refs[i], fixed[i], x0[i], update_state=False)
            XB.append(xb)
        XB = np.array(XB)
        t_ref = torch.from_numpy(refs).to(dev)
        t_x0 = torch.from_numpy(x0).to(dev)
        t_fixed = torch.from_numpy(fixed).to(dev) if fixed.shape[1] else None
        status = torch.zeros(n, dtype=torch.int32, device=dev)
        cost = torch.zeros(n, dtype=torch.float32, device=dev)
        proj = torch.zeros((n, opt._objective_spec().len_proj), dtype=torch.uint8, device=dev) if o.type == "dexpilot" else None
        q = opt.retarget_batch(t_ref, t_fixed, t_x0, status_out=status, cost_out=cost, projected=proj)
        torch.cuda.synchronize()
        q = q.cpu().numpy()
        st = status.cpu().numpy()
        dq = np.abs(q - XB).max(1)
        print(f"{key}: n={opt.opt_dof} iters mean {np.mean(st & 0xffff):.1f} max {np.max(st & 0xffff)} "
              f"rej mean {np.mean((st >> 16) & 0xff):.1f} flags {np.unique(st >> 24)} |dq| median {np.median(dq):.2e} "
              f"p90 {np.percentile(dq, 90):.2e} max {dq.max():.2e} frac<1e-4 {(dq < 1e-4).mean():.2f}", flush=True)
        # throughput
        B = 65536
        reps = (B + n - 1) // n
        big_ref = t_ref.repeat(reps, 1, 1)[:B].contiguous()
        big_x0 = t_x0.repeat(reps, 1)[:B].contiguous()
        big_fixed = t_fixed.repeat(reps, 1)[:B].contiguous() if t_fixed is not None else None
        big_proj = torch.zeros((B, proj.shape[1]), dtype=torch.uint8, device=dev) if proj is not None else None
        out = torch.empty((B, opt.opt_dof), dtype=torch.float32, device=dev)
        for _ in range(2):
            opt.retarget_batch(big_ref, big_fixed, big_x0, out=out, projected=big_proj)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            opt.retarget_batch(big_ref, big_fixed, big_x0, out=out, projected=big_proj)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"    B={B}: {ms:.3f} ms/launch -> {B / ms * 1e3:.3e} frames/s  {opt.engine().launch_info()}", flush=True)


if __name__ == "__main__":
    main()
