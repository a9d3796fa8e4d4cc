#!/usr/bin/env python
"""Find frames the solver flags (max_iters / non-finite) on the DexPilot LEAP stream workload and compare them with the oracle."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from helpers import build_oracle, build_product, keypoint_trajectory  # noqa: E402
from oracle.solvers import solve_converged  # noqa: E402


def main():
    key = sys.argv[1] if len(sys.argv) > 1 else "teleop/leap_hand_right_dexpilot"
    dev = torch.device("cuda", 0)
    seq = build_product(key)
    opt = seq.optimizer
    o = build_oracle(key)
    S, T = 64, 300
    rng = np.random.RandomState(7)
    base = keypoint_trajectory()[:T].astype(np.float32)
    kp = base[None] + rng.randn(S, 1, 21, 3).astype(np.float32) * 0.002
    kp[:, :, 0] = 0
    tk = torch.from_numpy(np.ascontiguousarray(kp)).to(dev)
    st = seq.make_stream_state(S)
    flagged = []
    for t in range(T):
        last_in = st.last_qpos.clone()
        proj_in = st.projected.clone() if st.projected is not None else None
        status = torch.zeros(S, dtype=torch.int32, device=dev)
        cost = torch.zeros(S, dtype=torch.float32, device=dev)
        q = opt.retarget_batch(keypoints=tk[:, t].contiguous(), last_qpos=st.last_qpos, projected=st.projected, clip_init=True,
                               status_out=status, cost_out=cost)
        torch.cuda.synchronize()
        sw = status.cpu().numpy()
        for s in np.nonzero(sw >> 24)[0]:
            flagged.append((t, int(s), int(sw[s]), float(cost[s]), last_in[s].cpu().numpy(), q[s].cpu().numpy(),
                            proj_in[s].cpu().numpy() if proj_in is not None else None))
        st.last_qpos = q
    print(f"{key}: {len(flagged)} flagged of {S * T}")
    for t, s, sw, c, last, q, proj in flagged[:12]:
        ref = o.ref_from_keypoints(kp[s, t]).astype(np.float32)
        lastc = np.clip(last, o.joint_limits[:, 0], o.joint_limits[:, 1])
        if proj is not None:
            o.projected[:] = proj.astype(bool)
        xb, kkt, Fb = solve_converged(o, ref, np.zeros(0), lastc, update_state=True)
        obj = o.make_objective(ref, np.zeros(0), lastc, update_state=False)
        if proj is not None:
            o.projected[:] = proj.astype(bool)
            obj = o.make_objective(ref, np.zeros(0), lastc, update_state=True)
        print(f"  t={t} s={s} status iters={sw & 0xffff} rej={(sw >> 16) & 0xff} flags={sw >> 24} cost={c:.6e} "
              f"F(gpu)={obj.consistent(q.astype(np.float64)):.6e} F(oracle)={Fb:.6e} dq={np.abs(q - xb).max():.2e} "
              f"at-bounds={int(((q <= o.lower + 1e-6) | (q >= o.upper - 1e-6)).sum())}")


if __name__ == "__main__":
    main()
