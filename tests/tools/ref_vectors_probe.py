#!/usr/bin/env python
"""GPU box: per-problem comparison of the CUDA path with tests/golden/reference_vectors.npz and the oracle's mode B."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from helpers import GOLDEN, build_oracle, build_product  # noqa: E402
from oracle.solvers import solve_converged  # noqa: E402

V = np.load(GOLDEN / "reference_vectors.npz")
dev = torch.device("cuda", 0)
np.set_printoptions(linewidth=200, precision=5)
for case in sys.argv[1:] or ["leap_dexpilot", "ability_dexpilot_mimic"]:
    key = str(V[f"{case}/key"])
    seq, o = build_product(key), build_oracle(key)
    opt = seq.optimizer
    refs, fixed, x0 = V[f"{case}/ref_value"], V[f"{case}/fixed_qpos"], V[f"{case}/last_qpos"]
    B = refs.shape[0]
    status = torch.zeros(B, dtype=torch.int32, device=dev)
    cost = torch.zeros(B, dtype=torch.float32, device=dev)
    proj = torch.zeros((B, opt._objective_spec().len_proj), dtype=torch.uint8, device=dev) if o.type == "dexpilot" else None
    q = opt.retarget_batch(torch.from_numpy(refs).to(dev), torch.from_numpy(fixed).to(dev) if fixed.shape[1] else None,
                           torch.from_numpy(x0).to(dev), status_out=status, cost_out=cost, projected=proj)
    torch.cuda.synchronize()
    q, cost, status = q.cpu().numpy(), cost.cpu().numpy(), status.cpu().numpy()
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    np.save(ROOT / "gpurun_out" / f"probe_q_{case}.npy", q)
    for i in range(B):
        if o.type == "dexpilot":
            o.projected[:] = False
        xb, kkt, fb = solve_converged(o, refs[i], fixed[i], x0[i], update_state=False)
        obj = o.make_objective(refs[i], fixed[i], x0[i], update_state=False)
        print(f"{case}[{i}] F_gpu {cost[i]:.6f} (oracle at gpu x {obj.consistent(q[i].astype(float)):.6f}) F_ref {V[f'{case}/retarget_cost'][i]:.6f} "
              f"F_B {fb:.6f} |gpu-B| {np.abs(q[i]-xb).max():.2e} |gpu-ref| {np.abs(q[i]-V[f'{case}/retarget'][i]).max():.2e} "
              f"iters {status[i] & 0xffff} rej {(status[i] >> 16) & 0xff} flags {status[i] >> 24} proj {proj[i].cpu().numpy() if proj is not None else ''}")
