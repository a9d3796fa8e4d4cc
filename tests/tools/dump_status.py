#!/usr/bin/env python
"""Run the BASELINE workloads (tools/workloads.py, first input set) on cuda:0 and save status words + results, so that frames
the solver flags (max iterations) or that take unusually many iterations can be replayed on the CPU (host emulation of the solver
source, oracle) from their indices -- the inputs are seed-deterministic.

  python tests/tools/dump_status.py gpurun_out/status_default.npz            # DEXR_LIBRARY selects the build
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import workloads as W  # noqa: E402


def main(out):
    dev = torch.device("cuda", 0)
    res = {}

    def frames(tag, key, n, seed, **kw):
        seq = W.build(key, device=0)
        opt = seq.optimizer
        kp, x0, fixed, info = W.frames(seq, n, seed, **kw)
        q = torch.empty((n, opt.opt_dof), dtype=torch.float32, device=dev)
        st = torch.zeros((n,), dtype=torch.int32, device=dev)
        cost = torch.zeros((n,), dtype=torch.float32, device=dev)
        proj = torch.zeros((n, opt._objective_spec().len_proj), dtype=torch.uint8, device=dev) if opt.retargeting_type == "DEXPILOT" else None
        opt.retarget_batch(keypoints=torch.from_numpy(kp).to(dev), last_qpos=torch.from_numpy(x0).to(dev),
                           fixed_qpos=torch.from_numpy(fixed).to(dev) if fixed is not None else None, out=q, status_out=st,
                           cost_out=cost, projected=proj)
        torch.cuda.synchronize()
        s = st.cpu().numpy()
        res[f"{tag}/status"] = s  # (results are regenerated on demand: 65 536 x n floats per workload would not fit gpurun_out)
        bad = np.nonzero((s >> 24) != 0)[0]
        res[f"{tag}/flagged_index"], res[f"{tag}/flagged_q"], res[f"{tag}/flagged_cost"] = bad, q.cpu().numpy()[bad], cost.cpu().numpy()[bad]
        it = s & 0xffff
        print(f"{tag}: iterations mean {it.mean():.3f} p99 {np.percentile(it, 99):.0f} max {it.max()}  flagged {(s >> 24 != 0).sum()}  "
              f"rejects mean {((s >> 16) & 0x7f).mean():.3f}", flush=True)

    frames("allegro_vector", W.METRIC_KEY, 65536, W.METRIC_SEED)
    frames("allegro_vector_cold", W.METRIC_KEY, 65536, W.METRIC_SEED, sigma=0.5)
    frames("shadow_position_narrow", W.SHADOW_POS_KEY, 65536, W.SHADOW_SEED, narrow_dummy=True)
    frames("shadow_position_shipped", W.SHADOW_POS_KEY, 65536, W.SHADOW_SEED, narrow_dummy=False)
    frames("leap_dexpilot_frames", W.LEAP_DEXPILOT_KEY, 65536, W.SHADOW_SEED)
    seq = W.build(W.LEAP_DEXPILOT_KEY, device=0)
    kp = W.streams(2048, 300)
    o = torch.empty((2048, 300, seq.optimizer.robot.dof), dtype=torch.float32, device=dev)
    st = torch.zeros((2048, 300), dtype=torch.int32, device=dev)
    seq.retarget_sequences(torch.from_numpy(kp).to(dev), out=o, status_out=st)
    torch.cuda.synchronize()
    s = st.cpu().numpy()
    res["leap_dexpilot_streams/status"] = s
    it = s & 0xffff
    print(f"leap_dexpilot_streams: iterations mean {it.mean():.3f} max {it.max()} flagged {(s >> 24 != 0).sum()}", flush=True)
    np.savez_compressed(out, **res)


if __name__ == "__main__":
    main(sys.argv[1])
