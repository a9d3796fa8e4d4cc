#!/usr/bin/env python
"""Run the bodies of the frames-mode GPU parity tests (tests/test_gpu_parity.py) against the HOST EMULATION of the solver
source instead of the CUDA library -- for a chosen set of compile-time experiment switches.  A dry run on the CPU of what
`pytest -m gpu` will check on a B200 (minus GPU arithmetic in the last bits, minus the kernels of dexr.cu).

  python tests/tools/emu_gpu_tests.py                         # default build of the solver
  python tests/tools/emu_gpu_tests.py DEXR_EXP_PDFALLBACK DEXR_EXP_MERGEDRES
"""
import sys
import time
import traceback
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
for p in (ROOT, ROOT / "tests"):
    sys.path.insert(0, str(p))
import emu_host  # noqa: E402
import test_gpu_parity as T  # noqa: E402

DEFINES = tuple(sys.argv[1:])


def emu_solve(opt, refs=None, fixed=None, x0=None, keypoints=None, clip_init=False, want_proj=False, proj_init=None):
    B = x0.shape[0]
    proj = None
    if opt.retargeting_type == "DEXPILOT":
        lp = opt._objective_spec().len_proj
        proj = np.zeros((B, lp), np.uint8) if proj_init is None else np.ascontiguousarray(proj_init, dtype=np.uint8).copy()
    fx = fixed if (fixed is not None and fixed.shape[1] > 0) else None
    q, status, cost = emu_host.solve_frames(opt, x0, keypoints=keypoints, ref_value=None if keypoints is not None else refs,
                                           fixed_qpos=fx, projected=proj, defines=DEFINES, clip_init=clip_init)
    full = np.zeros((B, opt.robot.dof), np.float32)  # scatter + mimic like the kernel's optional robot_qpos output
    full[:, opt.idx_pin2target] = q
    if fx is not None:
        full[:, opt.idx_pin2fixed] = fx
    if opt.adaptor is not None:
        full = np.stack([opt.adaptor.forward_qpos(r.astype(np.float64)) for r in full]).astype(np.float32)
    res = dict(q=q, status=status, cost=cost, robot_qpos=full)
    if proj is not None:
        res["projected"] = proj
    return res


def main():
    T.gpu_solve = emu_solve
    jobs = [(T.test_synthetic_warm_start_parity, dict(key=k, ov=ov)) for k, ov in T.FAMILIES]
    jobs += [(T.test_recorded_trajectory_parity, dict(key=k)) for k in
             ["teleop/allegro_hand_right", "teleop/shadow_hand_right", "teleop/schunk_svh_hand_right", "teleop/leap_hand_right_dexpilot"]]
    for mark in getattr(T.test_reference_test_protocol, "pytestmark", []):
        if mark.name == "parametrize":
            jobs += [(T.test_reference_test_protocol, dict(zip(("key", "kind"), v))) for v in mark.args[1]]
    jobs += [(T.test_nonfinite_input_does_not_poison_neighbours, {}), (T.test_bounds_are_respected_and_active, {})]
    failed = 0
    for fn, kw in jobs:
        t0 = time.time()
        try:
            fn(**kw)
            print(f"PASS {fn.__name__} {kw} [{time.time() - t0:.1f}s]", flush=True)
        except Exception:
            failed += 1
            print(f"FAIL {fn.__name__} {kw}\n{traceback.format_exc(limit=3)}", flush=True)
    print(f"{len(jobs) - failed}/{len(jobs)} passed with defines {DEFINES or '(default)'}")
    return failed


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
