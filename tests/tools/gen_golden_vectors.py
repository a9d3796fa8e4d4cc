#!/usr/bin/env python
"""Generate tests/golden/oracle_vectors.npz: seeded inputs and the oracle's outputs for a set of configurations.

The reference itself cannot run offline (pinocchio / nlopt absent), so these are the ORACLE's answers, not the
reference's: mode B (converged minimiser, the joint-space parity target) and mode A (reference-faithful early-stopped
SLSQP) for the same inputs.  They pin the oracle against regressions (tests/test_golden_vectors.py, CPU) and give
the GPU suite fixed expected outputs that need no CPU solve at test time.

Usage: python tests/tools/gen_golden_vectors.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from helpers import build_oracle, keypoint_trajectory, synth_problems  # noqa: E402
from oracle.solvers import solve_converged, solve_reference  # noqa: E402

CASES = [  # (name, config key, override, n, init noise, target noise, seed)
    ("allegro_vector", "teleop/allegro_hand_right", {}, 32, 0.05, 0.01, 101),
    ("shadow_position_dummy", "offline/shadow_hand_right", {}, 16, 0.05, 0.005, 102),
    ("leap_dexpilot", "teleop/leap_hand_right_dexpilot", {}, 24, 0.05, 0.01, 103),
    ("svh_vector_mimic", "teleop/schunk_svh_hand_right", {}, 24, 0.05, 0.01, 104),
    ("inspire_position_mimic_dummy", "offline/inspire_hand_right", {}, 16, 0.05, 0.005, 105),
    ("panda_vector_prismatic", "teleop/panda_gripper", {}, 16, 0.01, 0.002, 106),
    ("shadow_dexpilot_5finger", "teleop/shadow_hand_right_dexpilot", {}, 16, 0.05, 0.01, 107),
]


def main():
    out = {}
    for name, key, ov, n, noise, tnoise, seed in CASES:
        o = build_oracle(key, ov)
        rng = np.random.RandomState(seed)
        refs, fixed, x0, _ = synth_problems(o, n, rng, init_noise=noise, target_noise=tnoise)
        XB, FB, XA = [], [], []
        for i in range(n):
            if o.type == "dexpilot":
                o.projected[:] = False
            xb, kkt, fb = solve_converged(o, refs[i], fixed[i], x0[i], update_state=False)
            assert kkt < 1e-8, (name, i, kkt)
            if o.type == "dexpilot":
                o.projected[:] = False
            xa, _ = solve_reference(o, refs[i], fixed[i], x0[i])
            XB.append(xb); FB.append(fb); XA.append(xa)
        out[f"{name}/key"] = np.array(key)
        out[f"{name}/ref_value"] = refs
        out[f"{name}/fixed_qpos"] = fixed
        out[f"{name}/last_qpos"] = x0
        out[f"{name}/qpos_converged"] = np.array(XB)
        out[f"{name}/cost_converged"] = np.array(FB)
        out[f"{name}/qpos_reference_mode"] = np.array(XA)
        print(name, "done", np.abs(np.array(XB) - np.array(XA)).max())
    # a short stream of the recorded trajectory through the SeqRetargeting recurrence (mode B)
    from oracle.solvers import OracleSeqRetargeting

    kp = keypoint_trajectory()[0:120:4]
    oseq = OracleSeqRetargeting(build_oracle("teleop/allegro_hand_right"), mode="converged")
    traj = np.array([oseq.retarget(oseq.opt.ref_from_keypoints(k)) for k in kp])
    out["allegro_stream/keypoints"] = kp
    out["allegro_stream/robot_qpos"] = traj
    np.savez_compressed(ROOT / "tests" / "golden" / "oracle_vectors.npz", **out)


if __name__ == "__main__":
    main()
