"""Shared test helpers: build the product objects and the oracle from the committed fixtures."""
import json
from functools import lru_cache
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"
ROBOTS = GOLDEN / "robots"


@lru_cache(maxsize=None)
def configs():
    with open(GOLDEN / "configs.json") as f:
        return json.load(f)


def config_dict(key, override=None):
    cfg = dict(configs()[key])
    cfg["urdf_path"] = Path(cfg["urdf_path"]).stem + ".json"  # fixtures hold the joint trees as JSON
    if override:
        cfg.update(override)
    return cfg


def build_product(key, override=None, device=None):
    from dex_retargeting_b200.retargeting_config import RetargetingConfig

    RetargetingConfig.set_default_urdf_dir(str(ROBOTS))
    return RetargetingConfig.from_dict(config_dict(key), override).build(device=device)


def build_oracle(key, override=None):
    from oracle.objectives import OracleOptimizer

    return OracleOptimizer(configs()[key], str(ROBOTS), override)


def keypoint_trajectory():
    return np.load(GOLDEN / "human_joint_right.npy")


def synth_problems(o, n, rng, init_noise=0.05, target_noise=0.0):
    """Problems in the style of tests/test_optimizer.py:27-81 of the reference (reachable targets from a
    random pose, noisy warm start), optionally with unreachable targets (target_noise, metres)."""
    lim = o.robot.joint_limits
    refs, fixed, x0, qstar = [], [], [], []
    for _ in range(n):
        q = rng.uniform(lim[:, 0], lim[:, 1])
        if o.adaptor is not None:
            q = o.adaptor.forward_qpos(q)
        init = np.clip(q + rng.randn(o.robot.dof) * init_noise, lim[:, 0] + 1e-5, lim[:, 1] - 1e-5)
        o.robot.compute_forward_kinematics(q)
        pos = o.robot.link_positions(o.link_ids)
        ref = pos if o.type == "position" else pos[o.task_sel] - pos[o.origin_sel]
        ref = ref + rng.randn(*ref.shape) * target_noise
        refs.append(ref.astype(np.float32))
        fixed.append(q[o.idx_pin2fixed].astype(np.float32))
        x0.append(init[o.idx_pin2target].astype(np.float32))
        qstar.append(q)
    return np.array(refs), np.array(fixed).reshape(n, -1), np.array(x0), np.array(qstar)
