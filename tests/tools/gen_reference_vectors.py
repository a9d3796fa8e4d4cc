#!/usr/bin/env python
"""Generate tests/golden/reference_vectors.npz by EXECUTING THE REFERENCE'S OWN PYTHON for the hot path.

Build container only (needs /root/reference; the fixture travels, the reference does not).

What runs unmodified from /root/reference/src/dex_retargeting: optimizer.py (Optimizer.retarget and the three
get_objective_function closures: torch SmoothL1 loss, gradient assembly with the grad-only regulariser, the DexPilot
hysteresis / weights / projected targets), kinematics_adaptor.py (MimicJointKinematicAdaptor.forward_qpos /
backward_jacobian), seq_retarget.py (SeqRetargeting.__init__ / retarget), optimizer_utils.py (LPFilter).

What is absent offline and therefore shimmed -- this is all that stays unpinned:
  * pinocchio: robot_wrapper.py imports it, so an empty stand-in module is registered; the optimizers only need an
    object with RobotWrapper's methods (robot_wrapper.py:28-95), here `DuckRobot` over the oracle's pure-Python 4x4 FK
    (oracle/robot.py with use_c=False), returning the LOCAL-frame Jacobian exactly as computeFrameJacobian would so the
    reference's own `link_rot @ J[:3]` (optimizer.py:172-177) is exercised;
  * nlopt: `ShimOpt` implements the six calls optimizer.py makes (:41,59,60,96,98,136) on top of scipy's SLSQP;
  * pytransform3d: only SeqRetargeting.warm_start uses it; an empty module is registered.

Recorded per case: inputs, the reference closure's value and gradient at several points, the DexPilot `projected`
state it leaves behind, `Optimizer.retarget` results (through ShimOpt), and a SeqRetargeting stream over the recorded
keypoint trajectory.  tests/test_reference_vectors.py holds the oracle (and tests/test_gpu_golden.py the CUDA path) to
these numbers.

Usage: python tests/tools/gen_reference_vectors.py [/root/reference]
"""
import sys
import types
from pathlib import Path

import numpy as np
from scipy.optimize import minimize

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from helpers import ROBOTS, configs, keypoint_trajectory, synth_problems, build_oracle  # noqa: E402
from oracle.robot import OracleRobot  # noqa: E402


class ShimOpt:
    """The nlopt.opt surface optimizer.py uses, over scipy.optimize SLSQP (same Kraft code base)."""

    def __init__(self, algorithm, n):
        self.n, self.lb, self.ub, self.ftol, self.f = n, [-np.inf] * n, [np.inf] * n, 0.0, None

    def set_lower_bounds(self, lb):
        self.lb = list(lb)

    def set_upper_bounds(self, ub):
        self.ub = list(ub)

    def set_ftol_abs(self, tol):
        self.ftol = tol

    def set_min_objective(self, f):
        self.f = f

    def optimize(self, x0):
        def fun(x):
            g = np.zeros(self.n)
            return float(self.f(x, g)), g

        res = minimize(fun, np.asarray(x0, dtype=np.float64), jac=True, method="SLSQP", bounds=list(zip(self.lb, self.ub)),
                       options=dict(ftol=self.ftol, maxiter=1000))
        return res.x


def install_shims(ref_root):
    pin = types.ModuleType("pinocchio")
    pin.Model = pin.Data = pin.SE3 = object
    pin.BODY = 2
    nlopt = types.ModuleType("nlopt")
    nlopt.LD_SLSQP, nlopt.opt = 40, ShimOpt
    pt = types.ModuleType("pytransform3d")
    pt.rotations = types.ModuleType("pytransform3d.rotations")
    sys.modules.update({"pinocchio": pin, "nlopt": nlopt, "pytransform3d": pt, "pytransform3d.rotations": pt.rotations})
    sys.path.insert(0, str(Path(ref_root) / "src"))


class DuckRobot:
    """RobotWrapper's method surface (robot_wrapper.py:28-95) over the oracle's pure-Python kinematics."""

    def __init__(self, desc, add_dummy):
        self.r = OracleRobot(desc, add_dummy, use_c=False)

    dof = property(lambda s: s.r.dof)
    dof_joint_names = property(lambda s: list(s.r.dof_joint_names))
    joint_names = property(lambda s: ["universe"] + list(s.r.dof_joint_names))
    link_names = property(lambda s: list(s.r.link_names))
    joint_limits = property(lambda s: s.r.joint_limits.copy())

    def get_joint_index(self, name):
        return self.dof_joint_names.index(name)

    def get_link_index(self, name):
        return self.r.get_link_index(name)

    def compute_forward_kinematics(self, qpos):
        self.r.compute_forward_kinematics(np.asarray(qpos, dtype=np.float64))

    def get_link_pose(self, link_id):
        return self.r.get_link_pose(link_id)

    def compute_single_link_local_jacobian(self, qpos, link_id):
        self.r.compute_forward_kinematics(np.asarray(qpos, dtype=np.float64))
        Jw = self.r.link_jacobians([link_id])[0]  # world-aligned linear part
        R = self.r.get_link_pose(link_id)[:3, :3]
        J = np.zeros((6, self.dof))
        J[:3] = R.T @ Jw  # LOCAL frame, as pin.computeFrameJacobian's default
        return J


def build_reference(key, override=None):
    """RetargetingConfig.build (retargeting_config.py:167-257) with the URDF/pinocchio steps replaced by DuckRobot."""
    from dex_retargeting.kinematics_adaptor import MimicJointKinematicAdaptor
    from dex_retargeting.optimizer import DexPilotOptimizer, PositionOptimizer, VectorOptimizer
    from dex_retargeting.optimizer_utils import LPFilter
    from dex_retargeting.seq_retarget import SeqRetargeting

    cfg = dict(configs()[key])
    cfg.update(override or {})
    add_dummy = bool(cfg.get("add_dummy_free_joint", False))
    robot = DuckRobot(str(ROBOTS / (Path(cfg["urdf_path"]).stem + ".json")), add_dummy)
    names = cfg.get("target_joint_names")
    if add_dummy and names is not None:
        names = ["dummy_x_translation_joint", "dummy_y_translation_joint", "dummy_z_translation_joint",
                 "dummy_x_rotation_joint", "dummy_y_rotation_joint", "dummy_z_rotation_joint"] + list(names)
    joint_names = list(names) if names is not None else robot.dof_joint_names
    hi = cfg.get("target_link_human_indices")
    hi = np.array(hi) if hi is not None else None
    typ = cfg["type"].lower()
    nd, hd = cfg.get("normal_delta", 4e-3), cfg.get("huber_delta", 2e-2)
    if typ == "position":
        opt = PositionOptimizer(robot, joint_names, target_link_names=cfg["target_link_names"],
                                target_link_human_indices=hi, norm_delta=nd, huber_delta=hd)
    elif typ == "vector":
        opt = VectorOptimizer(robot, joint_names, target_origin_link_names=cfg["target_origin_link_names"],
                              target_task_link_names=cfg["target_task_link_names"], target_link_human_indices=hi,
                              scaling=cfg.get("scaling_factor", 1.0), norm_delta=nd, huber_delta=hd)
    else:
        opt = DexPilotOptimizer(robot, joint_names, finger_tip_link_names=cfg["finger_tip_link_names"],
                                wrist_link_name=cfg["wrist_link_name"], target_link_human_indices=hi,
                                scaling=cfg.get("scaling_factor", 1.0), project_dist=cfg.get("project_dist", 0.03),
                                escape_dist=cfg.get("escape_dist", 0.05))
    alpha = cfg.get("low_pass_alpha", 0.1)
    lp = LPFilter(alpha) if 0 <= alpha <= 1 else None
    src, mim, mul, off = robot.r.mimic_spec()
    if mim and not cfg.get("ignore_mimic_joint", False):
        opt.set_kinematic_adaptor(MimicJointKinematicAdaptor(robot, target_joint_names=joint_names, source_joint_names=src,
                                                             mimic_joint_names=mim, multipliers=mul, offsets=off))
    return SeqRetargeting(opt, has_joint_limits=cfg.get("has_joint_limits", True), lp_filter=lp)


CASES = [  # (name, config key, n, init noise, target noise, seed)
    ("allegro_vector", "teleop/allegro_hand_right", 12, 0.05, 0.01, 201),
    ("shadow_position_dummy", "offline/shadow_hand_right", 8, 0.05, 0.005, 202),
    ("leap_dexpilot", "teleop/leap_hand_right_dexpilot", 12, 0.05, 0.01, 203),
    ("svh_vector_mimic", "teleop/schunk_svh_hand_right", 12, 0.05, 0.01, 204),
    ("inspire_position_mimic_dummy", "offline/inspire_hand_right", 8, 0.05, 0.005, 205),
    ("panda_vector_prismatic", "teleop/panda_gripper", 8, 0.01, 0.002, 206),
    ("ability_dexpilot_mimic", "teleop/ability_hand_right_dexpilot", 12, 0.05, 0.01, 207),
]
STREAMS = [("allegro_vector", "teleop/allegro_hand_right"), ("leap_dexpilot", "teleop/leap_hand_right_dexpilot"),
           ("svh_vector_mimic", "teleop/schunk_svh_hand_right"), ("shadow_dexpilot_5finger", "teleop/shadow_hand_right_dexpilot")]
N_POINTS = 3  # points per problem at which the closure is evaluated: the warm start and two perturbations of it


def main():
    ref_root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    install_shims(ref_root)
    out = {}
    for name, key, n, noise, tnoise, seed in CASES:
        o = build_oracle(key)  # only used to synthesise reachable problems
        seq = build_reference(key)
        opt = seq.optimizer
        rng = np.random.RandomState(seed)
        refs, fixed, x0, _ = synth_problems(o, n, rng, init_noise=noise, target_noise=tnoise)
        if opt.retargeting_type == "DEXPILOT":  # pull some thumb-finger vectors inside the projection band
            for i in range(0, n, 2):
                k = rng.randint(0, opt.num_fingers - 1)
                refs[i, k] *= np.float32(0.02 / max(np.linalg.norm(refs[i, k]), 1e-6))
        pts = np.empty((n, N_POINTS, opt.opt_dof))
        vals = np.empty((n, N_POINTS))
        grads = np.empty((n, N_POINTS, opt.opt_dof))
        proj = []
        sols = np.empty((n, opt.opt_dof), dtype=np.float32)
        sol_cost = np.empty(n)
        for i in range(n):
            if opt.retargeting_type == "DEXPILOT":
                opt.projected[:] = False
            fn = opt.get_objective_function(refs[i], fixed[i], x0[i].astype(np.float32))
            for p in range(N_POINTS):
                x = x0[i].astype(np.float64) + (rng.randn(opt.opt_dof) * 0.03 if p else 0.0)
                g = np.zeros(opt.opt_dof)
                vals[i, p] = fn(x, g)
                pts[i, p], grads[i, p] = x, g
            if opt.retargeting_type == "DEXPILOT":
                proj.append(opt.projected.copy())
                opt.projected[:] = False
            sols[i] = opt.retarget(refs[i], fixed[i], x0[i])
            # the consistent objective (value + the regulariser the reference only puts in the gradient) at the reference's
            # answer, from the reference's closure: the CUDA path must end at or below it (tests/test_gpu_golden.py)
            xs = sols[i].astype(np.float64)
            sol_cost[i] = fn(xs, np.zeros(0)) + opt.norm_delta * float(((xs - x0[i].astype(np.float32)) ** 2).sum())
        out[f"{name}/key"] = np.array(key)
        out[f"{name}/ref_value"], out[f"{name}/fixed_qpos"], out[f"{name}/last_qpos"] = refs, fixed, x0
        out[f"{name}/points"], out[f"{name}/values"], out[f"{name}/grads"] = pts, vals, grads
        out[f"{name}/retarget"], out[f"{name}/retarget_cost"] = sols, sol_cost
        out[f"{name}/idx_pin2target"] = np.asarray(opt.idx_pin2target)
        out[f"{name}/idx_pin2fixed"] = np.asarray(opt.idx_pin2fixed)
        if proj:
            out[f"{name}/projected"] = np.array(proj)
        print(name, "values", vals[:, 0].mean(), "|grad|", np.abs(grads).max())
    kp = keypoint_trajectory()
    for name, key in STREAMS:
        seq = build_reference(key)
        idx = seq.optimizer.target_link_human_indices
        frames = kp[0:160:4]
        qs, flags = [], []
        for f in frames:
            ref = f[idx[1, :], :] - f[idx[0, :], :]  # example/profiling/profile_online_retargeting.py:24-30
            qs.append(seq.retarget(ref, fixed_qpos=np.zeros(len(seq.optimizer.idx_pin2fixed))))
            if seq.optimizer.retargeting_type == "DEXPILOT":
                flags.append(seq.optimizer.projected.copy())
        out[f"stream_{name}/key"] = np.array(key)
        out[f"stream_{name}/keypoints"] = frames.astype(np.float32)
        out[f"stream_{name}/robot_qpos"] = np.array(qs)
        if flags:
            out[f"stream_{name}/projected"] = np.array(flags)
        print("stream", name, np.array(qs).shape)
    np.savez_compressed(ROOT / "tests" / "golden" / "reference_vectors.npz", **out)
    print("wrote tests/golden/reference_vectors.npz")


if __name__ == "__main__":
    main()
