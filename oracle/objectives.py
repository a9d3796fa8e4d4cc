"""Oracle objectives: the three retargeting losses and their gradients, float64 numpy.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates, relative to /root/reference:
  src/dex_retargeting/retargeting_config.py:167-257  build(): model, joint names, optimizer ctor
        arguments (note :218-228 -- DexPilot receives only scaling/project_dist/escape_dist, so its
        huber_delta / norm_delta stay at the ctor defaults 0.03 / 4e-3 whatever the config says),
        low-pass filter, mimic adaptor wiring
  src/dex_retargeting/optimizer.py:18-75     index maps, bounds widened by epsilon=1e-3, mimic
        joints removed from the fixed set
  src/dex_retargeting/optimizer.py:138-200   position objective  (SmoothL1 per coordinate, mean)
  src/dex_retargeting/optimizer.py:241-306   vector objective    (SmoothL1 of the vector norm, mean)
  src/dex_retargeting/optimizer.py:407-454   DexPilot link indices and projection cache
  src/dex_retargeting/optimizer.py:456-577   DexPilot objective  (hysteresis flags, weights,
        projected reference vectors, weighted SmoothL1 sum / m)
`value_and_grad(x)` returns the pair exactly as the reference hands it to nlopt: the VALUE omits the
norm_delta term, the GRADIENT includes it (optimizer.py:166-167 vs :194).  `consistent(x)` returns
value + norm_delta*|x - x_last|^2, the function whose gradient that actually is.
`torch_value_and_grad` evaluates the same thing with torch SmoothL1Loss + autograd, verbatim in
structure, to check the closed forms.
"""
import json
from pathlib import Path

import numpy as np

from .robot import DUMMY_JOINTS, OracleMimic, OracleRobot


def smooth_l1(d, beta):
    """torch.nn.SmoothL1Loss element: 0.5 d^2 / beta if |d| < beta else |d| - 0.5 beta; and its derivative."""
    a = np.abs(d)
    quad = a < beta
    val = np.where(quad, 0.5 * d * d / beta, a - 0.5 * beta)
    der = np.where(quad, d / beta, np.sign(d))
    return val, der


def generate_link_indices(num_fingers):
    """optimizer.py:407-428."""
    origin, task = [], []
    for i in range(1, num_fingers):
        for j in range(i + 1, num_fingers + 1):
            origin.append(j)
            task.append(i)
    for i in range(1, num_fingers + 1):
        origin.append(0)
        task.append(i)
    return origin, task


def set_dexpilot_cache(num_fingers, eta1, eta2):
    """optimizer.py:430-454."""
    projected = np.zeros(num_fingers * (num_fingers - 1) // 2, dtype=bool)
    s2_origin, s2_task = [], []
    for i in range(0, num_fingers - 2):
        for j in range(i + 1, num_fingers - 1):
            s2_origin.append(j)
            s2_task.append(i)
    dist = np.array([eta1] * (num_fingers - 1) + [eta2] * ((num_fingers - 1) * (num_fingers - 2) // 2))
    return projected, s2_origin, s2_task, dist


class OracleOptimizer:
    """One object per (config, robot): index maps + loss parameters + DexPilot state."""

    def __init__(self, cfg, robots_dir, override=None):
        cfg = dict(cfg)
        if override:
            cfg.update(override)
        self.cfg = cfg
        self.type = cfg["type"].lower()
        if self.type not in ("vector", "position", "dexpilot"):
            raise ValueError("Retargeting type must be one of ['vector', 'position', 'dexpilot']")
        add_dummy = bool(cfg.get("add_dummy_free_joint", False))
        stem = Path(cfg["urdf_path"]).stem
        p = Path(robots_dir) / (stem + ".json")
        self.robot = OracleRobot(str(p) if p.exists() else str(Path(robots_dir) / cfg["urdf_path"]), add_dummy)
        robot = self.robot

        names = cfg.get("target_joint_names")
        if add_dummy and names is not None:
            names = DUMMY_JOINTS + list(names)
        self.target_joint_names = list(names) if names is not None else list(robot.dof_joint_names)
        for n in self.target_joint_names:
            if n not in robot.dof_joint_names:
                raise ValueError(f"Joint {n} given does not appear to be in robot XML.")
        self.idx_pin2target = np.array([robot.dof_joint_names.index(n) for n in self.target_joint_names])
        self.idx_pin2fixed = np.array([i for i in range(robot.dof) if i not in self.idx_pin2target], dtype=int)
        self.opt_dof = len(self.idx_pin2target)

        self.norm_delta = float(cfg.get("normal_delta", 4e-3))
        self.huber_delta = float(cfg.get("huber_delta", 2e-2))
        self.scaling = float(cfg.get("scaling_factor", 1.0))
        hi = cfg.get("target_link_human_indices")
        if self.type == "position":
            self.target_link_indices = [robot.get_link_index(n) for n in cfg["target_link_names"]]
            self.target_link_human_indices = np.asarray(hi).squeeze()
            self.link_ids = list(self.target_link_indices)
            self.ftol = 1e-5
            self.m = len(self.link_ids)
        else:
            if self.type == "dexpilot":
                tips = list(cfg["finger_tip_link_names"])
                if not 2 <= len(tips) <= 5:
                    raise ValueError("DexPilot optimizer can only be applied to hands with 2 to 5 fingers")
                self.num_fingers = len(tips)
                oi, ti = generate_link_indices(self.num_fingers)
                if hi is None:
                    hi = (np.stack([oi, ti], axis=0) * 4).astype(int)
                ln = [cfg["wrist_link_name"]] + tips
                origin_names, task_names = [ln[i] for i in oi], [ln[i] for i in ti]
                # retargeting_config.py:218-228 does not forward these two
                self.huber_delta, self.norm_delta = 0.03, 4e-3
                self.project_dist = float(cfg.get("project_dist", 0.03))
                self.escape_dist = float(cfg.get("escape_dist", 0.05))
                self.eta1, self.eta2 = 1e-4, 3e-2
                (self.projected, self.s2_origin, self.s2_task, self.projected_dist) = set_dexpilot_cache(
                    self.num_fingers, self.eta1, self.eta2)
            else:
                origin_names, task_names = list(cfg["target_origin_link_names"]), list(cfg["target_task_link_names"])
            self.target_link_human_indices = np.asarray(hi)
            computed = sorted(set(origin_names) | set(task_names))  # reference uses list(set()), order irrelevant
            self.origin_sel = np.array([computed.index(n) for n in origin_names])
            self.task_sel = np.array([computed.index(n) for n in task_names])
            self.link_ids = [robot.get_link_index(n) for n in computed]
            self.ftol = 1e-6
            self.m = len(origin_names)

        # mimic adaptor (retargeting_config.py:237-250)
        self.adaptor = None
        src, mim, mul, off = robot.mimic_spec()
        if mim and not cfg.get("ignore_mimic_joint", False):
            self.adaptor = OracleMimic(robot, self.target_joint_names, src, mim, mul, off)
            self.idx_pin2fixed = np.array([i for i in self.idx_pin2fixed if i not in self.adaptor.idx_pin2mimic],
                                          dtype=int)

        # SeqRetargeting.__init__ (seq_retarget.py:20-35) + set_joint_limit (optimizer.py:54-60)
        lim = np.ones_like(robot.joint_limits)
        lim[:, 0], lim[:, 1] = -1e4, 1e4
        if cfg.get("has_joint_limits", True):
            lim = robot.joint_limits.copy()
        self.joint_limits = lim[self.idx_pin2target]
        self.lower = self.joint_limits[:, 0] - 1e-3
        self.upper = self.joint_limits[:, 1] + 1e-3
        self.low_pass_alpha = float(cfg.get("low_pass_alpha", 0.1))

    # ---------------------------------------------------------------------------------------
    def ref_from_keypoints(self, kp):
        """Caller-side gather, example/profiling/profile_online_retargeting.py:24-30."""
        idx = self.target_link_human_indices
        if self.type == "position":
            return kp[idx, :]
        return kp[idx[1, :], :] - kp[idx[0, :], :]

    def full_qpos(self, x, fixed_qpos):
        q = np.zeros(self.robot.dof)
        q[self.idx_pin2fixed] = fixed_qpos
        q[self.idx_pin2target] = x
        if self.adaptor is not None:
            q = self.adaptor.forward_qpos(q)
        return q

    def prepare(self, ref_value, update_state=True):
        """Per-frame constants: (target array, per-residual weights).  DexPilot updates `projected`."""
        ref_value = np.asarray(ref_value)
        if self.type == "position":
            return ref_value.astype(np.float64), None
        if self.type == "vector":
            return (ref_value * ref_value.dtype.type(self.scaling)).astype(np.float64), np.ones(self.m)
        # dexpilot, optimizer.py:460-508
        len_proj = len(self.projected)
        len_s2 = len(self.s2_task)
        len_s1 = len_proj - len_s2
        proj = self.projected if update_state else self.projected.copy()
        dist = np.linalg.norm(ref_value[:len_proj], axis=1)
        proj[:len_s1][dist[0:len_s1] < self.project_dist] = True
        proj[:len_s1][dist[0:len_s1] > self.escape_dist] = False
        proj[len_s1:len_proj] = np.logical_and(proj[:len_s1][self.s2_origin], proj[:len_s1][self.s2_task])
        proj[len_s1:len_proj] = np.logical_and(proj[len_s1:len_proj], dist[len_s1:len_proj] <= 0.03)
        normal_w = np.ones(len_proj, dtype=np.float32)
        high_w = np.array([200] * len_s1 + [400] * len_s2, dtype=np.float32)
        w = np.where(proj, high_w, normal_w)
        w = np.concatenate([w, np.ones(self.num_fingers, dtype=np.float32) * len_proj + self.num_fingers])
        normal_vec = ref_value * self.scaling
        dir_vec = ref_value[:len_proj] / (dist[:, None] + 1e-6)
        projected_vec = dir_vec * self.projected_dist[:, None]
        ref = np.where(proj[:, None], projected_vec, normal_vec[:len_proj])
        ref = np.concatenate([ref, normal_vec[len_proj:]], axis=0).astype(np.float32)
        return ref.astype(np.float64), w.astype(np.float64)

    # ---------------------------------------------------------------------------------------
    def make_objective(self, ref_value, fixed_qpos, last_qpos, update_state=True):
        target, weights = self.prepare(ref_value, update_state)
        return FrameObjective(self, target, weights, np.asarray(fixed_qpos, float), np.asarray(last_qpos, np.float32).astype(float))


class FrameObjective:
    def __init__(self, opt, target, weights, fixed_qpos, last_qpos):
        self.o, self.target, self.weights, self.fixed, self.last = opt, target, weights, fixed_qpos, last_qpos
        self.n_eval = 0

    def _kin(self, x, need_jac):
        o = self.o
        q = o.full_qpos(x, self.fixed)
        o.robot.compute_forward_kinematics(q)
        pos = o.robot.link_positions(o.link_ids)
        if not need_jac:
            return pos, None
        J = o.robot.link_jacobians(o.link_ids)
        J = o.adaptor.backward_jacobian(J) if o.adaptor is not None else J[..., o.idx_pin2target]
        return pos, J

    def _loss(self, pos, need_grad):
        """value, d value / d pos  (len(link_ids), 3)."""
        o = self.o
        if o.type == "position":
            v, d = smooth_l1(pos - self.target, o.huber_delta)
            return v.mean(), (d / v.size if need_grad else None)
        vec = pos[o.task_sel] - pos[o.origin_sel]
        diff = vec - self.target
        dist = np.linalg.norm(diff, axis=1)
        v, d = smooth_l1(dist, o.huber_delta)
        w = self.weights / o.m
        val = float((v * w).sum())
        if not need_grad:
            return val, None
        with np.errstate(invalid="ignore", divide="ignore"):
            unit = np.where(dist[:, None] > 0, diff / dist[:, None], 0.0)  # torch.norm backward at 0 -> 0
        gvec = unit * (d * w)[:, None]
        gpos = np.zeros_like(pos)
        np.add.at(gpos, o.task_sel, gvec)
        np.add.at(gpos, o.origin_sel, -gvec)
        return val, gpos

    def value(self, x):
        pos, _ = self._kin(np.asarray(x, float), False)
        return float(self._loss(pos, False)[0])

    def value_and_grad(self, x):
        """(value without the norm_delta term, gradient with it) -- what nlopt is given."""
        self.n_eval += 1
        x = np.asarray(x, float)
        pos, J = self._kin(x, True)
        val, gpos = self._loss(pos, True)
        grad = np.einsum("lc,lcn->n", gpos, J) + 2.0 * self.o.norm_delta * (x - self.last)
        return float(val), grad

    def consistent(self, x):
        x = np.asarray(x, float)
        return self.value(x) + self.o.norm_delta * float(((x - self.last) ** 2).sum())

    def task_error(self, x):
        """Mean Euclidean error as asserted by tests/test_optimizer.py (:130-141, :196-209)."""
        pos, _ = self._kin(np.asarray(x, float), False)
        if self.o.type == "position":
            return float(np.linalg.norm(pos - self.target, axis=-1).mean())
        return float(np.linalg.norm(pos[self.o.task_sel] - pos[self.o.origin_sel] - self.target, axis=-1).mean())

    def torch_value_and_grad(self, x):
        """Same quantity through torch SmoothL1Loss + autograd, structured like optimizer.py."""
        import torch

        o = self.o
        x = np.asarray(x, float)
        pos, J = self._kin(x, True)
        tp = torch.as_tensor(pos).requires_grad_()
        tt = torch.as_tensor(self.target)
        if o.type == "position":
            loss = torch.nn.SmoothL1Loss(beta=o.huber_delta)(tp, tt)
        else:
            rv = tp[torch.as_tensor(o.task_sel), :] - tp[torch.as_tensor(o.origin_sel), :]
            dist = torch.norm(rv - tt, dim=1, keepdim=False)
            if o.type == "vector":
                loss = torch.nn.SmoothL1Loss(beta=o.huber_delta, reduction="mean")(dist, torch.zeros_like(dist))
            else:
                loss = (torch.nn.SmoothL1Loss(beta=o.huber_delta, reduction="none")(dist, torch.zeros_like(dist))
                        * torch.as_tensor(self.weights) / rv.shape[0]).sum()
        loss.backward()
        gpos = tp.grad.numpy()[:, None, :]
        g = np.matmul(gpos, J).mean(1).sum(0) + 2 * o.norm_delta * (x - self.last)
        return float(loss.item()), g


def load_configs(golden_dir):
    with open(Path(golden_dir) / "configs.json") as f:
        return json.load(f)
