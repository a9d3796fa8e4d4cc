"""Mimic-joint constraint, host side (interface of src/dex_retargeting/kinematics_adaptor.py:9-113).

The constraint is held as ONE affine map from the optimised joints to the full joint vector,

    q_pin = E x_target + S q_pin_fixed_part,        q_pin[mimic_i] = multiplier_i * q_pin[source_i] + offset_i

i.e. a (dof x n_target) matrix `fold` whose column j has a 1 in the row of target joint j and `multiplier_i` in the row of
every joint that mimics it.  `backward_jacobian` is then a single product `J @ fold` (kinematics_adaptor.py:107-113 does the
same with a gather and a loop) and the robot-table compiler reads the very same structure (`mimic_src / mimic_mult /
mimic_off`, `group_*` in include/dexr.h) for the kernel's "mimic fold".  User code and tests keep calling
`adaptor.forward_qpos` / `adaptor.backward_jacobian` / `adaptor.idx_pin2mimic` exactly as with the reference.
"""
from typing import List, Sequence

import numpy as np

from .robot_wrapper import RobotWrapper


class KinematicAdaptor:
    """Base: remembers the robot and where the optimised joints sit in pinocchio order."""

    def __init__(self, robot: RobotWrapper, target_joint_names: List[str]):
        self.robot = robot
        self.target_joint_names = target_joint_names
        self.idx_pin2target = np.fromiter((robot.get_joint_index(n) for n in target_joint_names), dtype=int,
                                          count=len(target_joint_names))

    def forward_qpos(self, qpos: np.ndarray) -> np.ndarray:  # pragma: no cover - interface
        raise NotImplementedError

    def backward_jacobian(self, jacobian: np.ndarray) -> np.ndarray:  # pragma: no cover - interface
        raise NotImplementedError


def _lookup(names: Sequence[str], table: Sequence[str]) -> np.ndarray:
    pos = {n: i for i, n in enumerate(table)}
    missing = [n for n in names if n not in pos]
    if missing:
        raise ValueError(f"{missing} is not in list {list(table)}")
    return np.array([pos[n] for n in names], dtype=int)


class MimicJointKinematicAdaptor(KinematicAdaptor):
    def __init__(self, robot: RobotWrapper, target_joint_names: List[str], source_joint_names: List[str],
                 mimic_joint_names: List[str], multipliers: List[float], offsets: List[float]):
        super().__init__(robot, target_joint_names)
        self.source_joint_names, self.mimic_joint_names = list(source_joint_names), list(mimic_joint_names)
        self.multipliers = np.asarray(multipliers, dtype=float).reshape(-1)
        self.offsets = np.asarray(offsets, dtype=float).reshape(-1)

        counts = dict(source=len(self.source_joint_names), mimic=len(self.mimic_joint_names),
                      multiplier=self.multipliers.size, offset=self.offsets.size)
        if len(set(counts.values())) > 1:
            raise ValueError("Mimic joints setting dimension mismatch: " + ", ".join(f"{k} {v}" for k, v in counts.items()))
        if len(set(self.mimic_joint_names)) != counts["mimic"]:
            raise ValueError(f"Redundant mimic joint names: {mimic_joint_names}")
        optimised_mimics = [n for n in self.mimic_joint_names if n in set(target_joint_names)]
        if optimised_mimics:
            raise ValueError(
                f"Mimic joint should not be one of the target joints: {optimised_mimics} are driven by other joints and cannot "
                f"be optimised.  List the target joints explicitly in the retargeting config of a robot with mimic joints "
                f"(given: {target_joint_names})")

        dof_names = robot.dof_joint_names
        self.idx_pin2source = _lookup(self.source_joint_names, dof_names)
        self.idx_pin2mimic = _lookup(self.mimic_joint_names, dof_names)
        self.idx_target2source = _lookup(self.source_joint_names, self.target_joint_names)
        self.num_active_joints = len(dof_names) - counts["mimic"]

        # the affine map's linear part: d q_pin / d x_target
        self.fold = np.zeros((len(dof_names), len(target_joint_names)))
        self.fold[self.idx_pin2target, np.arange(len(target_joint_names))] = 1.0
        np.add.at(self.fold, (self.idx_pin2mimic, self.idx_target2source), self.multipliers)

    def forward_qpos(self, pin_qpos: np.ndarray) -> np.ndarray:
        """In place, like the reference: the mimic entries of `pin_qpos` are overwritten from their sources."""
        np.put(pin_qpos, self.idx_pin2mimic, np.take(pin_qpos, self.idx_pin2source) * self.multipliers + self.offsets)
        return pin_qpos

    def backward_jacobian(self, jacobian: np.ndarray) -> np.ndarray:
        """(..., dof) in pinocchio order -> (..., n_target): chain rule through the affine map."""
        return np.asarray(jacobian) @ self.fold
