"""Mimic-joint constraint (host mirror of src/dex_retargeting/kinematics_adaptor.py:9-113).

On the device the same constraint is part of the robot table: `mimic_src/mult/off` drive the joint
value and the `group_*` lists fold Jacobian / Hessian columns (csrc/dexr_kernels.cuh, "mimic fold").
These host classes exist so that user code and tests can keep calling `adaptor.forward_qpos`.
"""
from abc import abstractmethod
from typing import List

import numpy as np

from .robot_wrapper import RobotWrapper


class KinematicAdaptor:
    def __init__(self, robot: RobotWrapper, target_joint_names: List[str]):
        self.robot = robot
        self.target_joint_names = target_joint_names
        self.idx_pin2target = np.array([robot.get_joint_index(n) for n in target_joint_names])

    @abstractmethod
    def forward_qpos(self, qpos: np.ndarray) -> np.ndarray:
        """Apply the kinematic constraint to a qpos in pinocchio joint order (same shape out)."""

    @abstractmethod
    def backward_jacobian(self, jacobian: np.ndarray) -> np.ndarray:
        """Map a Jacobian in pinocchio joint order to target joint order."""


class MimicJointKinematicAdaptor(KinematicAdaptor):
    def __init__(self, robot: RobotWrapper, target_joint_names: List[str], source_joint_names: List[str],
                 mimic_joint_names: List[str], multipliers: List[float], offsets: List[float]):
        super().__init__(robot, target_joint_names)
        self.multipliers = np.array(multipliers)
        self.offsets = np.array(offsets)
        self.source_joint_names = list(source_joint_names)
        self.mimic_joint_names = list(mimic_joint_names)

        clash = set(mimic_joint_names) & set(target_joint_names)
        if clash:
            raise ValueError(
                f"Mimic joint should not be one of the target joints.\n"
                f"Mimic joints: {mimic_joint_names}.\n"
                f"Target joints: {target_joint_names}\n"
                f"You need to specify the target joint names explicitly in your retargeting config"
                f" for robot with mimic joint constraints: {target_joint_names}"
            )
        self.idx_pin2source = np.array([robot.get_joint_index(n) for n in source_joint_names])
        self.idx_pin2mimic = np.array([robot.get_joint_index(n) for n in mimic_joint_names])
        self.idx_target2source = np.array([self.target_joint_names.index(n) for n in source_joint_names])

        sizes = (len(self.idx_target2source), len(self.idx_pin2mimic), len(self.multipliers), len(self.offsets))
        if len(set(sizes)) != 1:
            raise ValueError(
                f"Mimic joints setting dimension mismatch.\n"
                f"Source joints: {sizes[0]}, mimic joints: {sizes[1]}, multiplier: {sizes[2]}, offset: {sizes[3]}"
            )
        self.num_active_joints = len(robot.dof_joint_names) - sizes[1]
        if len(mimic_joint_names) != len(np.unique(mimic_joint_names)):
            raise ValueError(f"Redundant mimic joint names: {mimic_joint_names}")

    def forward_qpos(self, pin_qpos: np.ndarray) -> np.ndarray:
        pin_qpos[self.idx_pin2mimic] = pin_qpos[self.idx_pin2source] * self.multipliers + self.offsets
        return pin_qpos

    def backward_jacobian(self, jacobian: np.ndarray) -> np.ndarray:
        out = jacobian[..., self.idx_pin2target]
        scaled = jacobian[..., self.idx_pin2mimic] * self.multipliers
        for i, tgt in enumerate(self.idx_target2source):
            out[..., tgt] += scaled[..., i]
        return out
