"""pinocchio-free RobotWrapper (host, float64).

API mirror of src/dex_retargeting/robot_wrapper.py:8-95 so that code written against the reference
(`robot.dof`, `robot.joint_limits`, `robot.compute_forward_kinematics`, `robot.get_link_pose`, ...)
keeps working without pinocchio.  The hot path does NOT go through this class: the solver kernel has
its own fp32 forward kinematics built from the same `KinematicModel` (table.py).  This is the
float64 bookkeeping / evaluation side (test-data generation, task-space error reports, warm_start).
"""
from types import SimpleNamespace
from typing import List

import numpy as np
import numpy.typing as npt

from .urdf import KinematicModel


class RobotWrapper:
    """This class does not take mimic joint into consideration (same contract as the reference)."""

    def __init__(self, urdf_path, use_collision=False, use_visual=False, add_dummy_free_joints: bool = False):
        if use_visual or use_collision:
            raise NotImplementedError
        if isinstance(urdf_path, KinematicModel):
            self.kin = urdf_path
        else:
            self.kin = KinematicModel.load(urdf_path, add_dummy_free_joints)
        k = self.kin
        # minimal stand-in for the attributes of pin.Model user code / tests touch
        self.model = SimpleNamespace(nq=k.dof, nv=k.dof, names=["universe"] + list(k.dof_joint_names),
                                     lowerPositionLimit=k.joint_limits[:, 0].copy(),
                                     upperPositionLimit=k.joint_limits[:, 1].copy())
        self.q0 = np.zeros(k.dof)
        self._Rw = None
        self._pw = None

    # ---------------------------------------------------------------- properties
    @property
    def joint_names(self) -> List[str]:
        return list(self.model.names)

    @property
    def dof_joint_names(self) -> List[str]:
        return list(self.kin.dof_joint_names)

    @property
    def dof(self) -> int:
        return self.kin.dof

    @property
    def link_names(self) -> List[str]:
        # pinocchio lists every frame: bodies and joints.  Bodies first here; only membership
        # tests and `"dummy" in name` counts are made on this list (optimizer.py:47-48).
        return list(self.kin.link_names) + [j.name for j in self.kin.joints]

    @property
    def joint_limits(self):
        return self.kin.joint_limits.copy()

    # ---------------------------------------------------------------- queries
    def get_joint_index(self, name: str):
        return self.dof_joint_names.index(name)

    def get_link_index(self, name: str):
        if name not in self.link_names:
            raise ValueError(f"{name} is not a link name. Valid link names: \n{self.link_names}")
        return self.kin.link_index(name)

    def get_joint_parent_child_frames(self, joint_name: str):
        j = self.kin.joint_map.get(joint_name)
        if j is None:
            raise ValueError(f"Can not find child link of {joint_name}")
        return self.kin.link_index(j.parent), self.kin.link_index(j.child)

    # ---------------------------------------------------------------- kinematics
    def compute_forward_kinematics(self, qpos: npt.NDArray):
        self._Rw, self._pw = self.kin.forward_kinematics(np.asarray(qpos, dtype=np.float64))

    def _pose(self, link_id: int):
        if self._Rw is None:
            raise RuntimeError("compute_forward_kinematics must be called first")
        return self.kin.link_pose(self._Rw, self._pw, link_id)

    def get_link_pose(self, link_id: int) -> npt.NDArray:
        R, p = self._pose(link_id)
        T = np.eye(4)
        T[:3, :3], T[:3, 3] = R, p
        return T

    def get_link_pose_inv(self, link_id: int) -> npt.NDArray:
        R, p = self._pose(link_id)
        T = np.eye(4)
        T[:3, :3], T[:3, 3] = R.T, -R.T @ p
        return T

    def compute_single_link_local_jacobian(self, qpos, link_id: int) -> npt.NDArray:
        """6 x dof frame Jacobian expressed in the LOCAL link frame (rows 0-2 linear, 3-5 angular)."""
        k = self.kin
        self.compute_forward_kinematics(qpos)
        R, p = self._pose(link_id)
        J = np.zeros((6, k.dof))
        j = int(k.link_parent[link_id])
        while j >= 0:
            a = self._Rw[j] @ k.joint_axis[j]
            if k.joint_type[j] == 0:
                J[:3, j] = R.T @ np.cross(a, p - self._pw[j])
                J[3:, j] = R.T @ a
            else:
                J[:3, j] = R.T @ a
            j = int(k.joint_parent[j])
        return J
