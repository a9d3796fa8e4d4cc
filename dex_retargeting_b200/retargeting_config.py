"""Retargeting configuration: dataclass + YAML loader + factory.

Keeps the reference schema and semantics (src/dex_retargeting/retargeting_config.py:18-285): the
same YAML keys and defaults, the `override` dict, the class-level default URDF directory, the same
construction order in `build()` -- including two reference quirks that change results:
  * DexPilot is built with only scaling / project_dist / escape_dist from the config; huber_delta and
    normal_delta stay at the DexPilotOptimizer defaults (retargeting_config.py:218-228);
  * with `add_dummy_free_joint` and explicit `target_joint_names`, the six dummy joints are
    prepended to the optimised joints (retargeting_config.py:190-191).
Differences: the URDF is read by `dex_retargeting_b200.urdf` (stdlib XML; no yourdfpy / pinocchio /
temp-file round trip) and `urdf_path` may also name a `.json` robot description.
"""
from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path
from typing import Any, Dict, List, Optional, Tuple, Union

import numpy as np
import yaml

from .kinematics_adaptor import MimicJointKinematicAdaptor
from .optimizer_utils import LPFilter
from .robot_wrapper import RobotWrapper
from .seq_retarget import SeqRetargeting
from .urdf import DUMMY_JOINT_NAMES, KinematicModel


@dataclass
class RetargetingConfig:
    type: str
    urdf_path: str

    # Free joint at the robot root: lets the hand move freely in space
    add_dummy_free_joint: bool = False

    # Human keypoint index of each target link (position) or of each vector's origin / task (2 x m)
    target_link_human_indices: Optional[np.ndarray] = None

    # Robot link corresponding to the human wrist (dexpilot)
    wrist_link_name: Optional[str] = None

    # Position retargeting
    target_link_names: Optional[List[str]] = None

    # Vector retargeting
    target_joint_names: Optional[List[str]] = None
    target_origin_link_names: Optional[List[str]] = None
    target_task_link_names: Optional[List[str]] = None

    # DexPilot retargeting
    finger_tip_link_names: Optional[List[str]] = None

    # Human -> robot size ratio (vector / dexpilot)
    scaling_factor: float = 1.0

    # Objective parameters
    normal_delta: float = 4e-3
    huber_delta: float = 2e-2

    # DexPilot projection thresholds
    project_dist: float = 0.03
    escape_dist: float = 0.05

    has_joint_limits: bool = True
    ignore_mimic_joint: bool = False

    # Low pass filter: smaller alpha = smoother and laggier; outside [0, 1] = no filter
    low_pass_alpha: float = 0.1

    _TYPE = ["vector", "position", "dexpilot"]
    _DEFAULT_URDF_DIR = "./"

    def __post_init__(self):
        self.type = self.type.lower()
        if self.type not in self._TYPE:
            raise ValueError(f"Retargeting type must be one of {self._TYPE}")

        if self.type == "vector":
            if self.target_origin_link_names is None or self.target_task_link_names is None:
                raise ValueError("Vector retargeting requires: target_origin_link_names + target_task_link_names")
            if len(self.target_task_link_names) != len(self.target_origin_link_names):
                raise ValueError("Vector retargeting origin and task links dim mismatch")
            if self.target_link_human_indices is None:
                raise ValueError("Vector retargeting requires: target_link_human_indices")
            self.target_link_human_indices = np.asarray(self.target_link_human_indices)
            if self.target_link_human_indices.shape != (2, len(self.target_origin_link_names)):
                raise ValueError("Vector retargeting link names and link indices dim mismatch")
        elif self.type == "position":
            if self.target_link_names is None:
                raise ValueError("Position retargeting requires: target_link_names")
            if self.target_link_human_indices is None:
                raise ValueError("Position retargeting requires: target_link_human_indices")
            self.target_link_human_indices = np.asarray(self.target_link_human_indices).squeeze()
            if self.target_link_human_indices.shape != (len(self.target_link_names),):
                raise ValueError("Position retargeting link names and link indices dim mismatch")
        elif self.type == "dexpilot":
            if self.finger_tip_link_names is None or self.wrist_link_name is None:
                raise ValueError("Position retargeting requires: finger_tip_link_names + wrist_link_name")
            if self.target_link_human_indices is not None:
                print(
                    "\033[33m",
                    "Target link human indices is provided in the DexPilot retargeting config, which is uncommon.\n"
                    "If you do not know exactly how it is used, please leave it to None for default.\n"
                    "\033[00m",
                )

        urdf_path = Path(self.urdf_path)
        if not urdf_path.is_absolute():
            urdf_path = (Path(self._DEFAULT_URDF_DIR) / urdf_path).absolute()
        if not urdf_path.exists():
            # robot descriptions exported as flat JSON joint trees (tests/golden/robots) are accepted too
            alt = (Path(self._DEFAULT_URDF_DIR) / (urdf_path.stem + ".json")).absolute()
            if not alt.exists():
                raise ValueError(f"URDF path {urdf_path} does not exist")
            urdf_path = alt
        self.urdf_path = str(urdf_path)

    @classmethod
    def set_default_urdf_dir(cls, urdf_dir: Union[str, Path]):
        path = Path(urdf_dir)
        if not path.exists():
            raise ValueError(f"URDF dir {urdf_dir} not exists.")
        cls._DEFAULT_URDF_DIR = urdf_dir

    @classmethod
    def load_from_file(cls, config_path: Union[str, Path], override: Optional[Dict] = None):
        path = Path(config_path)
        if not path.is_absolute():
            path = path.absolute()
        with path.open("r") as f:
            cfg = yaml.safe_load(f)["retargeting"]
        return cls.from_dict(cfg, override)

    @classmethod
    def from_dict(cls, cfg: Dict[str, Any], override: Optional[Dict] = None):
        cfg = dict(cfg)
        if cfg.get("target_link_human_indices") is not None:
            cfg["target_link_human_indices"] = np.array(cfg["target_link_human_indices"])
        if override is not None:
            cfg.update(override)
        return RetargetingConfig(**cfg)

    def build(self, device: Optional[int] = None) -> SeqRetargeting:
        from .optimizer import DexPilotOptimizer, PositionOptimizer, VectorOptimizer

        model = KinematicModel.load(self.urdf_path, add_dummy_free_joints=self.add_dummy_free_joint)
        robot = RobotWrapper(model)

        # the 6 dummy joints are optimised too
        if self.add_dummy_free_joint and self.target_joint_names is not None:
            self.target_joint_names = DUMMY_JOINT_NAMES + self.target_joint_names
        joint_names = self.target_joint_names if self.target_joint_names is not None else robot.dof_joint_names

        if self.type == "position":
            optimizer = PositionOptimizer(
                robot, joint_names, target_link_names=self.target_link_names,
                target_link_human_indices=self.target_link_human_indices,
                norm_delta=self.normal_delta, huber_delta=self.huber_delta, device=device,
            )
        elif self.type == "vector":
            optimizer = VectorOptimizer(
                robot, joint_names, target_origin_link_names=self.target_origin_link_names,
                target_task_link_names=self.target_task_link_names,
                target_link_human_indices=self.target_link_human_indices, scaling=self.scaling_factor,
                norm_delta=self.normal_delta, huber_delta=self.huber_delta, device=device,
            )
        elif self.type == "dexpilot":
            optimizer = DexPilotOptimizer(
                robot, joint_names, finger_tip_link_names=self.finger_tip_link_names,
                wrist_link_name=self.wrist_link_name, target_link_human_indices=self.target_link_human_indices,
                scaling=self.scaling_factor, project_dist=self.project_dist, escape_dist=self.escape_dist,
                device=device,
            )
        else:
            raise RuntimeError()

        lp_filter = LPFilter(self.low_pass_alpha) if 0 <= self.low_pass_alpha <= 1 else None

        has_mimic, source_names, mimic_names, multipliers, offsets = parse_mimic_joint(model)
        if has_mimic and not self.ignore_mimic_joint:
            optimizer.set_kinematic_adaptor(MimicJointKinematicAdaptor(
                robot, target_joint_names=joint_names, source_joint_names=source_names,
                mimic_joint_names=mimic_names, multipliers=multipliers, offsets=offsets,
            ))
        return SeqRetargeting(optimizer, has_joint_limits=self.has_joint_limits, lp_filter=lp_filter)


def get_retargeting_config(config_path: Union[str, Path]) -> RetargetingConfig:
    return RetargetingConfig.load_from_file(config_path)


def parse_mimic_joint(robot_model: KinematicModel) -> Tuple[bool, List[str], List[str], List[float], List[float]]:
    source_names, mimic_names, multipliers, offsets = robot_model.mimic_joints()
    return len(mimic_names) > 0, source_names, mimic_names, multipliers, offsets
