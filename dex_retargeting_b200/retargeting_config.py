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
    # the kinematics-only URDFs shipped with the package (the reference defaults to "./" and expects the caller to point
    # it at a dex-urdf checkout, retargeting_config.py:60; set_default_urdf_dir() still does that)
    _DEFAULT_URDF_DIR = str(Path(__file__).resolve().parent / "assets" / "robots" / "hands")  # = packaged_urdf_dir()

    # what each retargeting type needs from the config: required keys, and the shape of target_link_human_indices as a
    # function of the number of targets (None = optional, DexPilot derives it from the finger count)
    _SCHEMA = {
        "vector": dict(label="Vector", required=("target_origin_link_names", "target_task_link_names"),
                       count=lambda c: len(c.target_origin_link_names), index_shape=lambda m: (2, m)),
        "position": dict(label="Position", required=("target_link_names",),
                         count=lambda c: len(c.target_link_names), index_shape=lambda m: (m,)),
        "dexpilot": dict(label="DexPilot", required=("finger_tip_link_names", "wrist_link_name"), count=None, index_shape=None),
    }

    def __post_init__(self):
        self.type = str(self.type).lower()
        schema = self._SCHEMA.get(self.type)
        if schema is None:
            raise ValueError(f"Retargeting type must be one of {self._TYPE}")
        label = schema["label"]
        absent = [k for k in schema["required"] if getattr(self, k) is None]
        if absent:
            raise ValueError(f"{label} retargeting requires: {' + '.join(schema['required'])} (missing: {', '.join(absent)})")
        if self.type == "vector" and len(self.target_task_link_names) != len(self.target_origin_link_names):
            raise ValueError("Vector retargeting origin and task links dim mismatch: "
                             f"{len(self.target_origin_link_names)} origins, {len(self.target_task_link_names)} tasks")
        if schema["index_shape"] is not None:
            if self.target_link_human_indices is None:
                raise ValueError(f"{label} retargeting requires: target_link_human_indices")
            idx = np.asarray(self.target_link_human_indices)
            want = schema["index_shape"](schema["count"](self))
            if len(want) == 1:
                idx = idx.squeeze()
            if idx.shape != want:
                raise ValueError(f"{label} retargeting link names and link indices dim mismatch: indices {idx.shape}, expected {want}")
            self.target_link_human_indices = idx
        elif self.target_link_human_indices is not None:
            print("\033[33m DexPilot derives target_link_human_indices from the finger count; a config that sets it "
                  "overrides that default -- leave it out unless the keypoint layout really differs.\033[00m")
        self.urdf_path = str(self._resolve_robot_file(self.urdf_path))

    @classmethod
    def _resolve_robot_file(cls, name: Union[str, Path]) -> Path:
        """Absolute path of the robot description (`.urdf`, or a `.json` joint tree): as given, else under the default
        URDF directory.  A path that does not exist is an error, as in the reference (retargeting_config.py:93-96)."""
        path = Path(name)
        if not path.is_absolute():
            path = (Path(cls._DEFAULT_URDF_DIR) / path).absolute()
        if not path.exists():
            raise ValueError(f"URDF path {path} does not exist")
        return path

    @staticmethod
    def packaged_urdf_dir() -> Path:
        """Directory of the kinematics-only hand URDFs shipped with the package (same relative paths as dex-urdf's
        `robots/hands`, which is what the `urdf_path` entries of the YAML files are relative to)."""
        return Path(__file__).resolve().parent / "assets" / "robots" / "hands"

    @classmethod
    def set_default_urdf_dir(cls, urdf_dir: Union[str, Path]):
        if not Path(urdf_dir).exists():
            raise ValueError(f"URDF dir {urdf_dir} not exists.")
        cls._DEFAULT_URDF_DIR = urdf_dir

    @classmethod
    def load_from_file(cls, config_path: Union[str, Path], override: Optional[Dict] = None):
        with Path(config_path).absolute().open("r") as f:
            return cls.from_dict(yaml.safe_load(f)["retargeting"], override)

    @classmethod
    def from_dict(cls, cfg: Dict[str, Any], override: Optional[Dict] = None):
        merged = {**cfg, **(override or {})}
        if merged.get("target_link_human_indices") is not None:
            merged["target_link_human_indices"] = np.array(merged["target_link_human_indices"])
        return cls(**merged)

    def build(self, device: Optional[int] = None) -> SeqRetargeting:
        from .optimizer import DexPilotOptimizer, PositionOptimizer, VectorOptimizer

        model = KinematicModel.load(self.urdf_path, add_dummy_free_joints=self.add_dummy_free_joint)
        robot = RobotWrapper(model)

        # with a free-flying base the six dummy joints are optimised as well (retargeting_config.py:190-191)
        if self.add_dummy_free_joint and self.target_joint_names is not None:
            self.target_joint_names = DUMMY_JOINT_NAMES + self.target_joint_names
        joints = robot.dof_joint_names if self.target_joint_names is None else self.target_joint_names

        common = dict(target_link_human_indices=self.target_link_human_indices, device=device)
        loss = dict(norm_delta=self.normal_delta, huber_delta=self.huber_delta)
        if self.type == "position":
            optimizer = PositionOptimizer(robot, joints, target_link_names=self.target_link_names, **loss, **common)
        elif self.type == "vector":
            optimizer = VectorOptimizer(robot, joints, target_origin_link_names=self.target_origin_link_names,
                                        target_task_link_names=self.target_task_link_names, scaling=self.scaling_factor,
                                        **loss, **common)
        else:  # dexpilot: huber_delta / normal_delta are NOT forwarded (retargeting_config.py:218-228)
            optimizer = DexPilotOptimizer(robot, joints, finger_tip_link_names=self.finger_tip_link_names,
                                          wrist_link_name=self.wrist_link_name, scaling=self.scaling_factor,
                                          project_dist=self.project_dist, escape_dist=self.escape_dist, **common)

        has_mimic, sources, mimics, multipliers, offsets = parse_mimic_joint(model)
        if has_mimic and not self.ignore_mimic_joint:
            optimizer.set_kinematic_adaptor(MimicJointKinematicAdaptor(
                robot, target_joint_names=joints, source_joint_names=sources, mimic_joint_names=mimics,
                multipliers=multipliers, offsets=offsets))
        smoothing = LPFilter(self.low_pass_alpha) if 0 <= self.low_pass_alpha <= 1 else None
        return SeqRetargeting(optimizer, has_joint_limits=self.has_joint_limits, lp_filter=smoothing)


def get_retargeting_config(config_path: Union[str, Path]) -> RetargetingConfig:
    return RetargetingConfig.load_from_file(config_path)


def parse_mimic_joint(robot_model: KinematicModel) -> Tuple[bool, List[str], List[str], List[float], List[float]]:
    source_names, mimic_names, multipliers, offsets = robot_model.mimic_joints()
    return len(mimic_names) > 0, source_names, mimic_names, multipliers, offsets
