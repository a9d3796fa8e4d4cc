"""Names shared by the whole package: robots, retargeting types, hand sides, the operator -> MANO frame change and the
lookup of the packaged default configurations.

Public names and values are those of src/dex_retargeting/constants.py:7-87 (user code imports them); the default
configuration files live under `dex_retargeting_b200/configs/{teleop,offline}/` with the reference package's file names, or
under the directory named by $DEX_RETARGETING_CONFIG_DIR.
"""
import enum
import os
from pathlib import Path
from typing import Optional

import numpy as np

# members are numbered from 1 in this order, like enum.auto() in the reference
RobotName = enum.Enum("RobotName", "allegro shadow svh leap ability inspire panda")
RetargetingType = enum.Enum("RetargetingType", "vector position dexpilot")  # teleop | offline hand-object data | teleop with a finger-closing prior
HandType = enum.Enum("HandType", "right left")

# file stem of each robot's URDF and configuration files
_STEMS = dict(allegro="allegro_hand", shadow="shadow_hand", svh="schunk_svh_hand", leap="leap_hand", ability="ability_hand",
              inspire="inspire_hand", panda="panda_gripper")
ROBOT_NAME_MAP = {RobotName[name]: stem for name, stem in _STEMS.items()}
ROBOT_NAMES = list(ROBOT_NAME_MAP)


def _signed_permutation(*rows: str) -> np.ndarray:
    """Rows given as signed axis names ("-z" = the row (0, 0, -1)): integer matrix of a frame change."""
    m = np.zeros((3, 3), dtype=int)
    for r, spec in enumerate(rows):
        m[r, "xyz".index(spec[-1])] = -1 if spec.startswith("-") else 1
    return m


# operator (wrist-frame estimate of the detector) -> MANO convention; the left hand mirrors the y axis of the right one
OPERATOR2MANO_RIGHT = _signed_permutation("-z", "-x", "+y")
OPERATOR2MANO_LEFT = _signed_permutation("-z", "+x", "-y")
OPERATOR2MANO = {HandType.right: OPERATOR2MANO_RIGHT, HandType.left: OPERATOR2MANO_LEFT}


def config_root() -> Path:
    env = os.environ.get("DEX_RETARGETING_CONFIG_DIR")
    return Path(env) if env else Path(__file__).parent / "configs"


def get_default_config_path(robot_name: RobotName, retargeting_type: RetargetingType, hand_type: HandType) -> Optional[Path]:
    """`<root>/<teleop|offline>/<stem>[_<hand>][_dexpilot].yml`; grippers have one hand-agnostic file per type."""
    parts = [ROBOT_NAME_MAP[robot_name]]
    if "gripper" not in parts[0]:
        parts.append(hand_type.name)
    if retargeting_type is RetargetingType.dexpilot:
        parts.append("dexpilot")
    folder = "offline" if retargeting_type is RetargetingType.position else "teleop"
    return config_root() / folder / ("_".join(parts) + ".yml")
