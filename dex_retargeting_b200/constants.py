"""Enums, robot names, default config lookup, operator->MANO frames.

API mirror of src/dex_retargeting/constants.py:7-87 (reference).  The default configuration files
are looked up under `dex_retargeting_b200/configs/{teleop,offline}/` (same file names as the
reference package) -- or under the directory named by $DEX_RETARGETING_CONFIG_DIR.
"""
import enum
import os
from pathlib import Path
from typing import Optional

import numpy as np

OPERATOR2MANO_RIGHT = np.array([[0, 0, -1], [-1, 0, 0], [0, 1, 0]])
OPERATOR2MANO_LEFT = np.array([[0, 0, -1], [1, 0, 0], [0, -1, 0]])


class RobotName(enum.Enum):
    allegro = enum.auto()
    shadow = enum.auto()
    svh = enum.auto()
    leap = enum.auto()
    ability = enum.auto()
    inspire = enum.auto()
    panda = enum.auto()


class RetargetingType(enum.Enum):
    vector = enum.auto()    # teleoperation, no finger closing prior
    position = enum.auto()  # offline data, hand-object interaction
    dexpilot = enum.auto()  # teleoperation, finger closing prior


class HandType(enum.Enum):
    right = enum.auto()
    left = enum.auto()


ROBOT_NAME_MAP = {
    RobotName.allegro: "allegro_hand",
    RobotName.shadow: "shadow_hand",
    RobotName.svh: "schunk_svh_hand",
    RobotName.leap: "leap_hand",
    RobotName.ability: "ability_hand",
    RobotName.inspire: "inspire_hand",
    RobotName.panda: "panda_gripper",
}
ROBOT_NAMES = list(ROBOT_NAME_MAP.keys())


def config_root() -> Path:
    env = os.environ.get("DEX_RETARGETING_CONFIG_DIR")
    return Path(env) if env else Path(__file__).parent / "configs"


def get_default_config_path(robot_name: RobotName, retargeting_type: RetargetingType, hand_type: HandType) -> Optional[Path]:
    sub = "offline" if retargeting_type is RetargetingType.position else "teleop"
    stem = ROBOT_NAME_MAP[robot_name]
    if "gripper" not in stem:  # grippers have a single, hand-agnostic file
        stem = f"{stem}_{hand_type.name}"
    if retargeting_type is RetargetingType.dexpilot:
        stem += "_dexpilot"
    return config_root() / sub / f"{stem}.yml"


OPERATOR2MANO = {HandType.right: OPERATOR2MANO_RIGHT, HandType.left: OPERATOR2MANO_LEFT}
