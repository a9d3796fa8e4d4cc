"""The reference's tests/test_retargeting_config.py protocol against the drop-in (CPU: building a retargeting object compiles
the robot table but touches no GPU): config files by path, a config from a dict, a list of dict configs with mixed types
(incl. the mixed-case `type: DexPilot`), and the free-flying-base override."""
import pytest
import yaml

from helpers import ROBOTS
from dex_retargeting_b200.constants import config_root
from dex_retargeting_b200.retargeting_config import RetargetingConfig
from dex_retargeting_b200.seq_retarget import SeqRetargeting

VECTOR = ["teleop/allegro_hand_right.yml", "teleop/allegro_hand_left.yml", "teleop/shadow_hand_right.yml",
          "teleop/schunk_svh_hand_right.yml", "teleop/leap_hand_right.yml", "teleop/ability_hand_right.yml",
          "teleop/ability_hand_left.yml"]
POSITION = ["offline/allegro_hand_right.yml", "offline/shadow_hand_right.yml", "offline/schunk_svh_hand_right.yml",
            "offline/leap_hand_right.yml", "offline/ability_hand_right.yml"]
DEXPILOT = ["teleop/allegro_hand_right_dexpilot.yml", "teleop/allegro_hand_left_dexpilot.yml", "teleop/shadow_hand_right_dexpilot.yml",
            "teleop/schunk_svh_hand_right_dexpilot.yml", "teleop/leap_hand_right_dexpilot.yml"]


@pytest.fixture(autouse=True)
def _urdf_dir():
    RetargetingConfig.set_default_urdf_dir(str(ROBOTS))


@pytest.mark.parametrize("config_path", VECTOR + POSITION + DEXPILOT)
def test_path_config_parsing(config_path):
    retargeting = RetargetingConfig.load_from_file(config_root() / config_path).build()
    assert isinstance(retargeting, SeqRetargeting)


def test_dict_config_parsing():
    cfg = yaml.safe_load("""
    type: position
    urdf_path: ability_hand/ability_hand_right.urdf
    wrist_link_name: "base_link"
    target_joint_names: ['index_q1', 'middle_q1', 'pinky_q1', 'ring_q1', 'thumb_q1', 'thumb_q2']
    target_link_names: ["thumb_tip", "index_tip", "middle_tip", "ring_tip", "pinky_tip"]
    target_link_human_indices: [4, 8, 12, 16, 20]
    low_pass_alpha: 1
    """)
    retargeting = RetargetingConfig.from_dict(cfg).build()
    assert isinstance(retargeting, SeqRetargeting)
    assert retargeting.optimizer.retargeting_type == "POSITION" and retargeting.optimizer.opt_dof == 6


def test_multi_dict_config_parsing():
    cfgs = yaml.safe_load("""
    - type: vector
      urdf_path: allegro_hand/allegro_hand_right.urdf
      wrist_link_name: "wrist"
      target_joint_names: null
      target_origin_link_names: ["wrist", "wrist", "wrist", "wrist"]
      target_task_link_names: ["link_15.0_tip", "link_3.0_tip", "link_7.0_tip", "link_11.0_tip"]
      scaling_factor: 1.6
      target_link_human_indices: [[0, 0, 0, 0], [4, 8, 12, 16]]
      low_pass_alpha: 0.2
    - type: DexPilot
      urdf_path: leap_hand/leap_hand_right.urdf
      wrist_link_name: "base"
      target_joint_names: null
      finger_tip_link_names: ["thumb_tip_head", "index_tip_head", "middle_tip_head", "ring_tip_head"]
      scaling_factor: 1.6
      low_pass_alpha: 0.2
    """)
    kinds = []
    for cfg in cfgs:
        retargeting = RetargetingConfig.from_dict(cfg).build()
        assert isinstance(retargeting, SeqRetargeting)
        kinds.append(retargeting.optimizer.retargeting_type)
    assert kinds == ["VECTOR", "DEXPILOT"]


@pytest.mark.parametrize("config_path", POSITION)
def test_add_dummy_joint(config_path):
    path = config_root() / config_path
    retargeting = RetargetingConfig.load_from_file(path, {"add_dummy_free_joint": False}).build()
    robot_dof = retargeting.optimizer.robot.dof
    active_dof = len(retargeting.optimizer.target_joint_names)
    retargeting = RetargetingConfig.load_from_file(path, {"add_dummy_free_joint": True}).build()
    robot = retargeting.optimizer.robot
    assert robot.dof == robot_dof + 6
    assert retargeting.joint_limits.shape == (active_dof + 6, 2)
    assert all("dummy" in n for n in robot.dof_joint_names[:6])
