"""What the reference's tests/test_retargeting_config.py checks, against the drop-in (CPU: building a retargeting object
compiles the robot table but touches no GPU): every packaged config file builds (a superset of the 17 the reference lists),
configs given as dicts build -- one position config with explicit target joints on a mimic hand, and a list mixing a vector
config with a `DexPilot` one (mixed-case type, defaults for everything optional) -- and switching the free-flying base on adds
six leading dummy joints to the robot and to the optimised set."""
import pytest

from helpers import ROBOTS
from dex_retargeting_b200.constants import config_root
from dex_retargeting_b200.retargeting_config import RetargetingConfig
from dex_retargeting_b200.seq_retarget import SeqRetargeting

ALL_CONFIGS = sorted(p.relative_to(config_root()).as_posix() for p in config_root().glob("*/*.yml"))
OFFLINE_HANDS = [c for c in ALL_CONFIGS if c.startswith("offline/") and c.endswith("_right.yml")]


@pytest.fixture(autouse=True)
def _robots():
    RetargetingConfig.set_default_urdf_dir(str(RetargetingConfig.packaged_urdf_dir()))


def test_the_reference_lists_are_covered():
    must = ["teleop/allegro_hand_right.yml", "teleop/allegro_hand_left.yml", "teleop/shadow_hand_right.yml",
            "teleop/schunk_svh_hand_right.yml", "teleop/leap_hand_right.yml", "teleop/ability_hand_right.yml",
            "teleop/ability_hand_left.yml", "offline/allegro_hand_right.yml", "offline/shadow_hand_right.yml",
            "offline/schunk_svh_hand_right.yml", "offline/leap_hand_right.yml", "offline/ability_hand_right.yml",
            "teleop/allegro_hand_right_dexpilot.yml", "teleop/allegro_hand_left_dexpilot.yml",
            "teleop/shadow_hand_right_dexpilot.yml", "teleop/schunk_svh_hand_right_dexpilot.yml", "teleop/leap_hand_right_dexpilot.yml"]
    assert set(must) <= set(ALL_CONFIGS) and len(ALL_CONFIGS) == 39


@pytest.mark.parametrize("rel", ALL_CONFIGS)
def test_config_file_builds(rel):
    assert isinstance(RetargetingConfig.load_from_file(config_root() / rel).build(), SeqRetargeting)


def test_single_dict_config():
    cfg = dict(type="position", urdf_path="ability_hand/ability_hand_right.urdf", wrist_link_name="base_link",
               target_joint_names=["index_q1", "middle_q1", "pinky_q1", "ring_q1", "thumb_q1", "thumb_q2"],
               target_link_names=["thumb_tip", "index_tip", "middle_tip", "ring_tip", "pinky_tip"],
               target_link_human_indices=[4, 8, 12, 16, 20], low_pass_alpha=1)
    retargeting = RetargetingConfig.from_dict(cfg).build()
    assert isinstance(retargeting, SeqRetargeting)
    assert retargeting.optimizer.retargeting_type == "POSITION" and retargeting.optimizer.opt_dof == 6


def test_list_of_dict_configs_with_mixed_types():
    tips = ["link_15.0_tip", "link_3.0_tip", "link_7.0_tip", "link_11.0_tip"]
    cfgs = [dict(type="vector", urdf_path="allegro_hand/allegro_hand_right.urdf", wrist_link_name="wrist", target_joint_names=None,
                 target_origin_link_names=["wrist"] * 4, target_task_link_names=tips, scaling_factor=1.6,
                 target_link_human_indices=[[0] * 4, [4, 8, 12, 16]], low_pass_alpha=0.2),
            dict(type="DexPilot", urdf_path="leap_hand/leap_hand_right.urdf", wrist_link_name="base", target_joint_names=None,
                 finger_tip_link_names=[f"{f}_tip_head" for f in ("thumb", "index", "middle", "ring")], scaling_factor=1.6,
                 low_pass_alpha=0.2)]
    built = [RetargetingConfig.from_dict(c).build() for c in cfgs]
    assert all(isinstance(r, SeqRetargeting) for r in built)
    assert [r.optimizer.retargeting_type for r in built] == ["VECTOR", "DEXPILOT"]


@pytest.mark.parametrize("rel", OFFLINE_HANDS)
def test_free_flying_base_adds_six_leading_joints(rel):
    def build(flag):
        return RetargetingConfig.load_from_file(config_root() / rel, {"add_dummy_free_joint": flag}).build()

    fixed, flying = build(False), build(True)
    assert flying.optimizer.robot.dof == fixed.optimizer.robot.dof + 6
    assert flying.joint_limits.shape == (len(fixed.optimizer.target_joint_names) + 6, 2)
    assert all("dummy" in n for n in flying.optimizer.robot.dof_joint_names[:6])
