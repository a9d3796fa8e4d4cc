"""Host-side mirror of the reference's Python surface (no GPU): constructors, attributes, errors, state."""
import numpy as np
import pytest

from helpers import ROBOTS, build_product, config_dict, configs
from dex_retargeting_b200.constants import (HandType, RetargetingType, RobotName, ROBOT_NAMES, get_default_config_path,
                                            OPERATOR2MANO)
from dex_retargeting_b200.optimizer import DexPilotOptimizer, Optimizer, PositionOptimizer, VectorOptimizer
from dex_retargeting_b200.optimizer_utils import LPFilter
from dex_retargeting_b200.retargeting_config import RetargetingConfig
from dex_retargeting_b200.robot_wrapper import RobotWrapper
from dex_retargeting_b200.seq_retarget import SeqRetargeting, _intrinsic_xyz_from_matrix, _quat_to_matrix
from dex_retargeting_b200.urdf import DUMMY_JOINT_NAMES


def test_dexpilot_statics_match_reference_docstrings():
    # reference optimizer.py:411-412 and :434-438
    assert DexPilotOptimizer.generate_link_indices(4) == ([2, 3, 4, 3, 4, 4, 0, 0, 0, 0], [1, 1, 1, 2, 2, 3, 1, 2, 3, 4])
    proj, s2o, s2t, dist = DexPilotOptimizer.set_dexpilot_cache(4, 0.1, 0.2)
    assert proj.dtype == bool and proj.tolist() == [False] * 6
    assert (s2o, s2t) == ([1, 2, 2], [0, 0, 1])
    np.testing.assert_allclose(dist, [0.1, 0.1, 0.1, 0.2, 0.2, 0.2])
    for nf in (2, 3, 5):
        o, t = DexPilotOptimizer.generate_link_indices(nf)
        assert len(o) == len(t) == nf * (nf - 1) // 2 + nf


@pytest.mark.parametrize("key", sorted(configs()))
def test_build_every_config(key):
    # reference tests/test_retargeting_config.py:47-52
    seq = build_product(key)
    assert isinstance(seq, SeqRetargeting)
    opt = seq.optimizer
    kind = configs()[key]["type"].lower()
    assert opt.retargeting_type == kind.upper()
    assert isinstance(opt, {"vector": VectorOptimizer, "position": PositionOptimizer, "dexpilot": DexPilotOptimizer}[kind])
    assert seq.last_qpos.dtype == np.float32 and seq.last_qpos.shape == (opt.opt_dof,)
    np.testing.assert_allclose(seq.last_qpos, seq.joint_limits.mean(1), rtol=1e-6)
    assert len(opt.fixed_joint_names) == len(opt.idx_pin2fixed)
    assert seq.joint_names == opt.robot.dof_joint_names
    if kind == "dexpilot":  # reference quirk: huber/normal delta of the config are not forwarded
        assert (opt.huber_delta, opt.norm_delta) == (0.03, 4e-3)


def test_dummy_joint_config():
    # reference tests/test_retargeting_config.py:106-125
    base = build_product("offline/shadow_hand_right", dict(add_dummy_free_joint=False))
    seq = build_product("offline/shadow_hand_right")
    assert seq.optimizer.robot.dof == base.optimizer.robot.dof + 6
    assert seq.optimizer.robot.dof_joint_names[:6] == DUMMY_JOINT_NAMES
    assert seq.optimizer.has_free_joint and not base.optimizer.has_free_joint
    pg = build_product("offline/panda_gripper")
    assert pg.optimizer.target_joint_names[:6] == DUMMY_JOINT_NAMES and pg.optimizer.opt_dof == 7


def test_config_validation_errors():
    RetargetingConfig.set_default_urdf_dir(str(ROBOTS))
    good = config_dict("teleop/allegro_hand_right")
    with pytest.raises(ValueError, match="type must be one of"):
        RetargetingConfig.from_dict({**good, "type": "nope"})
    with pytest.raises(ValueError, match="dim mismatch"):
        RetargetingConfig.from_dict({**good, "target_task_link_names": ["link_15.0_tip"]})
    with pytest.raises(ValueError, match="link indices dim mismatch"):
        RetargetingConfig.from_dict({**good, "target_link_human_indices": [[0, 0], [4, 8]]})
    with pytest.raises(ValueError, match="does not exist"):
        RetargetingConfig.from_dict({**good, "urdf_path": "missing.urdf"})
    with pytest.raises(ValueError, match="not exists"):
        RetargetingConfig.set_default_urdf_dir("/no/such/dir")
    with pytest.raises(ValueError, match="not a link name"):
        RetargetingConfig.from_dict({**good, "target_task_link_names": ["a", "b", "c", "d"]}).build()
    pos = config_dict("offline/allegro_hand_right")
    with pytest.raises(ValueError, match="target_link_names"):
        RetargetingConfig.from_dict({k: v for k, v in pos.items() if k != "target_link_names"})
    dp = config_dict("teleop/leap_hand_right_dexpilot")
    with pytest.raises(ValueError, match="finger_tip_link_names"):
        RetargetingConfig.from_dict({k: v for k, v in dp.items() if k != "wrist_link_name"})


def test_optimizer_errors_and_mimic_target_clash():
    robot = RobotWrapper(ROBOTS / "schunk_svh_hand_right.json")
    with pytest.raises(ValueError, match="does not appear to be in robot XML"):
        Optimizer(robot, ["bogus"], np.zeros((2, 1)))
    with pytest.raises(ValueError, match="2 to 5 fingers"):
        DexPilotOptimizer(robot, robot.dof_joint_names, ["thtip"], "right_hand_base_link")
    seq = build_product("teleop/allegro_hand_right")
    with pytest.raises(ValueError, match="Expect joint limits have shape"):
        seq.optimizer.set_joint_limit(np.zeros((3, 2)))
    with pytest.raises(ValueError, match="non_target_qpos"):
        seq.optimizer.retarget(np.zeros((4, 3)), np.zeros(2), np.zeros(16))
    # all joints as targets on a mimic hand -> adaptor refuses (kinematics_adaptor.py:63-70)
    cfg = config_dict("teleop/schunk_svh_hand_right")
    cfg.pop("target_joint_names")
    RetargetingConfig.set_default_urdf_dir(str(ROBOTS))
    with pytest.raises(ValueError, match="Mimic joint should not be one of the target joints"):
        RetargetingConfig.from_dict(cfg).build()
    # ... unless mimic tags are ignored: then all 20 joints are optimised
    seq = RetargetingConfig.from_dict({**cfg, "ignore_mimic_joint": True}).build()
    assert seq.optimizer.opt_dof == 20 and seq.optimizer.adaptor is None


def test_fixed_joints_with_partial_targets():
    cfg = config_dict("teleop/allegro_hand_right")
    RetargetingConfig.set_default_urdf_dir(str(ROBOTS))
    names = ["joint_0.0", "joint_1.0", "joint_2.0", "joint_3.0"]
    seq = RetargetingConfig.from_dict({**cfg, "target_joint_names": names}).build()
    opt = seq.optimizer
    assert opt.opt_dof == 4 and len(opt.idx_pin2fixed) == 12
    t = opt.build_table()
    assert t.n_var == 4 and t.n_fixed == 12


def test_lp_filter_and_state_helpers():
    f = LPFilter(0.25)
    a = f.next(np.array([1.0, 2.0]))
    np.testing.assert_allclose(a, [1, 2])
    b = f.next(np.array([3.0, 2.0]))
    np.testing.assert_allclose(b, [1.5, 2.0])
    f.reset()
    assert not f.is_init and f.y is None
    seq = build_product("teleop/allegro_hand_right")
    q = np.arange(16, dtype=float) / 100
    seq.set_qpos(q)
    np.testing.assert_allclose(seq.get_qpos(), q)
    seq.reset()
    np.testing.assert_allclose(seq.last_qpos, seq.joint_limits.mean(1), rtol=1e-6)
    assert seq.low_pass_alpha == pytest.approx(0.2)
    assert build_product("teleop/allegro_hand_right", dict(low_pass_alpha=2.0)).filter is None


def test_default_config_paths_and_enums():
    assert len(ROBOT_NAMES) == 7
    p = get_default_config_path(RobotName.allegro, RetargetingType.vector, HandType.right)
    assert p.parts[-2:] == ("teleop", "allegro_hand_right.yml")
    p = get_default_config_path(RobotName.shadow, RetargetingType.position, HandType.left)
    assert p.parts[-2:] == ("offline", "shadow_hand_left.yml")
    p = get_default_config_path(RobotName.leap, RetargetingType.dexpilot, HandType.right)
    assert p.name == "leap_hand_right_dexpilot.yml"
    assert get_default_config_path(RobotName.panda, RetargetingType.dexpilot, HandType.left).name == "panda_gripper_dexpilot.yml"
    assert get_default_config_path(RobotName.panda, RetargetingType.position, HandType.left).parts[-2:] == ("offline", "panda_gripper.yml")
    for m in OPERATOR2MANO.values():
        np.testing.assert_allclose(m @ m.T, np.eye(3))


def test_warm_start_sets_dummy_joints():
    seq = build_product("offline/shadow_hand_right")
    rng = np.random.RandomState(0)
    quat = rng.randn(4)
    quat /= np.linalg.norm(quat)
    pos = np.array([0.3, -0.2, 0.5])
    seq.warm_start(pos, quat, HandType.right, is_mano_convention=False)
    assert seq.is_warm_started
    # FK at the warm-started pose: the wrist link (child of the last dummy joint) sits at the requested pose
    robot = seq.optimizer.robot
    q = np.zeros(robot.dof)
    q[:6] = seq.last_qpos[:6]
    robot.compute_forward_kinematics(q)
    wrist = robot.get_joint_parent_child_frames(DUMMY_JOINT_NAMES[5])[1]
    T = robot.get_link_pose(wrist)
    np.testing.assert_allclose(T[:3, 3], pos, atol=1e-6)
    np.testing.assert_allclose(T[:3, :3], _quat_to_matrix(quat), atol=1e-6)
    with pytest.raises(ValueError):
        seq.warm_start(np.zeros(2), quat)


def test_euler_extraction_roundtrip():
    rng = np.random.RandomState(1)
    for _ in range(20):
        a, b, c = rng.uniform(-1.5, 1.5, 3)
        Rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
        Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
        Rz = np.array([[np.cos(c), -np.sin(c), 0], [np.sin(c), np.cos(c), 0], [0, 0, 1]])
        np.testing.assert_allclose(_intrinsic_xyz_from_matrix(Rx @ Ry @ Rz), [a, b, c], atol=1e-12)


def test_packaged_default_configs_load_with_packaged_urdfs():
    """The 39 packaged YAML files (same names / schema as the reference package) load and build from the kinematics-only
    URDFs shipped under dex_retargeting_b200/assets (the real URDF reader, not the JSON fixtures); a path that does not
    exist raises like the reference (retargeting_config.py:131-132) instead of falling back to anything."""
    RetargetingConfig.set_default_urdf_dir(str(RetargetingConfig.packaged_urdf_dir()))
    with pytest.raises(ValueError, match="does not exist"):
        RetargetingConfig.load_from_file(get_default_config_path(RobotName.allegro, RetargetingType.vector, HandType.right),
                                         override=dict(urdf_path="allegro_hand/allegro_hand_rigth.urdf"))
    n = 0
    for robot in ROBOT_NAMES:
        for rtype in RetargetingType:
            for hand in HandType:
                if robot is RobotName.ability and rtype is RetargetingType.dexpilot and False:
                    continue
                path = get_default_config_path(robot, rtype, hand)
                assert path.exists(), path
                seq = RetargetingConfig.load_from_file(path).build()
                assert seq.optimizer.retargeting_type == rtype.name.upper()
                n += 1
    assert n == 42  # 7 robots x 3 types x 2 hands (the gripper's two hands share a file)
    seq = RetargetingConfig.load_from_file(get_default_config_path(RobotName.allegro, RetargetingType.vector, HandType.right),
                                           override=dict(scaling_factor=1.0, low_pass_alpha=0)).build()
    assert seq.optimizer.scaling == 1.0 and seq.filter.alpha == 0


def test_step_tol_default_and_env_override(monkeypatch):
    """`Optimizer.step_tol` defaults to 1e-5; DEXR_STEP_TOL (A/B switch) changes the default of
    optimizers constructed afterwards and reaches the launch parameters."""
    RetargetingConfig.set_default_urdf_dir(str(RetargetingConfig.packaged_urdf_dir()))
    path = get_default_config_path(RobotName.allegro, RetargetingType.vector, HandType.right)
    monkeypatch.delenv("DEXR_STEP_TOL", raising=False)
    opt = RetargetingConfig.load_from_file(path).build().optimizer
    assert opt.step_tol == 1e-5 and abs(opt.params().tol - 1e-5) < 1e-12
    monkeypatch.setenv("DEXR_STEP_TOL", "1e-4")
    opt = RetargetingConfig.load_from_file(path).build().optimizer
    assert opt.step_tol == 1e-4 and abs(opt.params().tol - 1e-4) < 1e-11
