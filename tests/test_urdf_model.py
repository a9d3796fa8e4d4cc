"""Robot description reader, pinocchio-compatible DoF order, host FK vs the oracle's independent FK."""
import json

import numpy as np
import pytest

from helpers import ROBOTS
from dex_retargeting_b200.robot_wrapper import RobotWrapper
from dex_retargeting_b200.urdf import DUMMY_JOINT_NAMES, KinematicModel, rpy_to_matrix
from oracle.robot import OracleRobot

# predicted pinocchio DoF order (depth first, siblings by joint name)
EXPECTED_ORDER = {
    "allegro_hand_right": ["joint_0.0", "joint_1.0", "joint_2.0", "joint_3.0", "joint_12.0", "joint_13.0", "joint_14.0",
                           "joint_15.0", "joint_4.0", "joint_5.0", "joint_6.0", "joint_7.0", "joint_8.0", "joint_9.0",
                           "joint_10.0", "joint_11.0"],
    "leap_hand_right": ["1", "0", "2", "3", "12", "13", "14", "15", "5", "4", "6", "7", "9", "8", "10", "11"],
    "shadow_hand_right": ["WRJ2", "WRJ1", "FFJ4", "FFJ3", "FFJ2", "FFJ1", "LFJ5", "LFJ4", "LFJ3", "LFJ2", "LFJ1", "MFJ4",
                          "MFJ3", "MFJ2", "MFJ1", "RFJ4", "RFJ3", "RFJ2", "RFJ1", "THJ5", "THJ4", "THJ3", "THJ2", "THJ1"],
    "ability_hand_right": ["index_q1", "index_q2", "middle_q1", "middle_q2", "pinky_q1", "pinky_q2", "ring_q1", "ring_q2",
                           "thumb_q1", "thumb_q2"],
}
ALL_ROBOTS = sorted(p.stem for p in ROBOTS.glob("*.json"))


def test_fixture_inventory():
    assert len(ALL_ROBOTS) == 13  # 6 hands x left/right + panda gripper


@pytest.mark.parametrize("name", sorted(EXPECTED_ORDER))
def test_pinocchio_dof_order(name):
    m = KinematicModel.load(ROBOTS / f"{name}.json")
    assert m.dof_joint_names == EXPECTED_ORDER[name]


def test_dummy_joints_first_and_limits():
    # reference tests/test_retargeting_config.py:106-125: exactly 6 more DoFs, named *dummy*, first in order
    base = KinematicModel.load(ROBOTS / "shadow_hand_right.json")
    m = KinematicModel.load(ROBOTS / "shadow_hand_right.json", add_dummy_free_joints=True)
    assert m.dof == base.dof + 6
    assert m.dof_joint_names[:6] == DUMMY_JOINT_NAMES
    assert all("dummy" in n for n in m.dof_joint_names[:6])
    np.testing.assert_allclose(m.joint_limits[:3], [[-5, 5]] * 3)
    np.testing.assert_allclose(m.joint_limits[3:6], [[-2 * np.pi, 2 * np.pi]] * 3)
    assert list(m.joint_type[:6]) == [1, 1, 1, 0, 0, 0]
    assert m.joint_depth.max() == 13


def test_rpy_convention():
    # URDF fixed-axis rpy: R = Rz(yaw) Ry(pitch) Rx(roll)
    r, p, y = 0.3, -0.7, 1.1
    Rx = np.array([[1, 0, 0], [0, np.cos(r), -np.sin(r)], [0, np.sin(r), np.cos(r)]])
    Ry = np.array([[np.cos(p), 0, np.sin(p)], [0, 1, 0], [-np.sin(p), 0, np.cos(p)]])
    Rz = np.array([[np.cos(y), -np.sin(y), 0], [np.sin(y), np.cos(y), 0], [0, 0, 1]])
    np.testing.assert_allclose(rpy_to_matrix([r, p, y]), Rz @ Ry @ Rx, atol=1e-15)


@pytest.mark.parametrize("name", ALL_ROBOTS)
@pytest.mark.parametrize("dummy", [False, True])
def test_fk_and_jacobian_against_oracle(name, dummy):
    """Product host FK (folded joint table) == oracle FK (4x4 products over every URDF joint);
    product LOCAL frame Jacobian rotated to world == oracle world Jacobian == finite differences."""
    m = KinematicModel.load(ROBOTS / f"{name}.json", add_dummy_free_joints=dummy)
    w = RobotWrapper(m)
    o = OracleRobot(str(ROBOTS / f"{name}.json"), dummy)
    assert o.dof_joint_names == m.dof_joint_names
    np.testing.assert_allclose(o.joint_limits, m.joint_limits)
    rng = np.random.RandomState(3)
    q = rng.uniform(m.joint_limits[:, 0], m.joint_limits[:, 1])
    w.compute_forward_kinematics(q)
    o.compute_forward_kinematics(q)
    for ln in m.link_names:
        np.testing.assert_allclose(w.get_link_pose(w.get_link_index(ln)), o.get_link_pose(o.get_link_index(ln)), atol=1e-12)
    probe = m.link_names[-1]
    Jo = o.link_jacobians([o.get_link_index(probe)])[0]
    lid = w.get_link_index(probe)
    Jl = w.compute_single_link_local_jacobian(q, lid)
    Jw = w.get_link_pose(lid)[:3, :3] @ Jl[:3]
    np.testing.assert_allclose(Jw, Jo, atol=1e-12)
    h = 1e-6
    for i in range(m.dof):
        qp, qm = q.copy(), q.copy()
        qp[i] += h
        qm[i] -= h
        o.compute_forward_kinematics(qp)
        pp = o.link_positions([o.get_link_index(probe)])[0]
        o.compute_forward_kinematics(qm)
        pm = o.link_positions([o.get_link_index(probe)])[0]
        np.testing.assert_allclose((pp - pm) / (2 * h), Jo[:, i], atol=2e-8)


def test_pose_inverse_and_errors():
    w = RobotWrapper(ROBOTS / "allegro_hand_right.json")
    w.compute_forward_kinematics(np.full(w.dof, 0.2))
    lid = w.get_link_index("link_3.0_tip")
    np.testing.assert_allclose(w.get_link_pose(lid) @ w.get_link_pose_inv(lid), np.eye(4), atol=1e-14)
    with pytest.raises(ValueError):
        w.get_link_index("no_such_link")
    with pytest.raises(NotImplementedError):
        RobotWrapper(ROBOTS / "allegro_hand_right.json", use_visual=True)


def test_special_joint_rejected(tmp_path):
    d = json.loads((ROBOTS / "panda_gripper_glb.json").read_text())
    d["joints"][1]["type"] = "continuous"  # nq != nv in pinocchio -> robot_wrapper.py:22-23 raises
    p = tmp_path / "bad.json"
    p.write_text(json.dumps(d))
    with pytest.raises(NotImplementedError):
        KinematicModel.load(p)


def test_urdf_xml_reader_roundtrip(tmp_path):
    """The XML reader and the JSON fixtures describe the same model (write a URDF from the fixture)."""
    d = json.loads((ROBOTS / "schunk_svh_hand_right.json").read_text())
    lines = ['<?xml version="1.0"?>', f'<robot name="{d["name"]}">']
    for ln in d["links"]:
        lines.append(f'  <link name="{ln}"/>')
    for j in d["joints"]:
        lines.append(f'  <joint name="{j["name"]}" type="{j["type"]}">')
        lines.append(f'    <parent link="{j["parent"]}"/><child link="{j["child"]}"/>')
        lines.append(f'    <origin xyz="{" ".join(map(repr, j["xyz"]))}" rpy="{" ".join(map(repr, j["rpy"]))}"/>')
        lines.append(f'    <axis xyz="{" ".join(map(repr, j["axis"]))}"/>')
        if "limit" in j:
            lines.append(f'    <limit lower="{j["limit"][0]!r}" upper="{j["limit"][1]!r}" effort="1" velocity="1"/>')
        if "mimic" in j:
            lines.append(f'    <mimic joint="{j["mimic"][0]}" multiplier="{j["mimic"][1]!r}" offset="{j["mimic"][2]!r}"/>')
        lines.append("  </joint>")
    lines.append("</robot>")
    p = tmp_path / "svh.urdf"
    p.write_text("\n".join(lines))
    a, b = KinematicModel.load(p), KinematicModel.from_dict(d)
    assert a.dof_joint_names == b.dof_joint_names
    np.testing.assert_allclose(a.joint_R, b.joint_R)
    np.testing.assert_allclose(a.joint_p, b.joint_p)
    assert a.mimic_joints() == b.mimic_joints()
    assert len(a.mimic_joints()[1]) == 11
