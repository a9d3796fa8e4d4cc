"""tests/golden/reference_fk_vectors.npz (tests/tools/gen_reference_fk_vectors.py): global link poses computed by the
reference's OWN tree forward kinematics (yourdfpy.py build_tree / update_kinematics / _forward_kinematics_joint, executed
unmodified; rotation arithmetic from OpenCV and scipy) for all 13 hand URDFs, with and without the free-flying base.

The oracle's FK (oracle/robot.py, Python walk and C restatement) and the product's host-side RobotWrapper restate
pinocchio's forwardKinematics + updateFramePlacement (robot_wrapper.py:82-87), which is absent offline: this fixture is the
external opinion that pins row a9 of SURVEY.md section 8 -- every link, every joint type, mimic joints, dummy joints."""
from pathlib import Path

import numpy as np
import pytest

from helpers import GOLDEN, ROBOTS
from dex_retargeting_b200.robot_wrapper import RobotWrapper
from oracle.robot import OracleRobot

VEC = np.load(GOLDEN / "reference_fk_vectors.npz")
STEMS = [Path(str(p)).stem for p in VEC["urdfs"]]
TOL = 1e-12


def _full_q(robot_names, mimic_spec, act_names, q_act):
    """The pinocchio-order joint vector that puts the robot in the reference tree's configuration: actuated joints by
    name, mimic joints = multiplier * source + offset with the source looked up among the actuated joints
    (yourdfpy.py:1017-1030: a source that is not actuated counts as 0)."""
    src, mim, mul, off = mimic_spec
    by_name = dict(zip(act_names, q_act))
    for s, m, a, b in zip(src, mim, mul, off):
        by_name[m] = (by_name[s] if s in act_names else 0.0) * a + b
    return np.array([by_name[n] for n in robot_names])


def test_inventory():
    assert len(STEMS) >= 13
    for s in STEMS:
        for kind in ("plain", "dummy"):
            assert VEC[f"{s}/{kind}/poses"].shape[0] >= 4
            # proper rigid transforms (the generator's third-party rotations did their job)
            P = VEC[f"{s}/{kind}/poses"]
            R = P[..., :3, :3]
            np.testing.assert_allclose(R @ np.swapaxes(R, -1, -2), np.broadcast_to(np.eye(3), R.shape), atol=1e-13)
            np.testing.assert_allclose(np.linalg.det(R), 1.0, atol=1e-13)


@pytest.mark.parametrize("stem", STEMS)
@pytest.mark.parametrize("dummy", [False, True])
@pytest.mark.parametrize("use_c", [False, True])
def test_oracle_fk_matches_reference_tree_fk(stem, dummy, use_c):
    tag = f"{stem}/{'dummy' if dummy else 'plain'}"
    o = OracleRobot(str(ROBOTS / f"{stem}.json"), dummy, use_c=use_c)
    act = [str(n) for n in VEC[f"{tag}/actuated"]]
    links = [str(n) for n in VEC[f"{tag}/links"]]
    assert set(act) <= set(o.dof_joint_names)
    worst = 0.0
    for q_act, poses in zip(VEC[f"{tag}/q"], VEC[f"{tag}/poses"]):
        q = _full_q(o.dof_joint_names, o.mimic_spec(), act, q_act)
        o.compute_forward_kinematics(q)
        for name, T in zip(links, poses):
            worst = max(worst, float(np.abs(o.get_link_pose(o.get_link_index(name)) - T).max()))
    assert worst < TOL, f"{tag}: oracle FK differs from the reference's tree FK by {worst:.3e}"


@pytest.mark.parametrize("stem", STEMS)
@pytest.mark.parametrize("dummy", [False, True])
def test_product_host_fk_matches_reference_tree_fk(stem, dummy):
    """The pinocchio-free RobotWrapper (fixed joints folded into a flat table -- a different algorithm from the tree walk)."""
    tag = f"{stem}/{'dummy' if dummy else 'plain'}"
    r = RobotWrapper(str(ROBOTS / f"{stem}.json"), add_dummy_free_joints=dummy)
    act = [str(n) for n in VEC[f"{tag}/actuated"]]
    links = [str(n) for n in VEC[f"{tag}/links"]]
    worst = 0.0
    for q_act, poses in zip(VEC[f"{tag}/q"], VEC[f"{tag}/poses"]):
        q = _full_q(r.dof_joint_names, r.kin.mimic_joints(), act, q_act)
        r.compute_forward_kinematics(q)
        for name, T in zip(links, poses):
            worst = max(worst, float(np.abs(r.get_link_pose(r.get_link_index(name)) - T).max()))
            np.testing.assert_allclose(r.get_link_pose_inv(r.get_link_index(name)) @ T, np.eye(4), atol=1e-11)
    assert worst < TOL, f"{tag}: product host FK differs from the reference's tree FK by {worst:.3e}"


@pytest.mark.parametrize("stem", ["allegro_hand_right", "shadow_hand_right", "schunk_svh_hand_right", "panda_gripper_glb"])
def test_jacobian_is_the_derivative_of_the_pinned_fk(stem):
    """With FK pinned to the reference's tree walk, the frame Jacobian (robot_wrapper.py:93-95) is pinned by differentiating
    it: central differences of the oracle FK against its analytic world-aligned linear Jacobian."""
    o = OracleRobot(str(ROBOTS / f"{stem}.json"), True)
    rng = np.random.RandomState(3)
    lim = o.joint_limits
    q = rng.uniform(np.maximum(lim[:, 0], -3), np.minimum(lim[:, 1], 3))
    ids = list(range(len(o.link_names)))
    o.compute_forward_kinematics(q)
    J = o.link_jacobians(ids)
    h = 1e-6
    for j in range(o.dof):
        dq = np.zeros(o.dof)
        dq[j] = h
        o.compute_forward_kinematics(q + dq)
        pp = o.link_positions(ids)
        o.compute_forward_kinematics(q - dq)
        pm = o.link_positions(ids)
        np.testing.assert_allclose(J[:, :, j], (pp - pm) / (2 * h), atol=2e-8)
