"""The C restatement of the oracle's kinematics (oracle/c/orc.c) equals the pure-Python one (oracle/robot.py)."""
import numpy as np
import pytest

from helpers import ROBOTS
from oracle.robot import OracleRobot, build_c_kinematics

ALL_ROBOTS = sorted(p.stem for p in ROBOTS.glob("*.json"))


def test_library_builds():
    assert build_c_kinematics().exists()


@pytest.mark.parametrize("name", ALL_ROBOTS)
@pytest.mark.parametrize("dummy", [False, True])
def test_c_matches_python(name, dummy):
    py = OracleRobot(str(ROBOTS / f"{name}.json"), dummy, use_c=False)
    cc = OracleRobot(str(ROBOTS / f"{name}.json"), dummy, use_c=True)
    assert py._c is None and cc._c is not None
    rng = np.random.RandomState(7)
    q = rng.uniform(py.joint_limits[:, 0], py.joint_limits[:, 1])
    py.compute_forward_kinematics(q)
    cc.compute_forward_kinematics(q)
    ids = list(range(len(py.link_names)))
    for i in ids:
        np.testing.assert_allclose(cc.get_link_pose(i), py.get_link_pose(i), atol=1e-14)
    np.testing.assert_allclose(cc.link_positions(ids), py.link_positions(ids), atol=1e-14)
    np.testing.assert_allclose(cc.link_jacobians(ids), py.link_jacobians(ids), atol=1e-14)
    g = rng.randn(len(ids), 3)
    np.testing.assert_allclose(cc.link_position_hessian_contraction(ids, g), py.link_position_hessian_contraction(ids, g), atol=1e-13)
