"""tests/golden/reference_urdf_vectors.npz (tests/tools/gen_reference_urdf_vectors.py): what the reference's OWN URDF code
(yourdfpy.py: _parse_robot / _parse_joint / _parse_origin / _parse_axis / _parse_limit / _parse_mimic / _add_dummy_joints,
executed with lxml / anytree / pytransform3d shimmed) produced for every hand URDF its configs use.  The product's URDF
reader, the committed joint-tree fixtures the GPU tests run on, and the oracle's robot model must say the same."""
from pathlib import Path

import numpy as np
import pytest

from helpers import GOLDEN, ROBOTS
from dex_retargeting_b200.urdf import DUMMY_JOINT_NAMES, KinematicModel, rpy_to_matrix
from oracle.robot import OracleRobot

VEC = np.load(GOLDEN / "reference_urdf_vectors.npz")
STEMS = [Path(str(p)).stem for p in VEC["urdfs"]]


def test_inventory():
    assert len(STEMS) >= 13 and {"allegro_hand_right", "shadow_hand_right", "schunk_svh_hand_right", "panda_gripper_glb"} <= set(STEMS)
    for s in STEMS:
        assert (ROBOTS / f"{s}.json").exists(), s


@pytest.mark.parametrize("stem", STEMS)
@pytest.mark.parametrize("dummy", [False, True])
def test_joint_tree_matches_reference_parser(stem, dummy):
    tag = f"{stem}/{'dummy' if dummy else 'plain'}"
    m = KinematicModel.load(ROBOTS / f"{stem}.json", add_dummy_free_joints=dummy)
    names = [str(n) for n in VEC[f"{tag}/joint_names"]]
    assert [j.name for j in m.joints] == names                      # file order, dummy chain first (yourdfpy.py:1985)
    assert list(m.urdf_link_names) == [str(n) for n in VEC[f"{tag}/links"]]
    for i, j in enumerate(m.joints):
        assert j.type == str(VEC[f"{tag}/joint_types"][i]), j.name
        assert j.parent == str(VEC[f"{tag}/joint_parent"][i]) and j.child == str(VEC[f"{tag}/joint_child"][i]), j.name
        T = np.eye(4)
        T[:3, :3], T[:3, 3] = rpy_to_matrix(j.rpy), j.xyz            # yourdfpy.py:1375-1387
        np.testing.assert_allclose(T, VEC[f"{tag}/joint_origin"][i], atol=1e-15, err_msg=j.name)
        if j.type != "fixed":
            np.testing.assert_allclose(j.axis, VEC[f"{tag}/joint_axis"][i], atol=0, err_msg=j.name)  # :1631-1643
            lo, hi = VEC[f"{tag}/joint_limit"][i]
            assert (j.lower, j.upper) == (lo, hi), j.name                                             # :1652-1661
    src, mim, mul, off = m.mimic_joints()                                                               # :1107-1115
    assert mim == [str(n) for n in VEC[f"{tag}/mimic_joint"]] and src == [str(n) for n in VEC[f"{tag}/mimic_source"]]
    np.testing.assert_array_equal(mul, VEC[f"{tag}/mimic_mult"])
    np.testing.assert_array_equal(off, VEC[f"{tag}/mimic_off"])
    if dummy:
        assert names[:6] == DUMMY_JOINT_NAMES
        np.testing.assert_allclose(VEC[f"{tag}/joint_limit"][:3], [[-5, 5]] * 3)
        np.testing.assert_allclose(VEC[f"{tag}/joint_limit"][3:6], [[-2 * np.pi, 2 * np.pi]] * 3)


@pytest.mark.parametrize("rel", [str(p) for p in VEC["urdfs"]])
def test_packaged_urdf_equals_fixture(rel):
    """The kinematics-only URDFs shipped under dex_retargeting_b200/assets (what bench.py and the default configs load through
    the real URDF reader) describe exactly the joint trees held to the reference parser above."""
    from dex_retargeting_b200.retargeting_config import RetargetingConfig

    a = KinematicModel.from_urdf(RetargetingConfig.packaged_urdf_dir() / rel)
    b = KinematicModel.load(ROBOTS / f"{Path(rel).stem}.json")
    assert a.to_dict() == b.to_dict()


@pytest.mark.parametrize("stem", STEMS)
@pytest.mark.parametrize("dummy", [False, True])
def test_dof_set_and_oracle_model_match_reference(stem, dummy):
    """The movable, non-mimic joints are exactly the reference's `actuated_joint_names`; product and oracle agree on the DoF
    order (pinocchio's, restated -- that ORDER is not something yourdfpy knows, only the set) and on the joint limits."""
    tag = f"{stem}/{'dummy' if dummy else 'plain'}"
    m = KinematicModel.load(ROBOTS / f"{stem}.json", add_dummy_free_joints=dummy)
    o = OracleRobot(str(ROBOTS / f"{stem}.json"), dummy)
    assert list(m.dof_joint_names) == list(o.dof_joint_names)
    mimic = {str(n) for n in VEC[f"{tag}/mimic_joint"]}
    assert set(m.dof_joint_names) - mimic == {str(n) for n in VEC[f"{tag}/actuated"]}
    by_name = {str(n): VEC[f"{tag}/joint_limit"][i] for i, n in enumerate(VEC[f"{tag}/joint_names"])}
    for k, n in enumerate(m.dof_joint_names):
        np.testing.assert_array_equal(m.joint_limits[k], by_name[n])
        np.testing.assert_array_equal(o.joint_limits[k], by_name[n])
    root = str(VEC[f"{tag}/base_link"])
    assert o.root == root
