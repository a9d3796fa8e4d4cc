#!/usr/bin/env python
"""Generate tests/golden/reference_fk_vectors.npz: link poses from the REFERENCE'S OWN tree forward kinematics
(src/dex_retargeting/yourdfpy.py, unmodified: URDF.load -> build_tree :1862-1896, update_kinematics :1898-1934,
_forward_kinematics_joint :1014-1050, get_link_global_transform :1936-1939) on every hand URDF, with and without the
dummy free joints, at seeded random configurations.

Why: the hot path's FK is pinocchio's (robot_wrapper.py:82-87), which is absent offline, and the oracle restates it
(oracle/robot.py).  This gives the oracle's FK a second opinion that was not written by us: the tree walk, the
`origin @ joint_motion` composition order, the mimic evaluation and the actuated-joint bookkeeping below are the
reference's code, and the rotation arithmetic underneath comes from two third-party libraries (OpenCV's Rodrigues for
axis-angle, scipy's Rotation for the URDF rpy convention) instead of our own formulas.

Build container only (needs /root/reference, cv2, scipy).  Stand-ins for the absent packages (lxml, anytree: as in
gen_reference_urdf_vectors.py).  pytransform3d:
  * `transform_from(R, p)`                      -> 4x4 from R, p
  * `matrix_from_euler(e, 0, 1, 2, extrinsic)`  -> scipy Rotation.from_euler("xyz", e) (lower case = extrinsic x, y, z,
                                                   the URDF rpy convention, i.e. Rz(e2) Ry(e1) Rx(e0))
  * `matrix_from_axis_angle([axis, q])`         -> cv2.Rodrigues(axis * q), RETURNED AS A 4x4 HOMOGENEOUS MATRIX.
    The real function returns 3x3 and the reference multiplies it onto the 4x4 origin (`origin @ R`, yourdfpy.py:1044),
    which raises for every revolute joint -- the reference's tree FK is dead code the hot path never calls.  Returning
    the homogeneous embedding of the same rotation is the one repair that makes the line well formed (the prismatic
    branch two lines above does exactly that with `pt.transform_from(np.eye(3), q * axis)`).  The reference's code is
    not edited.

Recorded per URDF x {plain, dummy}: the actuated joint names in the reference's order, Q configurations (uniform inside
the joint limits), the link names and the [Q, links, 4, 4] global link poses.  tests/test_reference_fk_vectors.py holds
oracle/robot.py (Python and C), and the product's host RobotWrapper, to them.

Usage: python tests/tools/gen_reference_fk_vectors.py [/root/reference]
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT / "tests" / "tools"))
from helpers import configs  # noqa: E402
from gen_reference_urdf_vectors import _install_shims  # noqa: E402

N_Q = 6


def _third_party_rotations():
    import cv2
    from scipy.spatial.transform import Rotation

    rot = sys.modules["pytransform3d.rotations"]

    def matrix_from_axis_angle(a):
        a = np.asarray(a, dtype=np.float64)
        R, _ = cv2.Rodrigues((a[:3] / np.linalg.norm(a[:3]) * a[3]).reshape(3, 1))
        T = np.eye(4)
        T[:3, :3] = R
        return T

    def matrix_from_euler(e, i, j, k, extrinsic):
        assert (i, j, k, extrinsic) == (0, 1, 2, True)
        return Rotation.from_euler("xyz", np.asarray(e, dtype=np.float64)).as_matrix()

    rot.matrix_from_axis_angle, rot.matrix_from_euler = matrix_from_axis_angle, matrix_from_euler


def main():
    ref_root = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    _install_shims(ref_root)
    _third_party_rotations()
    from dex_retargeting import yourdfpy as urdf

    hands = ref_root / "assets" / "robots" / "hands"
    stems = sorted({Path(c["urdf_path"]).as_posix() for c in configs().values()})
    out = {"urdfs": np.array(stems)}
    rng = np.random.RandomState(20260923)
    for rel in stems:
        for dummy in (False, True):
            u = urdf.URDF.load(str(hands / rel), add_dummy_free_joints=dummy, build_scene_graph=False)
            u._base_link = u._determine_base_link()          # what the constructor does when it builds a scene graph (:611-612)
            u.tree_root = u.build_tree()
            tag = f"{Path(rel).stem}/{'dummy' if dummy else 'plain'}"
            act = list(u.actuated_joint_names)
            lim = np.array([[u.joint_map[n].limit.lower, u.joint_map[n].limit.upper] for n in act], dtype=np.float64)
            links = [l.name for l in u.robot.links]
            qs, poses = [], []
            for _ in range(N_Q):
                q = rng.uniform(lim[:, 0], lim[:, 1])
                u._cfg = q.copy()                            # mimic joints read their source from the stored configuration (:1017-1023)
                u.update_kinematics(q)
                qs.append(q)
                poses.append(np.array([u.get_link_global_transform(n) for n in links]))
            out[f"{tag}/actuated"] = np.array(act)
            out[f"{tag}/q"] = np.array(qs)
            out[f"{tag}/links"] = np.array(links)
            out[f"{tag}/poses"] = np.array(poses)
            print(f"{tag:40s} {len(act):2d} actuated joints, {len(links):2d} links")
    dst = ROOT / "tests" / "golden" / "reference_fk_vectors.npz"
    np.savez_compressed(dst, **out)
    print("wrote", dst, dst.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
