#!/usr/bin/env python
"""Generate tests/golden/reference_warm_start.npz by running the REFERENCE'S OWN `SeqRetargeting.warm_start`
(src/dex_retargeting/seq_retarget.py:45-110, unmodified) on seeded wrist poses, for every offline (free-flying base) config.

Stand-ins for the absent packages (as in gen_reference_vectors.py: pinocchio -> DuckRobot over the oracle's FK, which is held
to the reference's tree FK by tests/test_reference_fk_vectors.py; nlopt unused here).  pytransform3d's two functions the
method calls are served by a THIRD-PARTY implementation of the same published conventions, not by our formulas:
  rotations.matrix_from_quaternion(q)                 q = (w, x, y, z)  -> scipy Rotation.from_quat([x, y, z, w]).as_matrix()
  rotations.euler_from_matrix(R, 0, 1, 2, extrinsic=False)  intrinsic x-y'-z'' angles with R = Rx(a) Ry(b) Rz(c)
                                                      -> scipy Rotation.from_matrix(R).as_euler("XYZ")  (upper case = intrinsic)
Recorded per config: wrist positions / quaternions, hand type, convention flag, and `last_qpos` after the call.
tests/test_warm_start_reference.py (CPU) and tests/test_gpu_next_rows.py hold `SeqRetargeting.warm_start` and the batched
`warm_start_batch` of the product to them.

Usage: python tests/tools/gen_reference_warm_start.py [/root/reference]
"""
import sys
from pathlib import Path

import numpy as np
from scipy.spatial.transform import Rotation

ROOT = Path(__file__).resolve().parent.parent.parent
for p in (ROOT, ROOT / "tests", ROOT / "tests" / "tools"):
    sys.path.insert(0, str(p))
import gen_reference_vectors as G  # noqa: E402
from helpers import configs  # noqa: E402


def main():
    ref_root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    G.install_shims(ref_root)
    rot = sys.modules["pytransform3d.rotations"]
    rot.matrix_from_quaternion = lambda q: Rotation.from_quat([q[1], q[2], q[3], q[0]]).as_matrix()

    def euler_from_matrix(R, i, j, k, extrinsic):
        assert (i, j, k, extrinsic) == (0, 1, 2, False)
        return Rotation.from_matrix(np.asarray(R)).as_euler("XYZ")

    rot.euler_from_matrix = euler_from_matrix

    # what warm_start needs from RobotWrapper beyond the surface gen_reference_vectors.DuckRobot already has
    G.DuckRobot.q0 = property(lambda s: np.zeros(s.r.dof))                                       # pin.neutral for revolute / prismatic
    G.DuckRobot.get_link_pose_inv = lambda s, link_id: np.linalg.inv(s.r.get_link_pose(link_id))  # robot_wrapper.py:89-91

    def parent_child(s, joint_name):  # robot_wrapper.py:69-78: frames either side of the joint
        j = next(j for j in s.r.urdf_joints if j["name"] == joint_name)
        return s.r.get_link_index(j["parent"]), s.r.get_link_index(j["child"])

    G.DuckRobot.get_joint_parent_child_frames = parent_child
    from dex_retargeting.constants import HandType

    out = {}
    rng = np.random.RandomState(77)
    keys = sorted(k for k, c in configs().items() if c.get("add_dummy_free_joint"))
    for key in keys:
        pos = rng.randn(6, 3) * 0.4
        quat = rng.randn(6, 4)
        quat /= np.linalg.norm(quat, axis=1, keepdims=True)
        for mano in (False, True):
            for hand in ("right", "left"):
                res = []
                for s in range(6):
                    seq = G.build_reference(key)
                    seq.warm_start(pos[s], quat[s], HandType[hand], is_mano_convention=mano)
                    res.append(np.asarray(seq.last_qpos, dtype=np.float64).copy())
                out[f"{key}/{hand}/{int(mano)}"] = np.array(res)
        out[f"{key}/pos"], out[f"{key}/quat"] = pos, quat
        print(key, "ok")
    out["keys"] = np.array(keys)
    dst = ROOT / "tests" / "golden" / "reference_warm_start.npz"
    np.savez_compressed(dst, **out)
    print("wrote", dst, dst.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
