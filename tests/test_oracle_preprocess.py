"""Oracle pre-processing restatement: the recorded trajectory IS the output of the reference's detector
post-processing (example/profiling/human_joint_right.pkl: wrist at the origin, wrist-frame aligned), so
re-applying the frame estimation to it must give (numerically) the MANO-convention identity mapping."""
import numpy as np

from helpers import keypoint_trajectory
from oracle.preprocess import OPERATOR2MANO, estimate_frame_from_hand_points, preprocess


def test_frame_is_orthonormal_and_right_handed_up_to_convention():
    kp = keypoint_trajectory().astype(np.float64)
    for f in range(0, kp.shape[0], 37):
        rot = estimate_frame_from_hand_points(kp[f] - kp[f][0:1])
        np.testing.assert_allclose(rot.T @ rot, np.eye(3), atol=1e-12)


def test_recorded_trajectory_is_a_fixed_point():
    kp = keypoint_trajectory().astype(np.float64)
    for f in range(0, kp.shape[0], 23):
        out, rot = preprocess(kp[f], "right")
        # already processed data: frame @ operator2mano is the identity, output == input
        np.testing.assert_allclose(rot @ OPERATOR2MANO["right"], np.eye(3), atol=2e-6)
        np.testing.assert_allclose(out, kp[f], atol=2e-6)


def test_left_right_mirror():
    kp = keypoint_trajectory()[100].astype(np.float64)
    r, _ = preprocess(kp, "right")
    l, _ = preprocess(kp, "left")
    np.testing.assert_allclose(l[:, 0], -r[:, 0], atol=1e-12)
    np.testing.assert_allclose(l[:, 1], -r[:, 1], atol=1e-12)
    np.testing.assert_allclose(l[:, 2], r[:, 2], atol=1e-12)
