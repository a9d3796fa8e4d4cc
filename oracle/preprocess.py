"""Oracle for the keypoint pre-processing step -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates example/vector_retargeting/single_hand_detector.py:100-103 (centre at the wrist, rotate into the
estimated wrist frame, then into the MANO convention) and :130-158 (estimate_frame_from_hand_points: plane
fit of landmarks {0,5,9} by SVD, Gram-Schmidt, sign fix), with the OPERATOR2MANO matrices of
src/dex_retargeting/constants.py:7-21.  float64, numpy's SVD exactly as the reference calls it.
"""
import numpy as np

OPERATOR2MANO = {
    "right": np.array([[0, 0, -1], [-1, 0, 0], [0, 1, 0]], dtype=np.float64),
    "left": np.array([[0, 0, -1], [1, 0, 0], [0, -1, 0]], dtype=np.float64),
}


def estimate_frame_from_hand_points(kp):
    assert kp.shape == (21, 3)
    points = kp[[0, 5, 9], :]
    x_vector = points[0] - points[2]
    points = points - np.mean(points, axis=0, keepdims=True)
    _, _, v = np.linalg.svd(points)
    normal = v[2, :]
    x = x_vector - np.sum(x_vector * normal) * normal
    x = x / np.linalg.norm(x)
    z = np.cross(x, normal)
    if np.sum(z * (points[1] - points[2])) < 0:
        normal = -normal
        z = -z
    return np.stack([x, normal, z], axis=1)


def preprocess(raw, hand="right"):
    """raw (21,3) detector landmarks -> (joint_pos (21,3), wrist_rot (3,3))."""
    kp = np.asarray(raw, dtype=np.float64)
    kp = kp - kp[0:1, :]
    rot = estimate_frame_from_hand_points(kp)
    return kp @ rot @ OPERATOR2MANO[hand], rot
