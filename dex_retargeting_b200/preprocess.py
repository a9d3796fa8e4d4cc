"""Keypoint pre-processing on device: detector landmarks -> wrist-centred, wrist-aligned MANO-convention keypoints.

The step right before the hot path in the reference's teleoperation pipeline
(example/vector_retargeting/single_hand_detector.py:100-103 and :130-158; OPERATOR2MANO matrices
src/dex_retargeting/constants.py:7-21).  One launch for the whole batch (`dexr_preprocess_keypoints`);
its output is exactly what `Optimizer.retarget_batch(keypoints=...)` / `retarget_sequences` consume.
"""
from __future__ import annotations

import ctypes as C

from . import _native as N
from .constants import HandType


def preprocess_keypoints(raw, hand_type: HandType = HandType.right, out=None, wrist_rot_out=None, stream=None):
    """raw: float32 CUDA tensor [..., 21, 3] of detector world landmarks (any leading shape).
    Returns `out` (same shape): (raw - wrist) @ wrist_frame @ operator2mano.  `wrist_rot_out` [..., 3, 3], if
    given, receives the estimated wrist frame (what the reference returns as `mediapipe_wrist_rot`)."""
    import torch

    if raw.dim() < 2 or tuple(raw.shape[-2:]) != (N.NUM_KEYPOINTS, 3):
        raise ValueError(f"raw keypoints must have shape [...,21,3], got {tuple(raw.shape)}")
    if not raw.is_cuda or raw.dtype != torch.float32 or not raw.is_contiguous():
        raise ValueError("raw keypoints must be a contiguous float32 CUDA tensor")
    B = raw.numel() // (N.NUM_KEYPOINTS * 3)
    if out is None:
        out = torch.empty_like(raw)
    elif out.shape != raw.shape or out.dtype != raw.dtype or out.device != raw.device or not out.is_contiguous():
        raise ValueError("out must match raw (shape, dtype, device, contiguous)")
    rot_ptr = None
    if wrist_rot_out is not None:
        if (tuple(wrist_rot_out.shape) != tuple(raw.shape[:-2]) + (3, 3) or wrist_rot_out.dtype != torch.float32
                or wrist_rot_out.device != raw.device or not wrist_rot_out.is_contiguous()):
            raise ValueError("wrist_rot_out must be a contiguous float32 tensor [...,3,3] on the same device")
        rot_ptr = wrist_rot_out.data_ptr()
    s = stream if stream is not None else torch.cuda.current_stream(raw.device)
    lib = N.load()
    N.check(lib.dexr_preprocess_keypoints(raw.data_ptr(), out.data_ptr(), rot_ptr, 0 if hand_type is HandType.right else 1,
                                          B, raw.device.index, C.c_void_p(s.cuda_stream)), "dexr_preprocess_keypoints")
    return out
