"""Output adapters: the formats on the far side of the hot path.

  * joint-order remapping by name (reference README.md:84-106, example/vector_retargeting/
    show_realtime_retargeting.py:113-118): retargeting returns qpos in pinocchio joint order; simulators
    want their own order.  `joint_order_map` builds the gather index once, `remap_qpos` applies it to a
    whole batch (numpy or torch, any device).
  * the pickle trajectory format written by example/vector_retargeting/detect_from_video.py:60-69:
    {"data": [qpos, ...], "meta_data": {"config_path", "dof", "joint_names"}}.
"""
from __future__ import annotations

import pickle
from pathlib import Path
from typing import List, Sequence

import numpy as np


def joint_order_map(source_joint_names: Sequence[str], target_joint_names: Sequence[str]) -> np.ndarray:
    """index such that qpos_target = qpos_source[..., index] (every target joint must exist in the source)."""
    src = list(source_joint_names)
    missing = [n for n in target_joint_names if n not in src]
    if missing:
        raise ValueError(f"joints {missing} are not produced by the retargeting (available: {src})")
    return np.array([src.index(n) for n in target_joint_names], dtype=np.int64)


def remap_qpos(qpos, index):
    """Gather the last axis: works on numpy arrays and torch tensors of any leading shape."""
    if isinstance(qpos, np.ndarray):
        return qpos[..., index]
    import torch

    return qpos.index_select(-1, torch.as_tensor(index, device=qpos.device))


def save_trajectory(path, qpos_sequence, joint_names: List[str], config_path: str = ""):
    """Write the reference's pickle layout.  qpos_sequence: [T, dof] array-like (one stream)."""
    arr = np.asarray(qpos_sequence.detach().cpu() if hasattr(qpos_sequence, "detach") else qpos_sequence)
    if arr.ndim != 2 or arr.shape[1] != len(joint_names):
        raise ValueError(f"expected [T,{len(joint_names)}] joint positions, got {arr.shape}")
    meta = dict(config_path=str(config_path), dof=len(joint_names), joint_names=list(joint_names))
    out = Path(path)
    out.parent.mkdir(parents=True, exist_ok=True)
    with out.open("wb") as f:
        pickle.dump(dict(data=[row.copy() for row in arr], meta_data=meta), f)
    return out


def load_trajectory(path):
    with open(path, "rb") as f:
        d = pickle.load(f)
    return np.asarray(d["data"]), d["meta_data"]
