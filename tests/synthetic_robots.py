"""Programmatically generated robot descriptions for edge-case tests (maximum sizes, deep chains)."""
import json

import numpy as np


def serial_chain(n_joints=32, seed=0, prismatic_every=0):
    """A single open chain of `n_joints` movable joints with varied axes / offsets and a tip link every 8 joints."""
    rng = np.random.RandomState(seed)
    links = ["base"]
    joints = []
    axes = [[1.0, 0, 0], [0, 1.0, 0], [0, 0, 1.0], [0, -1.0, 0], [0.6, 0.8, 0.0]]
    for i in range(n_joints):
        links.append(f"l{i}")
        jt = "prismatic" if prismatic_every and (i % prismatic_every == prismatic_every - 1) else "revolute"
        joints.append(dict(name=f"j{i:02d}", type=jt, parent="base" if i == 0 else f"l{i - 1}", child=f"l{i}",
                           xyz=[0.0, 0.0, 0.03] if i else [0.0, 0.0, 0.0], rpy=[float(rng.uniform(-0.3, 0.3)), 0.0, float(rng.uniform(-0.3, 0.3))],
                           axis=axes[i % len(axes)], limit=[-0.1, 0.1] if jt == "prismatic" else [-0.6, 0.6]))
        if i % 8 == 7:
            links.append(f"tip{i}")
            joints.append(dict(name=f"fix{i:02d}", type="fixed", parent=f"l{i}", child=f"tip{i}", xyz=[0.01, 0.0, 0.02],
                               rpy=[0.0, 0.0, 0.0], axis=[1.0, 0, 0]))
    return dict(name=f"chain{n_joints}", links=links, joints=joints)


def write_chain(tmp_path, n_joints=32, seed=0, prismatic_every=0):
    desc = serial_chain(n_joints, seed, prismatic_every)
    p = tmp_path / f"chain{n_joints}.json"
    p.write_text(json.dumps(desc))
    tips = [f"tip{i}" for i in range(7, n_joints, 8)]
    cfg = dict(type="position", urdf_path=str(p), target_link_names=tips, target_link_human_indices=[4 * (k + 1) for k in range(len(tips))],
               low_pass_alpha=1.0)
    return p, cfg


def tree_hand(trunk=2, widths=(4, 5, 4, 4, 5), seed=0):
    """A hand-shaped tree: a chain of `trunk` joints from the base, then one chain of widths[f] joints per finger, all
    attached to the last trunk link, each ending in a tip link.  Joint names sort so that pinocchio's depth-first order is
    trunk first, then finger 0, finger 1, ... (contiguous lanes)."""
    rng = np.random.RandomState(seed)
    axes = [[1.0, 0, 0], [0, 1.0, 0], [0, 0, 1.0], [0, -1.0, 0], [0.6, 0.8, 0.0]]
    links, joints = ["base"], []
    parent = "base"
    for i in range(trunk):
        links.append(f"trunk{i}")
        joints.append(dict(name=f"a_trunk{i}", type="revolute", parent=parent, child=f"trunk{i}", xyz=[0.0, 0.0, 0.02 if i else 0.0],
                           rpy=[float(rng.uniform(-0.2, 0.2)), 0.0, 0.0], axis=axes[i % 5], limit=[-0.5, 0.5]))
        parent = f"trunk{i}"
    palm = parent
    tips = []
    for f, w in enumerate(widths):
        ang = (f - (len(widths) - 1) / 2) * 0.35
        par = palm
        for i in range(w):
            links.append(f"f{f}_l{i}")
            xyz = [0.03 * float(np.sin(ang)), 0.01 * f, 0.05 * float(np.cos(ang))] if i == 0 else [0.0, 0.0, 0.028]
            joints.append(dict(name=f"f{f}_j{i}", type="revolute", parent=par, child=f"f{f}_l{i}", xyz=xyz,
                               rpy=[0.0, float(rng.uniform(-0.2, 0.2)), ang if i == 0 else 0.0], axis=axes[(f + i) % 5], limit=[-0.2, 1.2]))
            par = f"f{f}_l{i}"
        links.append(f"f{f}_tip")
        joints.append(dict(name=f"f{f}_tipfix", type="fixed", parent=par, child=f"f{f}_tip", xyz=[0.0, 0.0, 0.025], rpy=[0.0, 0.0, 0.0],
                           axis=[1.0, 0, 0]))
        tips.append(f"f{f}_tip")
    return dict(name=f"tree{trunk}_" + "".join(str(w) for w in widths), links=links, joints=joints), tips


def write_tree_hand(tmp_path, trunk, widths, seed=0, kind="position"):
    desc, tips = tree_hand(trunk, widths, seed)
    p = tmp_path / f"{desc['name']}.json"
    p.write_text(json.dumps(desc))
    hi = [min(20, 4 * (k + 1)) for k in range(len(tips))]
    if kind == "position":
        cfg = dict(type="position", urdf_path=str(p), target_link_names=tips, target_link_human_indices=hi, low_pass_alpha=1.0)
    else:
        cfg = dict(type="vector", urdf_path=str(p), target_origin_link_names=["base"] * len(tips), target_task_link_names=tips,
                   target_link_human_indices=[[0] * len(tips), hi], scaling_factor=1.0, low_pass_alpha=1.0)
    return p, cfg
