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
