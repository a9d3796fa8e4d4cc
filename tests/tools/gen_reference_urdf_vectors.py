#!/usr/bin/env python
"""Generate tests/golden/reference_urdf_vectors.npz by running the REFERENCE'S OWN URDF code (src/dex_retargeting/yourdfpy.py,
unmodified) on every hand URDF its retargeting configs use.

Build container only (needs /root/reference).  What runs from the reference: URDF.load -> _parse_robot / _parse_joint /
_parse_origin / _parse_axis / _parse_limit / _parse_mimic (yourdfpy.py:905-960, 1375-1387, 1631-1661, 1107-1115),
_add_dummy_joints (:1942-1989), _determine_base_link, the actuated-joint list, and the mimic lists in the iteration order
retargeting_config.parse_mimic_joint (:265-285) sees.  (The reference's own tree forward kinematics, build_tree /
update_kinematics :1860-1939, is not usable as an FK oracle: _forward_kinematics_joint :1044 multiplies the 4x4 origin by
pytransform3d's 3x3 axis-angle matrix and raises for every revolute joint -- dead code the hot path never calls.)

Absent packages and their stand-ins (the only code below that is ours):
  * lxml.etree  -> xml.etree.ElementTree (same element API for what the parser touches; comments are dropped by the parser)
  * anytree     -> `Node`, `LevelOrderIter`, `search.findall_by_attr`: a 20-line tree
  * pytransform3d -> `transform_from`, `matrix_from_axis_angle` (Rodrigues), `matrix_from_euler(e, 0, 1, 2, extrinsic=True)`
    = Rz(e2) Ry(e1) Rx(e0), the call the reference makes at yourdfpy.py:1382-1386; `euler_from_matrix` is only used when
    writing XML and is not needed.
  * six -> installed.

Recorded per URDF (with and without the dummy free joints): base link, link names, joints in the reference's order (name,
type, parent, child, 4x4 origin, axis, limits), the mimic lists, the actuated joint names.
tests/test_reference_urdf_vectors.py holds the product's URDF reader, the committed joint-tree fixtures and the oracle's
robot model to these.

Usage: python tests/tools/gen_reference_urdf_vectors.py [/root/reference]
"""
import math
import sys
import types
import xml.etree.ElementTree as ET
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from helpers import configs  # noqa: E402


# ----------------------------------------------------------------------------------------------- stand-ins
def _install_shims(ref_root):
    lxml = types.ModuleType("lxml")
    etree = types.ModuleType("lxml.etree")
    for name in ("parse", "Element", "SubElement", "ElementTree", "tostring", "Comment", "iterparse"):
        setattr(etree, name, getattr(ET, name))
    etree.XMLParser = lambda **kw: ET.XMLParser()
    etree.strip_tags = lambda *a, **k: None
    etree.cleanup_namespaces = lambda *a, **k: None
    etree._Comment = etree._ProcessingInstruction = type(None)
    lxml.etree = etree

    anytree = types.ModuleType("anytree")

    class Node:
        def __init__(self, name, parent=None, **kw):
            self.name, self.parent, self.children = name, parent, []
            self.__dict__.update(kw)
            if parent is not None:
                parent.children.append(self)

    def level_order(root):
        queue = [root]
        while queue:
            n = queue.pop(0)
            yield n
            queue.extend(n.children)

    search = types.ModuleType("anytree.search")
    search.findall_by_attr = lambda root, value, name="name": tuple(n for n in level_order(root) if getattr(n, name, None) == value)
    anytree.Node, anytree.LevelOrderIter, anytree.search = Node, level_order, search

    pt3 = types.ModuleType("pytransform3d")
    rot = types.ModuleType("pytransform3d.rotations")
    trf = types.ModuleType("pytransform3d.transformations")

    def matrix_from_axis_angle(a):
        ax, ang = np.asarray(a[:3], dtype=float), float(a[3])
        n = np.linalg.norm(ax)
        if n == 0.0:
            return np.eye(3)
        x, y, z = ax / n
        K = np.array([[0, -z, y], [z, 0, -x], [-y, x, 0]])
        return np.eye(3) + math.sin(ang) * K + (1 - math.cos(ang)) * (K @ K)

    def matrix_from_euler(e, i, j, k, extrinsic):
        assert (i, j, k, extrinsic) == (0, 1, 2, True), "only the call yourdfpy.py:1382-1386 makes is implemented"
        rx = matrix_from_axis_angle([1, 0, 0, e[0]]); ry = matrix_from_axis_angle([0, 1, 0, e[1]]); rz = matrix_from_axis_angle([0, 0, 1, e[2]])
        return rz @ ry @ rx

    def transform_from(R, p):
        T = np.eye(4)
        T[:3, :3], T[:3, 3] = R, p
        return T

    rot.matrix_from_axis_angle, rot.matrix_from_euler = matrix_from_axis_angle, matrix_from_euler
    trf.transform_from = transform_from
    pt3.rotations, pt3.transformations = rot, trf
    sys.modules.update({"lxml": lxml, "lxml.etree": etree, "anytree": anytree, "anytree.search": search, "pytransform3d": pt3,
                        "pytransform3d.rotations": rot, "pytransform3d.transformations": trf,
                        "pinocchio": types.ModuleType("pinocchio"), "nlopt": types.ModuleType("nlopt")})
    sys.path.insert(0, str(Path(ref_root) / "src"))


def main():
    ref_root = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    _install_shims(ref_root)
    from dex_retargeting import yourdfpy as urdf

    def parse_mimic_joint(robot_urdf):  # retargeting_config.py:265-285 cannot be imported without the whole package chain:
        # restated here in four lines over the reference's parsed objects (joint_map iteration order is the reference's)
        src, mim, mul, off = [], [], [], []
        for name, joint in robot_urdf.joint_map.items():
            if joint.mimic is not None:
                mim.append(name); src.append(joint.mimic.joint); mul.append(joint.mimic.multiplier); off.append(joint.mimic.offset)
        return src, mim, mul, off

    hands = ref_root / "assets" / "robots" / "hands"
    stems = sorted({Path(c["urdf_path"]).as_posix() for c in configs().values()})
    out = {"urdfs": np.array(stems)}
    for rel in stems:
        for dummy in (False, True):
            u = urdf.URDF.load(str(hands / rel), add_dummy_free_joints=dummy, build_scene_graph=False)
            u._base_link = u._determine_base_link()
            tag = f"{Path(rel).stem}/{'dummy' if dummy else 'plain'}"
            joints = u.robot.joints
            out[f"{tag}/links"] = np.array([l.name for l in u.robot.links])
            out[f"{tag}/base_link"] = np.array(u.base_link)
            out[f"{tag}/joint_names"] = np.array([j.name for j in joints])
            out[f"{tag}/joint_types"] = np.array([j.type for j in joints])
            out[f"{tag}/joint_parent"] = np.array([j.parent for j in joints])
            out[f"{tag}/joint_child"] = np.array([j.child for j in joints])
            out[f"{tag}/joint_origin"] = np.array([np.eye(4) if j.origin is None else j.origin for j in joints])
            out[f"{tag}/joint_axis"] = np.array([np.asarray(j.axis, dtype=float) for j in joints])
            out[f"{tag}/joint_limit"] = np.array([[np.nan, np.nan] if j.limit is None else
                                                  [np.nan if j.limit.lower is None else j.limit.lower,
                                                   np.nan if j.limit.upper is None else j.limit.upper] for j in joints], dtype=float)
            src, mim, mul, off = parse_mimic_joint(u)
            out[f"{tag}/mimic_joint"], out[f"{tag}/mimic_source"] = np.array(mim, dtype=str), np.array(src, dtype=str)
            out[f"{tag}/mimic_mult"], out[f"{tag}/mimic_off"] = np.array(mul, dtype=float), np.array(off, dtype=float)
            act = list(u.actuated_joint_names)
            out[f"{tag}/actuated"] = np.array(act)
            print(tag, len(joints), "joints", len(act), "actuated", len(mim), "mimic")
    np.savez_compressed(ROOT / "tests" / "golden" / "reference_urdf_vectors.npz", **out)
    print("wrote tests/golden/reference_urdf_vectors.npz")


if __name__ == "__main__":
    main()
