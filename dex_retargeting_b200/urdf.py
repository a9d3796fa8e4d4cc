"""URDF subset reader and kinematic model (host side, init time only).

The reference parses URDFs with its vendored yourdfpy fork (lxml + pytransform3d + anytree) and
then hands a re-written XML file to pinocchio.  Neither is needed for the hot path: the solver only
needs the joint tree.  This module reads the same URDF subset with the standard library and exposes
a `KinematicModel` whose degree-of-freedom order reproduces pinocchio's.

Behaviour followed (reference file:line, relative to /root/reference):
  * origin  : src/dex_retargeting/yourdfpy.py:1375-1387   R = Rz(yaw) Ry(pitch) Rx(roll), extrinsic xyz
  * axis    : src/dex_retargeting/yourdfpy.py:1631-1643   default "1 0 0", unparsable token -> 0
  * limit   : src/dex_retargeting/yourdfpy.py:1652-1661
  * mimic   : src/dex_retargeting/yourdfpy.py:1107-1115   defaults multiplier 1, offset 0
  * dummies : src/dex_retargeting/yourdfpy.py:1942-1989   3 prismatic (x,y,z, +-5 m) then 3 revolute
              (x,y,z, +-2 pi) prepended to the root, identity origins
  * DoF order [third party: pinocchio urdf parser over urdfdom]: depth first from the root link,
    children of a link visited in lexicographic order of the *joint* name (urdfdom keeps joints in a
    std::map), fixed joints merged into their parent, one DoF per revolute / prismatic joint.
    robot_wrapper.py:19-23 rejects nq != nv models, so "continuous"/"floating"/"planar" raise here.

The model can be serialised to a small JSON document (`to_dict` / `from_dict`): that is the robot
description format used by the test fixtures under tests/golden/robots (the URDF assets are a git
submodule of the reference and do not travel to the GPU box).
"""
from __future__ import annotations

import json
import math
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, List, Optional, Tuple

import numpy as np

DUMMY_JOINT_NAMES = [f"dummy_{a}_translation_joint" for a in "xyz"] + [
    f"dummy_{a}_rotation_joint" for a in "xyz"
]
_DUMMY_LINK_NAMES = [f"dummy_{a}_translation_link" for a in "xyz"] + [
    f"dummy_{a}_rotation_link" for a in "xyz"
]

MOVABLE_TYPES = ("revolute", "prismatic")


def rpy_to_matrix(rpy) -> np.ndarray:
    """URDF fixed-axis roll/pitch/yaw -> rotation matrix, R = Rz(y) Ry(p) Rx(r)."""
    r, p, y = (float(v) for v in rpy)
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    return np.array(
        [
            [cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
            [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
            [-sp, cp * sr, cp * cr],
        ],
        dtype=np.float64,
    )


def _floats(text: Optional[str], default: str, lenient: bool = False) -> List[float]:
    out = []
    for tok in (text if text is not None else default).split():
        try:
            out.append(float(tok))
        except ValueError:
            if not lenient:
                raise
            out.append(0.0)
    return out


@dataclass
class JointSpec:
    name: str
    type: str
    parent: str
    child: str
    xyz: List[float] = field(default_factory=lambda: [0.0, 0.0, 0.0])
    rpy: List[float] = field(default_factory=lambda: [0.0, 0.0, 0.0])
    axis: List[float] = field(default_factory=lambda: [1.0, 0.0, 0.0])
    lower: Optional[float] = None
    upper: Optional[float] = None
    mimic: Optional[Tuple[str, float, float]] = None  # (source joint, multiplier, offset)

    def to_dict(self) -> dict:
        d = dict(name=self.name, type=self.type, parent=self.parent, child=self.child,
                 xyz=list(self.xyz), rpy=list(self.rpy), axis=list(self.axis))
        if self.lower is not None or self.upper is not None:
            d["limit"] = [self.lower, self.upper]
        if self.mimic is not None:
            d["mimic"] = list(self.mimic)
        return d

    @staticmethod
    def from_dict(d: dict) -> "JointSpec":
        lim = d.get("limit", [None, None])
        mim = d.get("mimic")
        return JointSpec(d["name"], d["type"], d["parent"], d["child"], list(d["xyz"]), list(d["rpy"]),
                         list(d["axis"]), lim[0], lim[1], tuple(mim) if mim is not None else None)


class KinematicModel:
    """Joint tree with fixed joints folded, in pinocchio DoF order.

    Attributes (n = number of DoFs, all arrays float64 unless noted):
      dof_joint_names   list[n]   movable joint names in DoF order
      joint_type        (n,) int  0 revolute, 1 prismatic
      joint_parent      (n,) int  DoF index of the closest movable ancestor, -1 for the world
      joint_R, joint_p  (n,3,3),(n,3)  placement of the joint frame in the parent joint frame
      joint_axis        (n,3)     unit axis in the joint frame
      joint_limits      (n,2)
      link_names        list      every URDF link
      link_parent       (L,) int  DoF index of the joint the link is rigidly attached to, -1 world
      link_R, link_p    (L,3,3),(L,3)  placement of the link frame in that joint frame
    """

    def __init__(self, name: str, link_names: List[str], joints: List[JointSpec]):
        self.name = name
        self.urdf_link_names = list(link_names)
        self.joints = list(joints)
        self.joint_map: Dict[str, JointSpec] = {j.name: j for j in self.joints}
        if len(self.joint_map) != len(self.joints):
            raise ValueError("Duplicate joint names in robot description")
        self._compile()

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_urdf(cls, path, add_dummy_free_joints: bool = False) -> "KinematicModel":
        root = ET.parse(str(path)).getroot()
        if root.tag != "robot":
            raise ValueError(f"{path}: root element is <{root.tag}>, expected <robot>")
        links = [e.attrib["name"] for e in root.findall("link")]
        joints = []
        for e in root.findall("joint"):
            jtype = e.get("type")
            origin = e.find("origin")
            axis = e.find("axis")
            limit = e.find("limit")
            mimic = e.find("mimic")
            spec = JointSpec(
                name=e.attrib["name"],
                type=jtype,
                parent=e.find("parent").get("link"),
                child=e.find("child").get("link"),
                xyz=_floats(origin.get("xyz") if origin is not None else None, "0 0 0"),
                rpy=_floats(origin.get("rpy") if origin is not None else None, "0 0 0"),
                axis=_floats(axis.get("xyz") if axis is not None else None, "1 0 0", lenient=True),
            )
            if limit is not None:
                lo, hi = limit.get("lower"), limit.get("upper")
                spec.lower = float(lo) if lo is not None else None
                spec.upper = float(hi) if hi is not None else None
            if mimic is not None:
                spec.mimic = (mimic.get("joint"), float(mimic.get("multiplier", 1.0)),
                              float(mimic.get("offset", 0.0)))
            joints.append(spec)
        model_name = root.get("name", Path(str(path)).stem)
        if add_dummy_free_joints:
            links, joints = _prepend_dummy_joints(links, joints)
        return cls(model_name, links, joints)

    @classmethod
    def from_dict(cls, d: dict, add_dummy_free_joints: bool = False) -> "KinematicModel":
        links = list(d["links"])
        joints = [JointSpec.from_dict(j) for j in d["joints"]]
        if add_dummy_free_joints:
            links, joints = _prepend_dummy_joints(links, joints)
        return cls(d.get("name", "robot"), links, joints)

    @classmethod
    def load(cls, path, add_dummy_free_joints: bool = False) -> "KinematicModel":
        """Load a `.urdf` (XML) or a `.json` robot description."""
        p = Path(str(path))
        if p.suffix.lower() == ".json":
            with p.open("r") as f:
                return cls.from_dict(json.load(f), add_dummy_free_joints)
        return cls.from_urdf(p, add_dummy_free_joints)

    def write_urdf(self, path) -> None:
        """Kinematics-only URDF of this joint tree (links without geometry or inertia; joints with origin, axis, limit,
        mimic), written with 17 significant digits so that reading it back reproduces every float exactly.  This is what
        ships under `dex_retargeting_b200/assets/robots/hands/` -- the retargeting path needs nothing else from a URDF."""
        num = lambda v: " ".join(repr(float(x)) for x in v)  # noqa: E731
        root = ET.Element("robot", name=self.name)
        for link in self.urdf_link_names:
            ET.SubElement(root, "link", name=link)
        for j in self.joints:
            e = ET.SubElement(root, "joint", name=j.name, type=j.type)
            ET.SubElement(e, "parent", link=j.parent)
            ET.SubElement(e, "child", link=j.child)
            ET.SubElement(e, "origin", xyz=num(j.xyz), rpy=num(j.rpy))
            if j.type != "fixed":
                ET.SubElement(e, "axis", xyz=num(j.axis))
            if j.lower is not None or j.upper is not None:
                lim = {}
                if j.lower is not None:
                    lim["lower"] = repr(float(j.lower))
                if j.upper is not None:
                    lim["upper"] = repr(float(j.upper))
                ET.SubElement(e, "limit", effort="1", velocity="1", **lim)
            if j.mimic is not None:
                ET.SubElement(e, "mimic", joint=j.mimic[0], multiplier=repr(float(j.mimic[1])), offset=repr(float(j.mimic[2])))
        ET.indent(root, space=" ")
        Path(str(path)).parent.mkdir(parents=True, exist_ok=True)
        ET.ElementTree(root).write(str(path), encoding="unicode", xml_declaration=True)

    def to_dict(self) -> dict:
        """URDF-level description (before any dummy joints were added by this object's caller)."""
        return dict(name=self.name, links=list(self.urdf_link_names), joints=[j.to_dict() for j in self.joints])

    # ------------------------------------------------------------------ compilation
    def _compile(self):
        children: Dict[str, List[JointSpec]] = {n: [] for n in self.urdf_link_names}
        is_child = set()
        for j in self.joints:
            if j.parent not in children or j.child not in children:
                raise ValueError(f"Joint {j.name} refers to an unknown link")
            children[j.parent].append(j)
            is_child.add(j.child)
        roots = [n for n in self.urdf_link_names if n not in is_child]
        if len(roots) != 1:
            raise ValueError(f"Robot description must have exactly one root link, found {roots}")
        self.root_link = roots[0]
        for lst in children.values():
            lst.sort(key=lambda s: s.name)

        names, jtype, jparent, jR, jp, jaxis, jlim = [], [], [], [], [], [], []
        link_names, link_parent, link_R, link_p = [], [], [], []
        link_parent_joint_name: Dict[str, Optional[str]] = {}

        def visit(link: str, sup: int, R: np.ndarray, p: np.ndarray, via: Optional[str]):
            # `sup` = DoF index of the movable joint this link rides on (-1: world);
            # (R, p) = placement of the link frame in that joint's frame.
            link_names.append(link)
            link_parent.append(sup)
            link_R.append(R)
            link_p.append(p)
            link_parent_joint_name[link] = via
            for j in children[link]:
                Rc = R @ rpy_to_matrix(j.rpy)
                pc = R @ np.asarray(j.xyz, dtype=np.float64) + p
                if j.type == "fixed":
                    visit(j.child, sup, Rc, pc, j.name)
                elif j.type in MOVABLE_TYPES:
                    axis = np.asarray(j.axis, dtype=np.float64)
                    nrm = np.linalg.norm(axis)
                    if nrm == 0:
                        raise ValueError(f"Joint {j.name} has a zero axis")
                    idx = len(names)
                    names.append(j.name)
                    jtype.append(0 if j.type == "revolute" else 1)
                    jparent.append(sup)
                    jR.append(Rc)
                    jp.append(pc)
                    jaxis.append(axis / nrm)
                    jlim.append((j.lower if j.lower is not None else 0.0,
                                 j.upper if j.upper is not None else 0.0))
                    visit(j.child, idx, np.eye(3), np.zeros(3), j.name)
                else:
                    raise NotImplementedError(
                        f"Can not handle robot with special joint: {j.name} is of type {j.type!r}"
                    )

        visit(self.root_link, -1, np.eye(3), np.zeros(3), None)

        self.dof_joint_names: List[str] = names
        self.dof = len(names)
        self.joint_type = np.asarray(jtype, dtype=np.int64).reshape(-1)
        self.joint_parent = np.asarray(jparent, dtype=np.int64).reshape(-1)
        self.joint_R = np.asarray(jR, dtype=np.float64).reshape(-1, 3, 3)
        self.joint_p = np.asarray(jp, dtype=np.float64).reshape(-1, 3)
        self.joint_axis = np.asarray(jaxis, dtype=np.float64).reshape(-1, 3)
        self.joint_limits = np.asarray(jlim, dtype=np.float64).reshape(-1, 2)
        self.link_names: List[str] = link_names
        self.link_parent = np.asarray(link_parent, dtype=np.int64)
        self.link_R = np.asarray(link_R, dtype=np.float64).reshape(-1, 3, 3)
        self.link_p = np.asarray(link_p, dtype=np.float64).reshape(-1, 3)
        self.link_parent_joint_name = link_parent_joint_name
        if len(link_names) != len(self.urdf_link_names):
            raise ValueError("Robot description is not a connected tree")
        # depth (number of movable ancestors, inclusive) per joint
        depth = np.zeros(self.dof, dtype=np.int64)
        for i in range(self.dof):
            depth[i] = 1 + (depth[self.joint_parent[i]] if self.joint_parent[i] >= 0 else 0)
        self.joint_depth = depth

    # ------------------------------------------------------------------ queries
    def link_index(self, name: str) -> int:
        if name not in self.link_names:
            raise ValueError(f"{name} is not a link name. Valid link names: \n{self.link_names}")
        return self.link_names.index(name)

    def mimic_joints(self):
        """(source names, mimic names, multipliers, offsets) in URDF joint order.

        Follows retargeting_config.py:265-285 (iteration over `joint_map`, i.e. file order)."""
        src, mim, mul, off = [], [], [], []
        for j in self.joints:
            if j.mimic is not None:
                mim.append(j.name)
                src.append(j.mimic[0])
                mul.append(float(j.mimic[1]))
                off.append(float(j.mimic[2]))
        return src, mim, mul, off

    # ------------------------------------------------------------------ float64 host kinematics
    def forward_kinematics(self, q: np.ndarray):
        """World placement (R, p) of every movable joint frame after its own motion."""
        q = np.asarray(q, dtype=np.float64)
        if q.shape != (self.dof,):
            raise ValueError(f"qpos must have shape ({self.dof},), got {q.shape}")
        Rw = np.empty((self.dof, 3, 3))
        pw = np.empty((self.dof, 3))
        for i in range(self.dof):
            par = self.joint_parent[i]
            if par >= 0:
                R0 = Rw[par] @ self.joint_R[i]
                p0 = Rw[par] @ self.joint_p[i] + pw[par]
            else:
                R0, p0 = self.joint_R[i], self.joint_p[i]
            a = self.joint_axis[i]
            if self.joint_type[i] == 0:
                Rw[i] = R0 @ _axis_angle(a, q[i])
                pw[i] = p0
            else:
                Rw[i] = R0
                pw[i] = p0 + R0 @ (a * q[i])
        return Rw, pw

    def link_pose(self, Rw, pw, link: int):
        par = self.link_parent[link]
        if par < 0:
            return self.link_R[link].copy(), self.link_p[link].copy()
        return Rw[par] @ self.link_R[link], Rw[par] @ self.link_p[link] + pw[par]

    def is_ancestor_table(self) -> np.ndarray:
        """anc[i, j] = True if joint j is joint i or one of its movable ancestors."""
        anc = np.zeros((self.dof, self.dof), dtype=bool)
        for i in range(self.dof):
            k = i
            while k >= 0:
                anc[i, k] = True
                k = self.joint_parent[k]
        return anc


def _axis_angle(a: np.ndarray, t: float) -> np.ndarray:
    c, s = math.cos(t), math.sin(t)
    x, y, z = a
    C = 1.0 - c
    return np.array(
        [
            [c + x * x * C, x * y * C - z * s, x * z * C + y * s],
            [y * x * C + z * s, c + y * y * C, y * z * C - x * s],
            [z * x * C - y * s, z * y * C + x * s, c + z * z * C],
        ]
    )


def _prepend_dummy_joints(links: List[str], joints: List[JointSpec]):
    is_child = {j.child for j in joints}
    roots = [n for n in links if n not in is_child]
    if len(roots) != 1:
        raise ValueError(f"Robot description must have exactly one root link, found {roots}")
    new_joints = []
    for i in range(6):
        axis = [0.0, 0.0, 0.0]
        axis[i % 3] = 1.0
        lo, hi = (-5.0, 5.0) if i < 3 else (-2 * math.pi, 2 * math.pi)
        new_joints.append(
            JointSpec(
                name=DUMMY_JOINT_NAMES[i],
                type="prismatic" if i < 3 else "revolute",
                parent=_DUMMY_LINK_NAMES[i],
                child=_DUMMY_LINK_NAMES[i + 1] if i < 5 else roots[0],
                axis=axis,
                lower=lo,
                upper=hi,
            )
        )
    return _DUMMY_LINK_NAMES + list(links), new_joints + list(joints)
