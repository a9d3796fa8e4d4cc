"""Oracle robot model: float64 FK / Jacobians by walking the URDF-level tree with 4x4 matrices.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Deliberately shares no code with
dex_retargeting_b200: it reads the JSON robot description (tests/golden/robots/*.json) or a URDF
on its own and keeps every fixed joint as a separate 4x4 product, whereas the product folds fixed
joints into a flat table.

Restates (relative to /root/reference):
  robot_wrapper.py:28-52   joint_names / dof_joint_names / dof / link_names / joint_limits
  robot_wrapper.py:82-87   compute_forward_kinematics + get_link_pose  (pin.forwardKinematics,
                           pin.updateFramePlacement -> 4x4 homogeneous)
  robot_wrapper.py:93-95   compute_single_link_local_jacobian (pin.computeFrameJacobian, LOCAL);
                           callers rotate rows 0-2 by the link rotation (optimizer.py:172-177), which
                           is the world-aligned linear Jacobian of the frame origin -- returned here
                           directly by `link_jacobians`.
  kinematics_adaptor.py:102-113  mimic forward_qpos / backward_jacobian
  yourdfpy.py:1375-1387, 1631-1643, 1942-1989  origin / axis / dummy joints
DoF order is pinocchio's [third party]: depth first, children by joint name, fixed joints merged.
"""
import ctypes as C
import json
import math
import os
import subprocess
import xml.etree.ElementTree as ET
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIBORC = None


def build_c_kinematics(force=False):
    """gcc oracle/c/orc.c -> oracle/_build/liborc.so (the C restatement of this module's kinematics)."""
    src, out = _HERE / "c" / "orc.c", _HERE / "_build" / "liborc.so"
    if force or not out.exists() or out.stat().st_mtime < src.stat().st_mtime:
        out.parent.mkdir(exist_ok=True)
        subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-o", str(out), str(src), "-lm"], check=True)
    return out


def _load_c():
    """The C library if it can be had (built on demand), else None -> pure Python kinematics."""
    global _LIBORC
    if _LIBORC is None:
        if os.environ.get("ORACLE_PURE_PYTHON") == "1":
            _LIBORC = False
        else:
            try:
                _LIBORC = C.CDLL(str(build_c_kinematics()))
            except Exception:
                _LIBORC = False
    return _LIBORC or None


class _OrcModel(C.Structure):
    _fields_ = [("n_links", C.c_int), ("n_joints", C.c_int), ("dof", C.c_int), ("parent_link", C.c_void_p),
                ("child_link", C.c_void_p), ("type", C.c_void_p), ("dof_index", C.c_void_p), ("origin", C.c_void_p),
                ("axis", C.c_void_p), ("chain_start", C.c_void_p), ("chain_joint", C.c_void_p)]

DUMMY_JOINTS = ["dummy_%s_translation_joint" % a for a in "xyz"] + ["dummy_%s_rotation_joint" % a for a in "xyz"]
DUMMY_LINKS = ["dummy_%s_translation_link" % a for a in "xyz"] + ["dummy_%s_rotation_link" % a for a in "xyz"]


def _origin(xyz, rpy):
    r, p, y = rpy
    Rx = np.array([[1, 0, 0], [0, math.cos(r), -math.sin(r)], [0, math.sin(r), math.cos(r)]])
    Ry = np.array([[math.cos(p), 0, math.sin(p)], [0, 1, 0], [-math.sin(p), 0, math.cos(p)]])
    Rz = np.array([[math.cos(y), -math.sin(y), 0], [math.sin(y), math.cos(y), 0], [0, 0, 1]])
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry @ Rx
    T[:3, 3] = xyz
    return T


def _rodrigues(axis, angle):
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + math.sin(angle) * K + (1 - math.cos(angle)) * (K @ K)


def _read_urdf(path):
    root = ET.parse(str(path)).getroot()
    links = [e.get("name") for e in root.findall("link")]
    joints = []
    for e in root.findall("joint"):
        o, a, l, m = e.find("origin"), e.find("axis"), e.find("limit"), e.find("mimic")
        j = dict(name=e.get("name"), type=e.get("type"), parent=e.find("parent").get("link"),
                 child=e.find("child").get("link"),
                 xyz=[float(v) for v in (o.get("xyz", "0 0 0") if o is not None else "0 0 0").split()],
                 rpy=[float(v) for v in (o.get("rpy", "0 0 0") if o is not None else "0 0 0").split()],
                 axis=[float(v) for v in (a.get("xyz", "1 0 0") if a is not None else "1 0 0").split()])
        if l is not None:
            j["limit"] = [float(l.get("lower", 0.0)), float(l.get("upper", 0.0))]
        if m is not None:
            j["mimic"] = [m.get("joint"), float(m.get("multiplier", 1.0)), float(m.get("offset", 0.0))]
        joints.append(j)
    return dict(name=root.get("name", "robot"), links=links, joints=joints)


class OracleRobot:
    def __init__(self, desc, add_dummy_free_joint=False, use_c=None):
        if isinstance(desc, (str, bytes)) or hasattr(desc, "__fspath__"):
            p = str(desc)
            if p.endswith(".json"):
                with open(p) as f:
                    desc = json.load(f)
            else:
                desc = _read_urdf(p)
        links = list(desc["links"])
        joints = [dict(j) for j in desc["joints"]]
        if add_dummy_free_joint:
            child = {j["child"] for j in joints}
            (root,) = [n for n in links if n not in child]
            dj = []
            for i in range(6):
                ax = [0.0, 0.0, 0.0]
                ax[i % 3] = 1.0
                dj.append(dict(name=DUMMY_JOINTS[i], type="prismatic" if i < 3 else "revolute",
                               parent=DUMMY_LINKS[i], child=DUMMY_LINKS[i + 1] if i < 5 else root,
                               xyz=[0.0, 0.0, 0.0], rpy=[0.0, 0.0, 0.0], axis=ax,
                               limit=[-5.0, 5.0] if i < 3 else [-2 * math.pi, 2 * math.pi]))
            joints = dj + joints
            links = DUMMY_LINKS + links
        self.urdf_joints = joints
        self.link_names = links
        child = {j["child"] for j in joints}
        (self.root,) = [n for n in links if n not in child]
        self._out = {n: sorted((j for j in joints if j["parent"] == n), key=lambda j: j["name"]) for n in links}
        # pinocchio DoF order: preorder DFS
        order = []
        self._chains = {self.root: ()}  # link -> names of the movable joints between the root and it

        def walk(link):
            for j in self._out[link]:
                chain = self._chains[link]
                if j["type"] in ("revolute", "prismatic"):
                    order.append(j["name"])
                    chain = chain + (j["name"],)
                elif j["type"] != "fixed":
                    raise NotImplementedError("Can not handle robot with special joint.")
                self._chains[j["child"]] = chain
                walk(j["child"])

        walk(self.root)
        self.dof_joint_names = order
        self.dof = len(order)
        self._dof_index = {n: i for i, n in enumerate(order)}
        jm = {j["name"]: j for j in joints}
        self.joint_limits = np.array([jm[n].get("limit", [0.0, 0.0]) for n in order], dtype=np.float64).reshape(-1, 2)
        self._T0 = {j["name"]: _origin(j["xyz"], j["rpy"]) for j in joints}
        self._axis = {j["name"]: np.asarray(j["axis"], float) / np.linalg.norm(j["axis"]) for j in joints
                      if j["type"] != "fixed"}
        self._poses = None
        self._joint_frames = None
        self._c = None
        lib = _load_c() if use_c in (None, True) else None
        if use_c is True and lib is None:
            raise RuntimeError("oracle C kinematics requested but liborc.so could not be built")
        if lib is not None:
            self._init_c(lib)

    # -- C restatement (oracle/c/orc.c): same arithmetic, used for speed ----------------------------
    def _init_c(self, lib):
        order, stack = [], [self.root]
        lidx = {self.root: 0}
        while stack:  # topological order of joints, links numbered as they are reached
            link = stack.pop()
            for j in self._out[link]:
                lidx[j["child"]] = len(lidx)
                order.append(j)
                stack.append(j["child"])
        jidx = {j["name"]: i for i, j in enumerate(order)}
        tmap = {"fixed": 0, "revolute": 1, "prismatic": 2}
        a = dict(
            parent=np.array([lidx[j["parent"]] for j in order], dtype=np.int32),
            child=np.array([lidx[j["child"]] for j in order], dtype=np.int32),
            type=np.array([tmap[j["type"]] for j in order], dtype=np.int32),
            dofi=np.array([self._dof_index.get(j["name"], -1) for j in order], dtype=np.int32),
            origin=np.ascontiguousarray([self._T0[j["name"]] for j in order], dtype=np.float64),
            axis=np.ascontiguousarray([self._axis.get(j["name"], np.zeros(3)) for j in order], dtype=np.float64),
        )
        start, flat = [0] * (len(lidx) + 1), []
        by_cidx = sorted(lidx, key=lidx.get)
        for k, name in enumerate(by_cidx):
            flat += [jidx[n] for n in self._chains[name]]
            start[k + 1] = len(flat)
        a["cstart"] = np.array(start, dtype=np.int32)
        a["cjoint"] = np.array(flat if flat else [0], dtype=np.int32)
        m = _OrcModel(len(lidx), len(order), self.dof, *(a[k].ctypes.data for k in ("parent", "child", "type", "dofi", "origin", "axis", "cstart", "cjoint")))
        self._c = dict(lib=lib, model=m, arrays=a, lidx=lidx,
                       poses=np.zeros((len(lidx), 16)), axis_w=np.zeros((len(order), 3)), origin_w=np.zeros((len(order), 3)),
                       link2c=np.array([lidx[n] for n in self.link_names], dtype=np.int32))

    def _c_ids(self, link_ids):
        return np.ascontiguousarray(self._c["link2c"][np.asarray(link_ids, dtype=np.int64)], dtype=np.int32)

    # -- names ---------------------------------------------------------------------------------
    def get_joint_index(self, name):
        return self.dof_joint_names.index(name)

    def get_link_index(self, name):
        if name not in self.link_names:
            raise ValueError(f"{name} is not a link name. Valid link names: \n{self.link_names}")
        return self.link_names.index(name)

    def mimic_spec(self):
        """retargeting_config.py:265-285 (joint_map iteration = file order)."""
        src, mim, mul, off = [], [], [], []
        for j in self.urdf_joints:
            if "mimic" in j:
                mim.append(j["name"]); src.append(j["mimic"][0]); mul.append(j["mimic"][1]); off.append(j["mimic"][2])
        return src, mim, mul, off

    # -- kinematics ----------------------------------------------------------------------------
    def compute_forward_kinematics(self, qpos):
        qpos = np.ascontiguousarray(qpos, dtype=np.float64)
        assert qpos.shape == (self.dof,)
        if self._c is not None:
            c = self._c
            c["lib"].orc_fk(C.byref(c["model"]), qpos.ctypes.data_as(C.c_void_p), c["poses"].ctypes.data_as(C.c_void_p),
                            c["axis_w"].ctypes.data_as(C.c_void_p), c["origin_w"].ctypes.data_as(C.c_void_p))
            return
        poses = {self.root: np.eye(4)}
        frames = {}  # movable joint name -> (world axis, world origin, type, chain of dof indices)
        stack = [self.root]
        while stack:
            link = stack.pop()
            for j in self._out[link]:
                T = poses[link] @ self._T0[j["name"]]
                if j["type"] != "fixed":
                    i = self._dof_index[j["name"]]
                    a = self._axis[j["name"]]
                    frames[j["name"]] = (T[:3, :3] @ a, T[:3, 3].copy(), j["type"], i)
                    M = np.eye(4)
                    if j["type"] == "revolute":
                        M[:3, :3] = _rodrigues(a, qpos[i])
                    else:
                        M[:3, 3] = a * qpos[i]
                    T = T @ M
                poses[j["child"]] = T
                stack.append(j["child"])
        self._poses, self._joint_frames = poses, frames

    def get_link_pose(self, link_id):
        if self._c is not None:
            return self._c["poses"][self._c["link2c"][link_id]].reshape(4, 4).copy()
        return self._poses[self.link_names[link_id]].copy()

    def get_link_pose_inv(self, link_id):
        return np.linalg.inv(self.get_link_pose(link_id))

    def _c_call(self, fn, link_ids, *extra):
        c = self._c
        ids = self._c_ids(link_ids)
        getattr(c["lib"], fn)(C.byref(c["model"]), c["poses"].ctypes.data_as(C.c_void_p), c["axis_w"].ctypes.data_as(C.c_void_p),
                              c["origin_w"].ctypes.data_as(C.c_void_p), C.c_int(len(ids)), ids.ctypes.data_as(C.c_void_p), *extra)

    def link_positions(self, link_ids):
        if self._c is not None:
            pos = np.empty((len(link_ids), 3))
            self._c_call("orc_positions_jacobians", link_ids, pos.ctypes.data_as(C.c_void_p), C.c_void_p(None))
            return pos
        return np.stack([self._poses[self.link_names[i]][:3, 3] for i in link_ids], axis=0)

    def link_jacobians(self, link_ids):
        """(len(link_ids), 3, dof): world-aligned linear Jacobian of each link origin."""
        if self._c is not None:
            pos = np.empty((len(link_ids), 3))
            J = np.empty((len(link_ids), 3, self.dof))
            self._c_call("orc_positions_jacobians", link_ids, pos.ctypes.data_as(C.c_void_p), J.ctypes.data_as(C.c_void_p))
            return J
        J = np.zeros((len(link_ids), 3, self.dof))
        for r, lid in enumerate(link_ids):
            name = self.link_names[lid]
            p = self._poses[name][:3, 3]
            for jn in self._chains[name]:
                a, o, typ, i = self._joint_frames[jn]
                J[r, :, i] = np.cross(a, p - o) if typ == "revolute" else a
        return J


    def link_position_hessian_contraction(self, link_ids, gpos):
        """sum_l sum_c gpos[l,c] * d2 p_l[c] / dq_i dq_j  as a (dof, dof) matrix (float64, explicit loops).

        Second derivatives of a point rigidly attached below joints i (closer to the root) and j:
          i revolute, j revolute : a_i x (a_j x (p - o_j))
          i revolute, j prismatic: a_i x a_j
          i prismatic            : 0
        (a = world axis, o = world joint origin).  Used by the oracle's Newton polish only."""
        if self._c is not None:
            S = np.empty((self.dof, self.dof))
            g = np.ascontiguousarray(gpos, dtype=np.float64)
            self._c_call("orc_hessian_contraction", link_ids, g.ctypes.data_as(C.c_void_p), S.ctypes.data_as(C.c_void_p))
            return S
        S = np.zeros((self.dof, self.dof))
        for r, lid in enumerate(link_ids):
            name = self.link_names[lid]
            p = self._poses[name][:3, 3]
            chain = self._chains[name]
            g = gpos[r]
            for u, ji in enumerate(chain):
                ai, oi, ti, i = self._joint_frames[ji]
                if ti != "revolute":
                    continue
                for jj in chain[u:]:
                    aj, oj, tj, j = self._joint_frames[jj]
                    d = np.cross(ai, np.cross(aj, p - oj)) if tj == "revolute" else np.cross(ai, aj)
                    v = float(g @ d)
                    S[i, j] += v
                    if i != j:
                        S[j, i] += v
        return S


class OracleMimic:
    """kinematics_adaptor.py:46-113."""

    def __init__(self, robot, target_joint_names, source_names, mimic_names, multipliers, offsets):
        inter = set(mimic_names) & set(target_joint_names)
        if inter:
            raise ValueError("Mimic joint should not be one of the target joints.")
        self.idx_pin2target = np.array([robot.get_joint_index(n) for n in target_joint_names])
        self.idx_pin2source = np.array([robot.get_joint_index(n) for n in source_names])
        self.idx_pin2mimic = np.array([robot.get_joint_index(n) for n in mimic_names])
        self.idx_target2source = np.array([list(target_joint_names).index(n) for n in source_names])
        self.multipliers = np.asarray(multipliers, float)
        self.offsets = np.asarray(offsets, float)

    def forward_qpos(self, q):
        q[self.idx_pin2mimic] = q[self.idx_pin2source] * self.multipliers + self.offsets
        return q

    def backward_jacobian(self, J):
        Jt = J[..., self.idx_pin2target].copy()
        Jm = J[..., self.idx_pin2mimic] * self.multipliers
        for i, t in enumerate(self.idx_target2source):
            Jt[..., t] += Jm[..., i]
        return Jt
