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
import json
import math
import xml.etree.ElementTree as ET

import numpy as np

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
    def __init__(self, desc, add_dummy_free_joint=False):
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
        qpos = np.asarray(qpos, dtype=np.float64)
        assert qpos.shape == (self.dof,)
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
        return self._poses[self.link_names[link_id]].copy()

    def get_link_pose_inv(self, link_id):
        return np.linalg.inv(self._poses[self.link_names[link_id]])

    def link_positions(self, link_ids):
        return np.stack([self._poses[self.link_names[i]][:3, 3] for i in link_ids], axis=0)

    def link_jacobians(self, link_ids):
        """(len(link_ids), 3, dof): world-aligned linear Jacobian of each link origin."""
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
