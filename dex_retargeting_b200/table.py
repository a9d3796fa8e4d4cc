"""Robot-table compiler: KinematicModel + optimizer wiring -> `dexr_table_t` (include/dexr.h).

Everything the reference keeps in Python objects and re-walks per objective evaluation is flattened
here once, at init time, into one ~8 KB plain-old-data struct that the kernel reads:
  * pinocchio Model (robot_wrapper.py:15-23): joint placements with fixed joints folded, axes,
    types; ancestor masks and pointer-jumping tables for the in-kernel forward kinematics
  * Optimizer index maps idx_pin2target / idx_pin2fixed (optimizer.py:25-38, 65-75)
  * nlopt bounds (optimizer.py:54-60) and SeqRetargeting's clip limits (seq_retarget.py:20-31)
  * MimicJointKinematicAdaptor tables (kinematics_adaptor.py:46-100)
  * per-optimizer link lists and human keypoint indices (optimizer.py:132-134, 226-237, 361-395)
(paths relative to /root/reference/src/dex_retargeting)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _native as N
from .urdf import KinematicModel


@dataclass
class ObjectiveSpec:
    """What the objective looks at (all in terms of link NAMES and keypoint ids)."""

    loss: int  # N.LOSS_*
    link_names: List[str]  # computed links, slot order
    res_task: List[int]  # slot per residual block
    res_origin: List[int]  # slot or -1
    res_human_task: List[int]
    res_human_origin: List[int]
    num_fingers: int = 0
    len_proj: int = 0
    len_s1: int = 0
    s2_origin: List[int] = field(default_factory=list)
    s2_task: List[int] = field(default_factory=list)


def _skew(a):
    return np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]], dtype=np.float64)


def compile_table(
    model: KinematicModel,
    target_joint_names: Sequence[str],
    objective: ObjectiveSpec,
    target_limits: np.ndarray,  # (n_var, 2) un-widened limits in target order
    epsilon: float = 1e-3,
    mimic: Optional[Tuple[Sequence[str], Sequence[str], Sequence[float], Sequence[float]]] = None,
    fixed_joint_names: Optional[Sequence[str]] = None,
) -> N.DexrTable:
    dof = model.dof
    if dof > N.MAX_LANES:
        raise ValueError(f"robot has {dof} movable joints; the solver supports at most {N.MAX_LANES}")
    names = model.dof_joint_names
    n_var = len(target_joint_names)
    if n_var < 1:
        raise ValueError("at least one target joint is required")
    target_limits = np.asarray(target_limits, dtype=np.float64)
    if target_limits.shape != (n_var, 2):
        raise ValueError(f"Expect joint limits have shape: {(n_var, 2)}, but get {target_limits.shape}")

    t = N.DexrTable()
    t.magic = N.TABLE_MAGIC
    t.nbytes = N.C.sizeof(N.DexrTable)
    t.dof, t.n_var = dof, n_var
    t.loss = int(objective.loss)

    var_index = [-1] * dof
    for k, jn in enumerate(target_joint_names):
        if jn not in names:
            raise ValueError(f"Joint {jn} given does not appear to be in robot XML.")
        if var_index[names.index(jn)] != -1:
            raise ValueError(f"Joint {jn} appears twice in target_joint_names")
        var_index[names.index(jn)] = k

    mimic_src = [-1] * dof
    mimic_mult = [0.0] * dof
    mimic_off = [0.0] * dof
    if mimic is not None:
        src_names, mim_names, mults, offs = mimic
        for s, mname, mu, of in zip(src_names, mim_names, mults, offs):
            mi = names.index(mname)
            si = names.index(s)
            if var_index[mi] != -1:
                raise ValueError("Mimic joint should not be one of the target joints.")
            if var_index[si] == -1:
                raise ValueError(f"Mimic source joint {s} must be one of the target joints")
            mimic_src[mi], mimic_mult[mi], mimic_off[mi] = si, float(mu), float(of)

    fixed_lanes = [i for i in range(dof) if var_index[i] == -1 and mimic_src[i] == -1]
    if fixed_joint_names is not None:
        expect = [names.index(n) for n in fixed_joint_names]
        if expect != fixed_lanes:
            raise ValueError("fixed joint list inconsistent with target / mimic joints")
    t.n_fixed = len(fixed_lanes)
    fixed_index = [-1] * dof
    for k, i in enumerate(fixed_lanes):
        fixed_index[i] = k

    anc = model.is_ancestor_table()  # anc[i, j]: j is i or an ancestor of i
    for c in range(N.MAX_LANES):
        if c < dof:
            R0 = model.joint_R[c]
            a = model.joint_axis[c]
            K = _skew(a)
            rev = model.joint_type[c] == 0
            RA = R0 @ K if rev else np.zeros((3, 3))
            RB = R0 @ K @ K if rev else np.zeros((3, 3))
            p0 = model.joint_p[c]
            d0 = R0 @ a
            jt = int(model.joint_type[c])
            am = sum(1 << j for j in range(dof) if anc[c, j])
            dm = sum(1 << i for i in range(dof) if anc[i, c])
        else:  # unused lanes: identity, prismatic with zero direction, never an ancestor of anything
            R0, RA, RB, p0, d0, a = np.eye(3), np.zeros((3, 3)), np.zeros((3, 3)), np.zeros(3), np.zeros(3), np.zeros(3)
            jt, am, dm = 1, (1 << c), (1 << c)
        for k in range(9):
            t.R0[c][k], t.RA[c][k], t.RB[c][k] = R0.flat[k], RA.flat[k], RB.flat[k]
        for k in range(3):
            t.p0[c][k], t.d0[c][k], t.axis[c][k] = p0[k], d0[k], a[k]
        t.jtype[c] = jt
        t.anc_mask[c], t.desc_mask[c] = am, dm
        t.var_index[c] = var_index[c] if c < dof else -1
        t.fixed_index[c] = fixed_index[c] if c < dof else -1
        t.mimic_src[c] = mimic_src[c] if c < dof else -1
        t.mimic_mult[c] = mimic_mult[c] if c < dof else 0.0
        t.mimic_off[c] = mimic_off[c] if c < dof else 0.0
        lo, hi, clo, chi = 0.0, 0.0, 0.0, 0.0
        if c < dof and var_index[c] >= 0:
            clo, chi = target_limits[var_index[c]]
            lo, hi = clo - epsilon, chi + epsilon
        t.lower[c], t.upper[c], t.clip_lo[c], t.clip_hi[c] = lo, hi, clo, chi

    # pointer jumping: 2^r-th ancestor
    ptr = [int(model.joint_parent[c]) for c in range(dof)]
    max_depth = int(model.joint_depth.max()) if dof else 1
    n_rounds = max(0, math.ceil(math.log2(max_depth))) if max_depth > 1 else 0
    if n_rounds > 5:
        raise ValueError("kinematic chains deeper than 32 joints are not supported")
    jump = [0] * N.MAX_LANES
    cur = list(ptr)
    for r in range(5):
        for c in range(N.MAX_LANES):
            v = cur[c] if (c < dof and r < n_rounds and cur[c] >= 0) else 63
            jump[c] |= (v & 63) << (6 * r)
        cur = [(cur[cur[c]] if cur[c] >= 0 else -1) for c in range(dof)]
    for c in range(N.MAX_LANES):
        t.jump[c] = jump[c]
    t.n_rounds = n_rounds

    # variable groups (the variable's own lane first, then the mimic joints it drives)
    has_mimic = 0
    for c in range(N.MAX_LANES):
        lanes, mults = [], []
        if c < dof and var_index[c] >= 0:
            lanes, mults = [c], [1.0]
            for j in range(dof):
                if mimic_src[j] == c:
                    lanes.append(j)
                    mults.append(mimic_mult[j])
                    has_mimic = 1
        if len(lanes) > N.MAX_GROUP:
            raise ValueError(f"joint {names[c]} drives {len(lanes) - 1} mimic joints; at most {N.MAX_GROUP - 1} supported")
        t.group_count[c] = len(lanes)
        for f in range(N.MAX_GROUP):
            t.group_lane[c][f] = lanes[f] if f < len(lanes) else 0
            t.group_mult[c][f] = mults[f] if f < len(mults) else 0.0
    t.has_mimic = has_mimic

    # links
    L = len(objective.link_names)
    if not 1 <= L <= N.MAX_LINKS:
        raise ValueError(f"objective uses {L} links; supported 1..{N.MAX_LINKS}")
    t.n_links = L
    for k, ln in enumerate(objective.link_names):
        li = model.link_index(ln)
        par = int(model.link_parent[li])
        t.link_parent[k] = par
        for e in range(3):
            t.link_off[k][e] = model.link_p[li][e]
        t.link_anc_mask[k] = t.anc_mask[par] if par >= 0 else 0

    for c in range(dof):
        riders = [objective.link_names[k] for k in range(L) if t.link_parent[k] == c]
        if len(riders) > N.MAX_LINKS_PER_LANE:
            raise ValueError(f"{len(riders)} objective links ({riders}) are attached to joint {names[c]}; "
                             f"at most {N.MAX_LINKS_PER_LANE} per joint are supported")

    m = len(objective.res_task)
    if not 1 <= m <= N.MAX_RES:
        raise ValueError(f"objective has {m} residual blocks; supported 1..{N.MAX_RES}")
    t.n_res = m
    for k in range(N.MAX_RES):
        if k < m:
            ht, ho = int(objective.res_human_task[k]), int(objective.res_human_origin[k])
            if not 0 <= ht < N.NUM_KEYPOINTS or not -1 <= ho < N.NUM_KEYPOINTS:
                raise ValueError("target_link_human_indices must index the 21 hand keypoints")
            t.res_task[k], t.res_origin[k] = int(objective.res_task[k]), int(objective.res_origin[k])
            t.res_human_task[k], t.res_human_origin[k] = ht, ho
        else:
            t.res_task[k], t.res_origin[k], t.res_human_task[k], t.res_human_origin[k] = 0, -1, 0, -1
    t.block_width = _block_width(t, dof, n_var, has_mimic)
    t.arrow = 0 if t.block_width else _arrow(t, dof, n_var, has_mimic)
    t.num_fingers, t.len_proj, t.len_s1 = objective.num_fingers, objective.len_proj, objective.len_s1
    for k in range(N.MAX_RES):
        t.s2_origin[k] = objective.s2_origin[k] if k < len(objective.s2_origin) else 0
        t.s2_task[k] = objective.s2_task[k] if k < len(objective.s2_task) else 0
    return t


def _block_width(t: N.DexrTable, dof: int, n_var: int, has_mimic: int) -> int:
    """Width (4 or 8) of the aligned lane windows over which the Newton system is block diagonal, or 0 (dense).

    Two joints are coupled if one is an ancestor of the other (kinematic curvature, shared Jacobian rows) or if
    some residual block depends on both.  When every coupled set sits inside one aligned window -- e.g. the four
    4-joint fingers of Allegro / LEAP with a palm-fixed origin link -- the solver factorises all windows side by
    side (dexr_kernels.cuh, "block mode")."""
    if has_mimic or n_var != dof:
        return 0
    parent = list(range(dof))

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    def union_mask(mask):
        lanes = [i for i in range(dof) if (mask >> i) & 1]
        for b in lanes[1:]:
            parent[find(b)] = find(lanes[0])

    for c in range(dof):
        union_mask(int(t.anc_mask[c]))
    for k in range(t.n_res):
        m = int(t.link_anc_mask[t.res_task[k]])
        if t.res_origin[k] >= 0:
            m |= int(t.link_anc_mask[t.res_origin[k]])
        union_mask(m)
    comps = {}
    for c in range(dof):
        comps.setdefault(find(c), []).append(c)
    for bw in (4, 8):
        if dof % bw == 0 and bw < dof and all(min(v) // bw == max(v) // bw for v in comps.values()):
            return bw
    return 0


def _arrow(t: N.DexrTable, dof: int, n_var: int, has_mimic: int) -> int:
    """1 + the number of trunk lanes if the Newton system has ARROW structure, else 0 (include/dexr.h: `arrow`).

    Trunk = lanes 0..tr-1 (a free-flying base and / or wrist joints: pinocchio's depth-first order puts the shared
    chain first); what remains must split into fingers -- contiguous lane runs of at most 8 joints headed by an
    ancestor of the rest -- such that no residual block and no ancestor relation joins two fingers.  The smallest such
    trunk (at most 8 lanes) is taken.  Only used by the 32-lane solver (dof > 16)."""
    if has_mimic or n_var != dof or dof <= 16:
        return 0
    anc = [int(t.anc_mask[c]) for c in range(dof)]
    desc = [int(t.desc_mask[c]) for c in range(dof)]
    supports = []
    for k in range(t.n_res):
        m = int(t.link_anc_mask[t.res_task[k]])
        if t.res_origin[k] >= 0:
            m |= int(t.link_anc_mask[t.res_origin[k]])
        supports.append(m)
    for tr in range(0, min(8, dof - 1) + 1):
        tmask = (1 << tr) - 1
        if any(anc[c] & ~tmask for c in range(tr)):
            continue
        finger_of, fingers, ok = {}, 0, True
        for c in range(tr, dof):
            chain = anc[c] & ~tmask
            fb = (chain & -chain).bit_length() - 1
            if fb == c:
                span = desc[c] | (1 << c)
                fw = span.bit_length() - c
                if fw > 8 or span != ((1 << fw) - 1) << c:
                    ok = False
                    break
                fingers += 1
                for i in range(c, c + fw):
                    finger_of[i] = span
            elif not (desc[fb] >> c) & 1:
                ok = False
                break
        if not ok or fingers > 6 or fingers < 2:
            continue
        if any(c not in finger_of or (anc[c] & ~tmask & ~finger_of[c]) for c in range(tr, dof)):
            continue
        if any((m & ~tmask) and ((m & ~tmask) & ~finger_of[((m & ~tmask) & -(m & ~tmask)).bit_length() - 1]) for m in supports):
            continue
        return 1 + tr
    return 0


def table_bytes(t: N.DexrTable) -> bytes:
    return bytes(memoryview(t))


def table_from_bytes(b: bytes) -> N.DexrTable:
    if len(b) != N.C.sizeof(N.DexrTable):
        raise ValueError("robot table has the wrong size")
    return N.DexrTable.from_buffer_copy(b)
