#!/usr/bin/env python
"""Design-time numpy emulation of the kernel's solver (batched, dtype selectable).

Not product code and not the oracle: a scratch model of the algorithm the CUDA kernel implements
(bounded Levenberg-Marquardt on the generalized Gauss-Newton model of the Huber objectives), used to
pick damping / active-set / stopping rules before writing CUDA and to study fp32 behaviour.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


class ProtoProblem:
    """Flattened problem description taken from an OracleOptimizer (design-time convenience)."""

    def __init__(self, o, dtype=np.float32):
        from dex_retargeting_b200.urdf import KinematicModel
        import json

        cfg = o.cfg
        stem = Path(cfg["urdf_path"]).stem
        with open(ROOT / "tests/golden/robots" / (stem + ".json")) as f:
            km = KinematicModel.from_dict(json.load(f), bool(cfg.get("add_dummy_free_joint", False)))
        assert km.dof_joint_names == o.robot.dof_joint_names
        self.km, self.o, self.dt = km, o, dtype
        self.link_ids = [km.link_index(o.robot.link_names[i]) for i in o.link_ids]
        # affine map q_pin = S x + c
        n, dof = o.opt_dof, km.dof
        S = np.zeros((dof, n))
        S[o.idx_pin2target, np.arange(n)] = 1.0
        self.off = np.zeros(dof)
        if o.adaptor is not None:
            a = o.adaptor
            for i in range(len(a.idx_pin2mimic)):
                S[a.idx_pin2mimic[i], a.idx_target2source[i]] = a.multipliers[i]
                self.off[a.idx_pin2mimic[i]] = a.offsets[i]
        self.S = S
        self.anc = km.is_ancestor_table()

    def fk(self, x, fixed):
        """x [B,n] -> link positions [B,L,3], J [B,L,3,n]"""
        km, dt = self.km, self.dt
        B = x.shape[0]
        q = x @ self.S.T.astype(dt) + self.off.astype(dt)
        if fixed is not None and fixed.size:
            q[:, self.o.idx_pin2fixed] = fixed
        Rw = np.zeros((B, km.dof, 3, 3), dt)
        pw = np.zeros((B, km.dof, 3), dt)
        aw = np.zeros((B, km.dof, 3), dt)
        for i in range(km.dof):
            par = km.joint_parent[i]
            R0 = km.joint_R[i].astype(dt)
            p0 = km.joint_p[i].astype(dt)
            if par >= 0:
                Rb = Rw[:, par] @ R0
                pb = np.einsum("bij,j->bi", Rw[:, par], p0) + pw[:, par]
            else:
                Rb = np.broadcast_to(R0, (B, 3, 3)).copy()
                pb = np.broadcast_to(p0, (B, 3)).copy()
            a = km.joint_axis[i].astype(dt)
            aw[:, i] = Rb @ a
            if km.joint_type[i] == 0:
                c, s = np.cos(q[:, i]), np.sin(q[:, i])
                K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]], dt)
                Rq = np.eye(3, dtype=dt)[None] + s[:, None, None] * K[None] + (1 - c)[:, None, None] * (K @ K)[None]
                Rw[:, i] = Rb @ Rq
                pw[:, i] = pb
            else:
                Rw[:, i] = Rb
                pw[:, i] = pb + aw[:, i] * q[:, i, None]
        L = len(self.link_ids)
        pos = np.zeros((B, L, 3), dt)
        Jp = np.zeros((B, L, 3, km.dof), dt)
        for r, l in enumerate(self.link_ids):
            par = km.link_parent[l]
            if par < 0:
                pos[:, r] = km.link_p[l].astype(dt)
                continue
            pos[:, r] = np.einsum("bij,j->bi", Rw[:, par], km.link_p[l].astype(dt)) + pw[:, par]
            for j in np.nonzero(self.anc[par])[0]:
                if km.joint_type[j] == 0:
                    # joint origin BEFORE its own rotation == after (revolute keeps the origin)
                    Jp[:, r, :, j] = np.cross(aw[:, j], pos[:, r] - pw[:, j])
                else:
                    Jp[:, r, :, j] = aw[:, j]
        J = Jp @ self.S.astype(dt)
        self._aw, self._Jp = aw, Jp
        return pos, J

    def curvature(self, gpos):
        """sum_l gpos_l . d2 p_l / dq dq  folded to variable space.  gpos [B,L,3] -> [B,n,n]"""
        km, dt = self.km, self.dt
        aw, Jp = self._aw, self._Jp
        t = np.cross(Jp.transpose(0, 1, 3, 2), gpos[:, :, None, :]).sum(1)  # [B,dof,3] = sum_l J_lj x g_l
        a_rev = aw * (km.joint_type == 0)[None, :, None]
        A = np.einsum("bic,bjc->bij", a_rev, t)  # a_i . t_j
        anc = self.anc  # anc[j,i]: i is ancestor-or-self of j
        up = A * anc.T[None]  # keep (i,j) with i in anc*(j)
        Sq = up + np.transpose(up, (0, 2, 1)) * (1 - np.eye(km.dof))[None]
        M = self.S.astype(dt)
        return np.einsum("ri,brs,sj->bij", M, Sq, M)


NOISE = 2e-6
RFAR = 0.2
LDOWN = 0.1
NO_DEC_AFTER_REJECT = False
LAMFAST = 1e-3
KEEP_LAM_ON_NOISE = True
LUP = 10.0
STATS = {'solves': 0}


def huber(d, beta):
    return np.where(d < beta, 0.5 * d * d / beta, d - 0.5 * beta)


def solve_batch(P: ProtoProblem, target, weights, fixed, x0, last, max_iter=40, tol=1e-6, ggn=True, lam0=1e-3,
                verbose=False, newton=False, hybrid=False, near=0.1, curv_far=True, pos_majorise=False, start_exact=False,
                chord=0.0, extrap=False, curv_skip_first=0, curv_pd_fallback=False, curv_after_small=0.0):
    """target [B,m,3] (already scaled / projected), weights [B,m] or None (position)."""
    o, dt = P.o, P.dt
    B, n = x0.shape
    lo, hi = o.lower.astype(dt), o.upper.astype(dt)
    nd = dt(o.norm_delta)
    beta = dt(o.huber_delta)
    x = np.clip(x0.astype(dt), lo, hi)
    last = last.astype(dt)
    target = target.astype(dt)

    def residuals(pos):
        if o.type == "position":
            return pos - target
        return pos[:, o.task_sel] - pos[:, o.origin_sel] - target

    def cost(r, x):
        if o.type == "position":
            a = np.abs(r)
            L = huber(a, beta).reshape(B, -1).mean(1)
        else:
            d = np.linalg.norm(r, axis=2)
            L = (huber(d, beta) * weights).sum(1) / dt(o.m)
        return L + nd * ((x - last) ** 2).sum(1)

    lam = np.full(B, lam0, dt)
    Afac = np.tile(np.eye(n, dtype=np.float64), (B, 1, 1))
    fac_act = np.zeros((B, n), bool)
    use_chord = np.zeros(B, bool)
    STATS['full'] = STATS.get('full', 0); STATS['chord'] = STATS.get('chord', 0)
    exact = (np.ones(B, bool) if start_exact else np.zeros(B, bool)) if hybrid else np.full(B, ggn)
    done = np.zeros(B, bool)
    iters = np.zeros(B, int)
    pos, J = P.fk(x, fixed)
    r = residuals(pos)
    F = cost(r, x)
    x_before_last = x.copy()
    curv_armed = np.zeros(B, bool)
    H_nocurv = None
    for it in range(max_iter):
        x_prev = x.copy()
        # gradient / GGN Hessian at x
        if o.type == "position":
            Jr = J.reshape(B, -1, n)
            rr = r.reshape(B, -1)
            quad = np.abs(rr) < beta
            c = dt(1.0 / rr.shape[1])
            gres = np.where(quad, rr / beta, np.sign(rr)) * c
            ex_pos = exact & (not pos_majorise)
            wrow = np.where(ex_pos[:, None], np.where(quad, 1 / beta, 0.0), 1.0 / np.maximum(np.abs(rr), beta)) * c
            g = np.einsum("br,brn->bn", gres, Jr)
            H = np.einsum("br,bri,brj->bij", wrow, Jr, Jr)
        else:
            Jv = J[:, o.task_sel] - J[:, o.origin_sel]  # B,m,3,n
            d = np.linalg.norm(r, axis=2)
            ck = weights / dt(o.m)
            quad = d < beta
            dsafe = np.maximum(d, 1e-12)
            u = r / dsafe[..., None]
            hp = np.where(quad, d / beta, 1.0)
            g = np.einsum("bk,bkc,bkcn->bn", ck * hp, u, Jv)
            a_iso = np.where(quad, 1 / beta, 1 / dsafe) * ck
            H = np.einsum("bk,bkci,bkcj->bij", a_iso, Jv, Jv)
            uJ = np.einsum("bkc,bkcn->bkn", u, Jv)
            a_rad = np.where(quad, 0.0, 1 / dsafe) * ck * exact[:, None]
            H -= np.einsum("bk,bki,bkj->bij", a_rad, uJ, uJ)
        if newton:
            if o.type == "position":
                gpos = gres.reshape(r.shape)
            else:
                gv = (ck * hp)[..., None] * u
                gpos = np.zeros_like(pos)
                for k in range(o.m):
                    gpos[:, o.task_sel[k]] += gv[:, k]
                    gpos[:, o.origin_sel[k]] -= gv[:, k]
            rmax = np.abs(r).reshape(B, -1).max(1) if o.type == 'position' else np.linalg.norm(r, axis=2).max(1)
            cw = np.ones(B, dt) if (curv_far or not hybrid) else (rmax < RFAR).astype(dt)
            if it < curv_skip_first:      # experiment: Gauss-Newton model for the first iterations
                cw = cw * 0
            if curv_after_small > 0:      # experiment: kinematic curvature only after an accepted step below this size
                cw = cw * curv_armed.astype(dt)
            Hcurv = P.curvature(gpos) * cw[:, None, None]
            H_nocurv = H + 2 * nd * np.eye(n, dtype=dt)[None]
            H = H + Hcurv
        g = g + 2 * nd * (x - last)
        H = H + 2 * nd * np.eye(n, dtype=dt)[None]
        act = ((x <= lo) & (g > 0)) | ((x >= hi) & (g < 0))
        g_f = np.where(act, 0, g)
        H_f = H * (~act)[:, :, None] * (~act)[:, None, :]
        H_f[:, np.arange(n), np.arange(n)] += act.astype(dt)
        diag = np.abs(H_f[:, np.arange(n), np.arange(n)]) + dt(1e-6)
        # chord (frozen factor) iteration for frames flagged by the previous step, if the active set is unchanged
        chord_now = use_chord & ~done & (act == fac_act).all(1)
        STATS['chord'] += int(chord_now.sum()); STATS['full'] += int((~chord_now & ~done).sum())
        use_chord = np.zeros(B, bool)
        chord_ok = np.zeros(B, bool)
        if chord_now.any():
            delta = -np.linalg.solve(Afac, g_f.astype(np.float64)[..., None])[..., 0].astype(dt)
            xn = np.clip(x + delta, lo, hi)
            posn, Jn = P.fk(xn, fixed)
            rn = residuals(posn)
            Fn = cost(rn, xn)
            step = np.abs(xn - x).max(1)
            noise = dt(NOISE) * np.abs(F)
            prev_step = np.abs(x - x_before_last).max(1)
            okc = chord_now & ((Fn <= F + noise) | (step < tol)) & (step < 0.5 * prev_step)
            chord_ok = okc
            x_before_last = np.where(okc[:, None], x, x_before_last)
            x = np.where(okc[:, None], xn, x)
            pos = np.where(okc[:, None, None], posn, pos)
            J = np.where(okc[:, None, None, None], Jn, J)
            r = np.where(okc[:, None, None], rn, r)
            F = np.where(okc, Fn, F)
            done |= okc & (step < tol)
            use_chord = okc & ~done
            iters += chord_now.astype(int)
        # inner loop: try lambdas
        accepted = done.copy() | chord_now   # chord frames skip the full solve this iteration (failed ones retry fully next time)
        switch = np.zeros(B, bool)
        for trial in range(8):
            STATS['solves'] += int((~accepted).sum())
            A = H_f.copy()
            A[:, np.arange(n), np.arange(n)] += lam[:, None] * diag
            pd = np.linalg.eigvalsh(A.astype(np.float64)).min(1) > 0
            if curv_pd_fallback and newton and H_nocurv is not None and (~pd).any():
                # experiment: indefinite with the kinematic curvature -> same damping on the PSD model without it
                A2 = H_nocurv * (~act)[:, :, None] * (~act)[:, None, :]
                A2[:, np.arange(n), np.arange(n)] += act.astype(dt) + lam[:, None] * diag
                A = np.where(pd[:, None, None], A, A2)
                pd = np.linalg.eigvalsh(A.astype(np.float64)).min(1) > 0
            if False and hybrid and start_exact and trial == 0:
                switch = exact & ~pd & ~accepted
                accepted = accepted | switch
            A[~pd] = np.eye(n, dtype=dt)
            delta = -np.linalg.solve(A.astype(np.float64), g_f.astype(np.float64)[..., None])[..., 0].astype(dt)
            delta[~pd] = 0
            xn = np.clip(x + delta, lo, hi)
            posn, Jn = P.fk(xn, fixed)
            rn = residuals(posn)
            Fn = cost(rn, xn)
            step = np.abs(xn - x).max(1)
            pred = 0.5 * ((xn - x) * (lam[:, None] * diag * (xn - x) - g_f)).sum(1)
            noise = dt(NOISE) * np.abs(F)
            ok = ((Fn <= F) | (step < tol) | (pred < noise)) & pd
            F_before = F.copy()
            upd = ok & ~accepted
            x = np.where(upd[:, None], xn, x)
            pos = np.where(upd[:, None, None], posn, pos)
            J = np.where(upd[:, None, None, None], Jn, J)
            r = np.where(upd[:, None, None], rn, r) if r.ndim == 3 else r
            F = np.where(upd, Fn, F)
            pstep = np.abs(x_prev - x_before_last).max(1)
            unclipped = (np.abs((x + delta) - xn).max(1) == 0)
            fast = extrap & (pstep > 0) & (step < 0.1 * pstep) & (step * step / np.maximum(pstep, 1e-30) < tol) & (step < 1e-3) & (lam <= dt(LAMFAST)) & unclipped & (trial == 0) & exact & (Fn <= F)
            newly_done = upd & ((step < tol) | fast)
            iters += (~done).astype(int) * (0 if trial else 1)
            done |= newly_done
            verified = Fn < F_before - noise if KEEP_LAM_ON_NOISE else np.ones(B, bool)
            dec_ok = verified & ((trial == 0) | (not NO_DEC_AFTER_REJECT))
            lam = np.where(upd & dec_ok, np.maximum(lam * dt(LDOWN), dt(1e-7)), np.where(accepted | upd, lam, lam * dt(LUP)))
            # remember the factor of accepted steps; request a chord iteration after small exact steps
            if extrap and not chord > 0:
                x_before_last = np.where(upd[:, None], x_prev, x_before_last)
            if chord > 0:
                Afac = np.where(upd[:, None, None], A.astype(np.float64), Afac)
                fac_act = np.where(upd[:, None], act, fac_act)
                x_before_last = np.where(upd[:, None], x_prev, x_before_last)
                use_chord = np.where(upd, (step < chord) & exact & ~newly_done, use_chord)
            accepted |= upd
            if accepted.all():
                break
        curv_armed = accepted & ~chord_now & (np.abs(x - x_prev).max(1) < curv_after_small)
        stuck = ~accepted & ~chord_now
        if hybrid:
            done |= stuck & ~exact  # the majoriser model failed too: at the (numerical) minimum
            laststep = np.abs(x - x_prev).max(1)
            exact = (~stuck) & (laststep < near) & ~switch
            if start_exact:
                lam = np.where(switch, dt(lam0), lam)
        else:
            done |= stuck  # could not improve: at (numerical) minimum
        if verbose:
            print(it, "active", (~done).sum(), "F mean", F.mean())
        if done.all():
            break
    return x, iters, F
