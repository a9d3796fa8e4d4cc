#!/usr/bin/env python
"""CPU probes of the solver's stopping and damping rules on the bench workloads, using the fp32 numpy model of the
kernel's iteration (tools/lm_prototype.py).  Design-time only: nothing here is a measurement of the CUDA path.

  python tests/tools/solver_model_probe.py tol   teleop/allegro_hand_right 2048   # stopping-threshold sweep
  python tests/tools/solver_model_probe.py lam   offline/shadow_hand_right 512    # damping-schedule sweep
  python tests/tools/solver_model_probe.py traj  offline/shadow_hand_right 256    # step / error per iteration
  python tests/tools/solver_model_probe.py start offline/shadow_hand_right 256    # Hessian spectrum at the warm start
  python tests/tools/solver_model_probe.py curv  offline/shadow_hand_right 256 [target noise, m]   # curvature gating rules
  python tests/tools/solver_model_probe.py stream teleop/allegro_hand_right 200  # same rules, recorded stream, warm starts

Numbers quoted in DESIGN.md section 5c come from these commands.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
for p in (ROOT, ROOT / "tests", ROOT / "tools"):
    sys.path.insert(0, str(p))
import lm_prototype as LP  # noqa: E402
from bench_configs import synth  # noqa: E402
from helpers import build_oracle, build_product  # noqa: E402

KERNEL = dict(newton=True, hybrid=True, start_exact=True, curv_far=False, pos_majorise=True)  # what the kernel does


def problem(key, B, seed=100, dtype=np.float32):
    seq, o = build_product(key), build_oracle(key)
    kp, x0, fixed, _ = synth(seq, B, seed)
    refs = np.stack([o.ref_from_keypoints(kp[i]) for i in range(B)]).astype(dtype)
    target, weights = frame_constants(o, refs, dtype, stream=False)
    fx = (fixed if fixed is not None else np.zeros((B, 0))).astype(dtype)
    return LP.ProtoProblem(o, dtype), o, target, weights, fx, x0.astype(dtype)


def frame_constants(o, refs, dtype, stream):
    """Targets and residual weights of every frame (oracle `prepare`: scaling; DexPilot flags, weights, projected targets --
    flags carried from frame to frame when `stream`, otherwise every frame starts with none set)."""
    if o.type == "position":
        return refs.astype(dtype), None
    tw = []
    for r in refs:
        if o.type == "dexpilot" and not stream:
            o.projected[:] = False
        tw.append(o.prepare(r, update_state=True))
    return np.stack([t for t, _ in tw]).astype(dtype), np.stack([w for _, w in tw]).astype(dtype)


def solve(P, target, weights, fx, x0, **kw):
    LP.STATS["solves"] = 0
    x, it, F = LP.solve_batch(P, target, weights, fx, x0.copy(), x0.copy(), **{**KERNEL, **kw})
    return x, it, F, LP.STATS["solves"] / x0.shape[0]


def cmd_tol(key, B):
    P, o, target, weights, fx, x0 = problem(key, B)
    P64, _, t64, w64, f64, x64 = problem(key, B, dtype=np.float64)
    xr = solve(P64, t64, w64, f64, x64, max_iter=64, lam0=1e-2, tol=1e-9)[0]
    for tol in (1e-5, 3e-5, 1e-4, 3e-4):
        x, it, F, s = solve(P, target, weights, fx, x0, max_iter=64, lam0=1e-2, tol=tol)
        d = np.abs(x - xr).max(1)
        print(f"tol {tol:.0e}: mean iterations {it.mean():.3f} (max {it.max()}), solves/frame {s:.3f}, |dq|inf vs f64-converged "
              f"median {np.median(d):.2e} p99 {np.percentile(d, 99):.2e} max {d.max():.2e}", flush=True)


def cmd_lam(key, B):
    P, o, target, weights, fx, x0 = problem(key, B)
    for lam0 in (1e-5, 1e-3, 1e-2, 1e-1, 1.0):
        for down in (0.1, 0.01):
            LP.LDOWN = down
            x, it, F, s = solve(P, target, weights, fx, x0, max_iter=64, lam0=lam0, tol=1e-5)
            print(f"lambda0 {lam0:.0e} decay {down}: mean iterations {it.mean():.3f} (max {it.max()}), solves/frame {s:.3f}, "
                  f"F mean {F.mean():.4e}", flush=True)
    LP.LDOWN = 0.1


def cmd_traj(key, B, iters=12):
    P, o, target, weights, fx, x0 = problem(key, B)
    xs = [np.clip(x0, o.lower, o.upper).astype(np.float32)]
    for k in range(1, iters + 1):  # the model is deterministic: re-run with a growing iteration cap
        xs.append(solve(P, target, weights, fx, x0, max_iter=k, lam0=1e-2, tol=1e-5)[0])
    xs = np.array(xs)
    steps, err = np.abs(xs[1:] - xs[:-1]).max(2), np.abs(xs - xs[-1][None]).max(2)
    np.set_printoptions(precision=2, linewidth=220)
    print("median step per iteration        ", np.median(steps, 1))
    print("p90 step per iteration           ", np.percentile(steps, 90, 1))
    print("median |x - x_final| by iteration", np.median(err, 1))
    print("p90 |x - x_final| by iteration   ", np.percentile(err, 90, 1))


def cmd_start(key, B):
    P, o, target, weights, fx, x0 = problem(key, B, dtype=np.float64)
    assert o.type == "position"
    x = np.clip(x0, o.lower, o.upper)
    pos, J = P.fk(x, fx)
    n = x.shape[1]
    rr, Jr = (pos - target).reshape(B, -1), J.reshape(B, -1, n)
    beta, c = o.huber_delta, 1.0 / rr.shape[1]
    quad = np.abs(rr) < beta
    gres = np.where(quad, rr / beta, np.sign(rr)) * c
    Hgn = np.einsum("br,bri,brj->bij", c / np.maximum(np.abs(rr), beta), Jr, Jr) + 2 * o.norm_delta * np.eye(n)
    H = Hgn + P.curvature(gres.reshape(pos.shape))
    ev, evgn = np.linalg.eigvalsh(H), np.linalg.eigvalsh(Hgn)
    print(f"residual coordinates beyond huber_delta at the warm start: {(~quad).mean():.3f}; median max|r| {np.median(np.abs(rr).max(1)):.3f} m")
    print(f"exact Hessian: min eigenvalue median {np.median(ev[:, 0]):.3e}, indefinite in {(ev[:, 0] < 0).mean():.3f} of the frames; "
          f"Gauss-Newton + regulariser: min {np.median(evgn[:, 0]):.3e}, max {np.median(evgn[:, -1]):.3e}")


CURV_RULES = [("kernel", {}), ("skip first 1", dict(curv_skip_first=1)), ("skip first 2", dict(curv_skip_first=2)),
              ("after step<.05", dict(curv_after_small=0.05)), ("pd fallback", dict(curv_pd_fallback=True))]


def _compare_rules(P, o, target, weights, fx, x0, skip=0):
    P64 = LP.ProtoProblem(o, np.float64)
    w64 = None if weights is None else weights.astype(np.float64)
    xr, _, Fr, _ = solve(P64, target.astype(np.float64), w64, fx.astype(np.float64), x0.astype(np.float64), max_iter=100, lam0=1e-2, tol=1e-9)
    for name, kw in CURV_RULES:
        x, it, F, s = solve(P, target, weights, fx, x0, max_iter=64, lam0=1e-2, tol=1e-5, **kw)
        d = np.abs(x - xr).max(1)
        print(f"  {name:15s}: mean iterations {it[skip:].mean():.3f} (max {it[skip:].max()}), solves/frame {s:.3f}, F - F_ref max "
              f"{np.max(F - Fr):.2e}, |dq|inf p99 {np.percentile(d, 99):.2e}, other basin {(d > 1e-4).mean():.3f}", flush=True)


def cmd_curv(key, B, noise=0.0):
    """When to include the kinematic (FK second derivative) curvature: the kernel's rule (residual below kFarResidual),
    Gauss-Newton for the first iterations, only after a small accepted step, or always but dropped when the damped
    Hessian is not positive definite (DEXR_EXP_PDFALLBACK).  `noise`: unreachable targets (std, metres)."""
    P, o, target, weights, fx, x0 = problem(key, B)
    target = (target + np.random.RandomState(5).randn(*target.shape) * float(noise)).astype(np.float32)
    print(key, "bench-style problems, target noise", noise)
    _compare_rules(P, o, target, weights, fx, x0)


def cmd_stream(key, T):
    """The same rules on the recorded keypoint trajectory, every frame warm-started from the previous frame's solution."""
    from helpers import keypoint_trajectory
    o = build_oracle(key)
    kp = keypoint_trajectory()[:T].astype(np.float32)
    refs = np.stack([o.ref_from_keypoints(kp[i]) for i in range(T)]).astype(np.float32)
    target, weights = frame_constants(o, refs, np.float32, stream=True)
    P, fx = LP.ProtoProblem(o, np.float32), np.zeros((T, 0), np.float32)
    x, starts = ((o.lower + o.upper) / 2).astype(np.float32)[None], []
    for t in range(T):
        starts.append(x[0].copy())
        x = solve(P, target[t:t + 1], None if weights is None else weights[t:t + 1], fx[:1], x, max_iter=64, lam0=1e-2, tol=1e-5)[0]
    print(key, "recorded stream,", T, "frames")
    _compare_rules(P, o, target, weights, fx, np.array(starts), skip=1)


if __name__ == "__main__":
    cmds = {"tol": cmd_tol, "lam": cmd_lam, "traj": cmd_traj, "start": cmd_start, "curv": cmd_curv, "stream": cmd_stream}
    cmds[sys.argv[1]](sys.argv[2], int(sys.argv[3]), *sys.argv[4:])
