#!/usr/bin/env python
"""Generate tests/golden/bench_parity.npz: the ORACLE's converged answer (mode B: float64 minimiser of
L(x) + norm_delta |x - x_last|^2 inside the widened bounds, SLSQP + exact-Hessian projected Newton polish to a KKT residual
< 1e-10; oracle/solvers.py:solve_converged) on the EXACT frames bench.py times, for every BASELINE.json configuration:

  metric        first 4096 frames of rank 0's first input batch (Vector Allegro, warm start 0.05 sigma)   [also config 2: its
                4096-frame batch IS this prefix]
  metric_cold   first 1024 frames of the cold-start arm (0.5 sigma, tests/test_optimizer.py:28-42)
  metric_real   first 1242 frames (two passes over the recording) of the real-trajectory arm
  shadow_narrow first 4096 frames of config 3 with the dummy joints drawn from +-0.5 m / +-pi
  shadow_ship   first 1024 frames of config 3 with the shipped +-5 m / +-2 pi range
  leap_frames   first 2048 frames of the DexPilot LEAP independent-frames arm (hysteresis flags start cleared)
  leap_streams  first 16 streams x 300 frames of config 4 (DexPilot LEAP, hysteresis + low-pass carried; sequential per stream,
                oracle SeqRetargeting in mode B), filtered robot qpos like the kernel's output
  mixed/<robot> first 256 frames of each robot group of config 5

The inputs are regenerated from tools/workloads.py (seeded numpy); a digest of every input slice is stored so that bench.py and
the tests can tell when the workload definition moved away from the fixture.  ~35 000 float64 solves: about a minute on 8 cores.

Usage: python tests/tools/gen_bench_parity.py
"""
import multiprocessing as mp
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT / "tools"))
import workloads as W  # noqa: E402

_O = {}


def _oracle(key):
    if key not in _O:
        from helpers import build_oracle

        _O[key] = build_oracle(key)
    return _O[key]


def _solve_frames(args):
    from oracle.solvers import solve_converged

    key, kp, x0, fixed = args
    o = _oracle(key)
    out, Fs, kk = [], [], []
    for i in range(kp.shape[0]):
        if hasattr(o, "projected") and o.type == "dexpilot":
            o.projected[:] = False
        ref = o.ref_from_keypoints(kp[i])
        fx = fixed[i] if fixed is not None else np.zeros(0)
        x, kkt, F = solve_converged(o, ref, fx, x0[i], update_state=False)
        out.append(x); Fs.append(F); kk.append(kkt)
    return np.array(out), np.array(Fs), np.array(kk)


def _solve_stream(args):
    from oracle.solvers import OracleSeqRetargeting

    key, kp = args
    o = _oracle(key)
    if o.type == "dexpilot":
        o.projected[:] = False
    seq = OracleSeqRetargeting(o, mode="converged")
    return np.array([seq.retarget(o.ref_from_keypoints(f)) for f in kp])


def frames_case(pool, key, kp, x0, fixed, n):
    kp, x0 = kp[:n], x0[:n]
    fixed = fixed[:n] if fixed is not None else None
    chunks = np.array_split(np.arange(n), min(n, 8 * 8))
    parts = pool.map(_solve_frames, [(key, kp[c], x0[c], fixed[c] if fixed is not None else None) for c in chunks])
    q = np.concatenate([p[0] for p in parts])
    return dict(q=q.astype(np.float32), F=np.concatenate([p[1] for p in parts]), kkt=np.concatenate([p[2] for p in parts]),
                digest=np.array(W.digest(kp, x0, fixed)), n=np.array(n))


def main():
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = "1"
    dst = ROOT / "tests" / "golden" / "bench_parity.npz"
    out = {}
    t0 = time.time()
    with mp.get_context("fork").Pool(8) as pool:
        def add(tag, rec):
            for k, v in rec.items():
                out[f"{tag}/{k}"] = v
            print(f"{tag}: {int(rec.get('n', 0))} frames, worst KKT residual {float(np.max(rec['kkt'])) if 'kkt' in rec else float('nan'):.2e}, "
                  f"{time.time() - t0:.0f} s", flush=True)

        seq = W.build(W.METRIC_KEY)
        kp, x0, fixed, _ = W.frames(seq, 65536, W.METRIC_SEED)
        add("metric", frames_case(pool, W.METRIC_KEY, kp, x0, fixed, 4096))
        kp, x0, fixed, _ = W.frames(seq, 65536, W.METRIC_SEED, sigma=0.5)
        add("metric_cold", frames_case(pool, W.METRIC_KEY, kp, x0, fixed, 1024))
        kp, x0 = W.real_frames(seq, 65536)
        add("metric_real", frames_case(pool, W.METRIC_KEY, kp, x0, None, 1242))

        seq = W.build(W.SHADOW_POS_KEY)
        kp, x0, fixed, _ = W.frames(seq, 65536, W.SHADOW_SEED, narrow_dummy=True)
        add("shadow_narrow", frames_case(pool, W.SHADOW_POS_KEY, kp, x0, fixed, 4096))
        kp, x0, fixed, _ = W.frames(seq, 65536, W.SHADOW_SEED, narrow_dummy=False)
        add("shadow_ship", frames_case(pool, W.SHADOW_POS_KEY, kp, x0, fixed, 1024))

        seq = W.build(W.LEAP_DEXPILOT_KEY)
        kp, x0, fixed, _ = W.frames(seq, 65536, W.SHADOW_SEED)
        add("leap_frames", frames_case(pool, W.LEAP_DEXPILOT_KEY, kp, x0, fixed, 2048))
        S = 16
        kp = W.streams(2048, 300)[:S]
        rq = np.array(pool.map(_solve_stream, [(W.LEAP_DEXPILOT_KEY, kp[s]) for s in range(S)]))
        add("leap_streams", dict(q=rq.astype(np.float32), digest=np.array(W.digest(kp)), n=np.array(S * 300)))

        for i, key in enumerate(W.MIXED_KEYS):
            seq = W.build(key)
            kp, x0, fixed, _ = W.frames(seq, 16384, W.MIXED_SEED + i)
            add("mixed/" + key.split("/")[1], frames_case(pool, key, kp, x0, fixed, 256))
    np.savez_compressed(dst, **out)
    print("wrote", dst, dst.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
