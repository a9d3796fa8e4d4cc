"""Joint-space parity of solver output against the committed oracle fixture tests/golden/bench_parity.npz
(tests/tools/gen_bench_parity.py: oracle mode B on the exact frames bench.py times).  Checker code: used by bench.py's
`parity` records and by tests/test_gpu_bench_parity.py, never by the product."""
from __future__ import annotations

from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
FIXTURE = ROOT / "tests" / "golden" / "bench_parity.npz"
TOL = 1e-4  # rad (BASELINE.json north_star: |dq|_inf < 1e-4)
_CACHE = {}


def fixture():
    if "f" not in _CACHE:
        _CACHE["f"] = np.load(FIXTURE)
    return _CACHE["f"]


def compare(tag, q, digest=None, status=None):
    """|dq|_inf per frame of `q` (the first n frames / streams of the workload, any float array shaped like the fixture's
    prefix) against the oracle.  Returns the record bench.py emits: n, median, p99, max, same_basin (fraction of frames
    within TOL -- the rest sit in another local minimum of the non-convex objective or were flagged), flagged."""
    f = fixture()
    ref = f[f"{tag}/q"]
    q = np.asarray(q, dtype=np.float64)
    n = min(len(ref), len(q))
    if n == 0:
        return {"config": tag, "n": 0}
    rec = {"config": tag, "fixture": "tests/golden/bench_parity.npz (oracle mode B, float64 KKT-polished)"}
    if digest is not None and str(f[f"{tag}/digest"]) != digest and n == len(ref):
        rec["error"] = "inputs differ from the fixture's (workload definition changed: regenerate tests/golden/bench_parity.npz)"
        return rec
    dq = np.abs(q[:n] - ref[:n].astype(np.float64))
    dq = dq.reshape(n, -1).max(1) if dq.ndim == 2 else dq.reshape(n, dq.shape[1], -1).max(2)  # streams: [S,T]
    flat = dq.reshape(-1)
    same = flat < TOL
    rec.update(n=int(flat.size), tol=TOL, median=float(np.median(flat)), p99=float(np.percentile(flat, 99)), max=float(flat.max()),
               same_basin=float(same.mean()), outside=int((~same).sum()),
               max_within_basin=float(flat[same].max()) if same.any() else None)
    if status is not None:
        st = np.asarray(status).reshape(-1)[: flat.size]
        rec["flagged"] = int(((st >> 24) != 0).sum())
        rec["flagged_outside"] = int((((st >> 24) != 0) & ~same).sum())
    return rec
