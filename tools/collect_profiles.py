#!/usr/bin/env python
"""Copy the record artefacts of one `tools/gpu_job.sh final` run from gpurun_out/job/ into profiles/ (tracked): the ncu
summaries, the launch list, the clocks record, the bench lines (ours + reference arm), the traffic file, and
profiles/r02/parity_r02.json assembled from the bench line's parity records."""
import json
import shutil
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
src = ROOT / "gpurun_out" / "job"
dst = ROOT / "profiles" / "r02"
dst.mkdir(parents=True, exist_ok=True)
for f in sorted(src.glob("prof_*.md")):
    shutil.copy(f, dst / f.name)
for name in ("launches_bench.csv", "clocks.csv"):
    if (src / name).exists():
        shutil.copy(src / name, dst / name)
shutil.copy(src / "roofline_traffic.json", ROOT / "profiles" / "roofline_traffic.json")
line = json.loads((src / "bench.json").read_text().strip().splitlines()[-1])
(dst / "bench_1gpu.json").write_text(json.dumps(line, indent=1) + "\n")
# the captures belong to the iteration counts of this run: bench.py withholds the derived issue / fp32 fractions when a later run
# of the same library solves at other counts (solver parameters changed)
tpath = ROOT / "profiles" / "roofline_traffic.json"
tj = json.loads(tpath.read_text())
iters = {"metric": line["solver"]["mean_iterations"], **{c["name"]: c["iterations_mean"] for c in line["configs"] if c.get("iterations_mean") is not None}}
for key, cap in tj["captures"].items():
    name = key.split("@")[0]
    if name in iters and (("@" not in key) or cap.get("frames_per_launch") == next((c.get("streams_per_gpu", 0) * c.get("steps", 0) for c in line["configs"] if c["name"] == name), None)):
        cap["iterations_mean_at_capture"] = round(float(iters[name]), 4)
tpath.write_text(json.dumps(tj, indent=1) + "\n")
ref = json.loads((src / "bench_reference.json").read_text().strip().splitlines()[-1])
(dst / "bench_reference_arm.json").write_text(json.dumps(ref, indent=1) + "\n")
par = {"build_id": line["solver"]["build_id"], "tolerance_rad": 1e-4,
       "oracle": "oracle mode B (float64 minimiser of the reference objective, KKT-polished): tests/golden/bench_parity.npz",
       "records": list(line.get("parity", []))}
for c in line["configs"]:
    p = c.get("parity")
    if isinstance(p, dict):
        par["records"].append({"bench_record": c["name"], **p})
    elif isinstance(p, list):
        par["records"] += [{"bench_record": c["name"], **q} for q in p]
(dst / "parity_r02.json").write_text(json.dumps(par, indent=1) + "\n")
print("collected into", dst, "build", par["build_id"])
for r in par["records"]:
    if "n" in r:
        print(f"  {r.get('bench_record', 'headline'):32s} {r['config']:28s} n {r['n']:5d} median {r['median']:.1e} p99 {r['p99']:.1e} max {r['max']:.1e} same_basin {r['same_basin']:.4f} flagged {r.get('flagged')}")
