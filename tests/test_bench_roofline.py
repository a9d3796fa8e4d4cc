"""bench.py's roofline bookkeeping (CPU): captures are tied to the library build AND to the iteration count they were taken at."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

PEAKS = dict(peak=6575.8, peak_src="test", peak_issue=4 * 148 * 1.965e9, peak_fp32=2 * 128 * 148 * 1.965e9)


def _file(tmp_path, build="abc"):
    p = tmp_path / "traffic.json"
    p.write_text(json.dumps({"build_id": build, "captures": {
        "metric": {"build_id": build, "frames_per_launch": 65536, "dram_bytes_per_launch": 20_000_000, "warp_inst_per_launch": 1.5e8,
                   "fp32_flop_per_launch": 1.7e9, "iterations_mean_at_capture": 2.43, "source": "x"},
        "leap_dexpilot_streams@256": {"build_id": build, "frames_per_launch": 76800, "dram_bytes_per_launch": 1, "warp_inst_per_launch": 1.0}}}))
    return p


def test_capture_of_this_build_and_iteration_count_is_reported(tmp_path):
    caps, note = bench.load_captures("abc", _file(tmp_path))
    assert note is None and set(caps) == {("metric", 65536), ("leap_dexpilot_streams", 76800)}
    r = bench.roofline_record(caps[("metric", 65536)], note, 65536, 0.232, 380, 2.44, **PEAKS)
    assert abs(r["achieved"] - 380 * 65536 / 0.232e-3 / 1e9) < 1e-9 and r["traffic"] == 20_000_000
    assert abs(r["issue"]["frac"] - 1.5e8 / 0.232e-3 / PEAKS["peak_issue"]) < 1e-12 and "fp32" in r and "note" not in r


def test_other_iteration_count_withholds_the_derived_fractions(tmp_path):
    caps, note = bench.load_captures("abc", _file(tmp_path))
    r = bench.roofline_record(caps[("metric", 65536)], note, 65536, 0.2, 380, 2.0, **PEAKS)
    assert "issue" not in r and "fp32" not in r and r["traffic"] == 20_000_000 and "withheld" in r["note"]
    # a capture without a recorded iteration count, or a record without one, is reported as before
    r = bench.roofline_record(caps[("leap_dexpilot_streams", 76800)], note, 76800, 10.0, 316, 4.2, **PEAKS)
    assert "issue" in r and "note" not in r


def test_other_build_withholds_everything(tmp_path):
    caps, note = bench.load_captures("another-build", _file(tmp_path))
    assert caps == {} and "another library build" in note
    r = bench.roofline_record(None, note, 65536, 0.232, 380, 2.43, **PEAKS)
    assert r["traffic"] is None and "issue" not in r and r["note"] == note


def test_committed_traffic_file_matches_the_committed_bench_record():
    t = json.loads((ROOT / "profiles" / "roofline_traffic.json").read_text())
    b = json.loads((ROOT / "profiles" / "r02" / "bench_1gpu.json").read_text())
    assert t["build_id"] == b["solver"]["build_id"] == b["roofline"]["capture"]["build_id"]
    assert all(c["build_id"] == t["build_id"] for c in t["captures"].values())


def test_committed_captures_belong_to_the_committed_library_sources():
    """The ncu captures under profiles/ are keyed by the sha256 stamp of csrc/dexr.cu + csrc/dexr_kernels.cuh + include/dexr.h: a
    source change without a new record run is caught here instead of by a `withheld` note in the next bench line."""
    from dex_retargeting_b200.build import source_id

    t = json.loads((ROOT / "profiles" / "roofline_traffic.json").read_text())
    assert t["build_id"] == source_id(), "library sources changed since the record run: re-run `tools/gpu_job.sh final` + tools/collect_profiles.py"
