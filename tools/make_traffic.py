#!/usr/bin/env python
"""Assemble profiles/roofline_traffic.json from `ncu --set full` captures of the bench kernels (one per configuration record
of bench.py), stamped with the build id of the library they were taken from.

  python tools/make_traffic.py <dir with prof_*.ncu-rep> <out.json> <build_id>
"""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
# capture file stem -> (bench record name, frames per launch, algorithmic bytes per frame)
CAPTURES = {
    "prof_frames_allegro": ("metric", 65536, 380),
    "prof_frames_shadowpos": ("shadow_position_narrowed", 65536, 492),
    "prof_frames_leapdp": ("leap_dexpilot_frames", 65536, 380),
    "prof_streams_256x300": ("leap_dexpilot_streams@256", 256 * 300, 316),     # one GPU's shard of the 8-GPU job
    "prof_streams_2048x300": ("leap_dexpilot_streams@2048", 2048 * 300, 316),  # the whole job on one GPU
}


def main():
    src, out, build_id = Path(sys.argv[1]), Path(sys.argv[2]), sys.argv[3]
    caps = {}
    for stem, (name, frames, bpf) in CAPTURES.items():
        rep = src / f"{stem}.ncu-rep"
        if not rep.exists():
            continue
        md, tj = src / f"{stem}.md", src / f"{stem}.traffic.json"
        subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_summary.py"), str(rep), str(md), f"{name}: {stem}", "--traffic", str(tj),
                        "--frames", str(frames), "--bytes-per-frame", str(bpf), "--build-id", build_id], check=True, capture_output=True)
        caps[name] = json.loads(tj.read_text())
        caps[name]["source"] = f"profiles/r02/{stem}.md (ncu --set full --clock-control none, one captured launch)"
    out.write_text(json.dumps({"build_id": build_id, "captures": caps}, indent=1) + "\n")
    print("wrote", out, "with", sorted(caps))


if __name__ == "__main__":
    main()
