#!/usr/bin/env python
"""Generate the data fixtures under tests/golden/ from the reference checkout (build container only).

`/root/reference` does not exist on the GPU box, so everything the tests, smoke() and bench.py
need from it is derived here ONCE and committed:

  tests/golden/robots/<urdf stem>.json   joint tree of every hand URDF the reference configs use
                                          (links, joints: type/parent/child/xyz/rpy/axis/limit/mimic;
                                          meshes, inertias, visuals dropped) -- written by
                                          dex_retargeting_b200.urdf.KinematicModel.to_dict()
  tests/golden/configs.json               every retargeting YAML of the reference, parsed, keyed by
                                          "<teleop|offline>/<file stem>"
  tests/golden/human_joint_right.npy      example/profiling/human_joint_right.pkl as float32 [621,21,3]

Usage: python tools/gen_fixtures.py [/root/reference]
"""
import json
import pickle
import sys
from pathlib import Path

import numpy as np
import yaml

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from dex_retargeting_b200.urdf import KinematicModel  # noqa: E402


def main():
    ref = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    out = ROOT / "tests" / "golden"
    (out / "robots").mkdir(parents=True, exist_ok=True)

    configs = {}
    urdfs = set()
    for yml in sorted((ref / "src/dex_retargeting/configs").glob("*/*.yml")):
        with yml.open() as f:
            cfg = yaml.safe_load(f)["retargeting"]
        configs[f"{yml.parent.name}/{yml.stem}"] = cfg
        urdfs.add(cfg["urdf_path"].strip())
    with (out / "configs.json").open("w") as f:
        json.dump(configs, f, indent=1, sort_keys=True)

    for rel in sorted(urdfs):
        model = KinematicModel.from_urdf(ref / "assets/robots/hands" / rel)
        with (out / "robots" / (Path(rel).stem + ".json")).open("w") as f:
            json.dump(model.to_dict(), f, separators=(",", ":"))
        print(f"{rel}: dof={model.dof} links={len(model.link_names)}")

    with (ref / "example/profiling/human_joint_right.pkl").open("rb") as f:
        traj = np.asarray(pickle.load(f), dtype=np.float32)
    assert traj.shape == (621, 21, 3)
    np.save(out / "human_joint_right.npy", traj)
    print("configs:", len(configs), "robots:", len(urdfs), "trajectory:", traj.shape)


if __name__ == "__main__":
    main()
