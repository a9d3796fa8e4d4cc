#!/usr/bin/env python
"""Throughput of the BASELINE.json configurations 2-5 on ONE GPU (device-resident inputs, CUDA-event timing).

  2. VectorOptimizer Allegro 16-DoF, batch 4096 synthetic 21-kpt frames
  3. PositionOptimizer Shadow (24 + 6 dummy = 30 DoF), batch 65536
  4. DexPilotOptimizer LEAP 16-DoF, streaming sequences x 300 frames, temporal state carried in-kernel
     (2048 streams = the whole 8-GPU job on one GPU, and 256 streams = one GPU's shard of it)
  5. mixed robots: {Allegro, Shadow, LEAP, Ability, SVH, Inspire} x 16384 frames each, one launch per robot on
     its own CUDA stream

Prints one JSON object per configuration; writes a markdown table when given --out.
"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from bench import make_batch  # noqa: E402
from helpers import build_product, keypoint_trajectory  # noqa: E402


def spec_of(seq):
    opt = seq.optimizer
    hi = np.asarray(opt.target_link_human_indices)
    if opt.retargeting_type == "POSITION":
        return list(opt.body_names), [int(v) for v in hi.reshape(-1)], 1.0, False
    names = list(opt.origin_link_names[:1]) + list(opt.task_link_names)
    human = [int(hi[0, 0])] + [int(v) for v in hi[1]]
    return names, human, float(opt.scaling), True


def synth(seq, n, seed):
    names, human, scale, centre = spec_of(seq)
    kin = seq.optimizer.robot.kin
    lim = kin.joint_limits.copy()
    narrowed = False
    for i, nm in enumerate(kin.dof_joint_names):  # dummy free joints: +-0.5 m / +-pi instead of +-5 m / +-2 pi
        if "dummy" in nm:
            lim[i] = [-0.5, 0.5] if "translation" in nm else [-np.pi, np.pi]
            narrowed = True
    saved = kin.joint_limits
    kin.joint_limits = lim
    try:
        kp, init = make_batch(kin, (names, human, scale), n, seed, centre=centre)
    finally:
        kin.joint_limits = saved
    opt = seq.optimizer
    x0 = init[:, opt.idx_pin2target]
    fixed = init[:, opt.idx_pin2fixed] if len(opt.idx_pin2fixed) else None
    return kp, np.ascontiguousarray(x0), fixed, narrowed


def time_launches(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def frames_config(key, B, dev, label, reps=10):
    seq = build_product(key)
    opt = seq.optimizer
    sets = []
    for s in range(4):
        kp, x0, fixed, narrowed = synth(seq, B, 100 + s)
        sets.append((torch.from_numpy(kp).to(dev), torch.from_numpy(x0).to(dev),
                     torch.from_numpy(np.ascontiguousarray(fixed)).to(dev) if fixed is not None else None))
    out = torch.empty((B, opt.opt_dof), dtype=torch.float32, device=dev)
    status = torch.zeros((B,), dtype=torch.int32, device=dev)
    proj = torch.zeros((B, opt._objective_spec().len_proj), dtype=torch.uint8, device=dev) if opt.retargeting_type == "DEXPILOT" else None
    it = [0]

    def run():
        k, x, f = sets[it[0] % len(sets)]
        it[0] += 1
        opt.retarget_batch(keypoints=k, fixed_qpos=f, last_qpos=x, out=out, status_out=status, projected=proj)

    ms = time_launches(run, reps)
    st = status.cpu().numpy()
    n, dof = opt.opt_dof, opt.robot.dof
    bytes_per_frame = 252 + 4 * n + 4 * n
    return dict(config=label, key=key, type=opt.retargeting_type, n_var=n, dof=dof, batch=B, ms_per_launch=ms,
                frames_per_s=B / ms * 1e3, mean_iterations=float((st & 0xffff).mean()), flagged=int(((st >> 24) != 0).sum()),
                bytes_per_frame=bytes_per_frame, hbm_gbs=bytes_per_frame * B / ms / 1e6, launch=opt.engine().launch_info(),
                dummy_limits_narrowed=narrowed)


def sequences_config(key, S, T, dev, label, reps=3):
    seq = build_product(key)
    opt = seq.optimizer
    rng = np.random.RandomState(7)
    base = keypoint_trajectory()[:T].astype(np.float32)  # recorded right-hand trajectory, first T frames
    kp = base[None] + rng.randn(S, 1, 21, 3).astype(np.float32) * 0.002  # per-stream 2 mm offsets
    kp[:, :, 0] = 0
    tk = torch.from_numpy(np.ascontiguousarray(kp)).to(dev)
    out = torch.empty((S, T, opt.robot.dof), dtype=torch.float32, device=dev)
    status = torch.zeros((S, T), dtype=torch.int32, device=dev)

    def run():
        seq.retarget_sequences(tk, out=out, status_out=status)

    ms = time_launches(run, reps, warm=1)
    st = status.cpu().numpy()
    return dict(config=label, key=key, type=opt.retargeting_type, n_var=opt.opt_dof, dof=opt.robot.dof, streams=S, steps=T,
                ms_per_launch=ms, frames_per_s=S * T / ms * 1e3, mean_iterations=float((st & 0xffff).mean()),
                flagged=int(((st >> 24) != 0).sum()), bytes_per_frame=252 + 4 * opt.robot.dof,
                hbm_gbs=(252 + 4 * opt.robot.dof) * S * T / ms / 1e6, launch=opt.engine().launch_info(),
                us_per_frame_per_stream=ms * 1e3 / T)


def mixed_config(dev, per_robot=16384, reps=5):
    keys = ["teleop/allegro_hand_right", "teleop/shadow_hand_right", "teleop/leap_hand_right", "teleop/ability_hand_right",
            "teleop/schunk_svh_hand_right", "teleop/inspire_hand_right"]
    jobs = []
    for i, key in enumerate(keys):
        seq = build_product(key)
        kp, x0, fixed, _ = synth(seq, per_robot, 300 + i)
        jobs.append((seq.optimizer, torch.from_numpy(kp).to(dev), torch.from_numpy(x0).to(dev),
                     torch.empty((per_robot, seq.optimizer.opt_dof), dtype=torch.float32, device=dev), torch.cuda.Stream(dev)))
    main = torch.cuda.current_stream(dev)

    def run():
        for opt, k, x, o, s in jobs:
            s.wait_stream(main)
            opt.retarget_batch(keypoints=k, last_qpos=x, out=o, stream=s)
        for *_, s in jobs:
            main.wait_stream(s)

    ms = time_launches(run, reps)
    total = per_robot * len(keys)
    return dict(config="5 mixed robots x %d frames, 6 concurrent launches" % per_robot, keys=keys, batch=total, ms_per_launch=ms,
                frames_per_s=total / ms * 1e3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    rows = [
        frames_config("teleop/allegro_hand_right", 4096, dev, "2 Vector Allegro, batch 4096"),
        frames_config("teleop/allegro_hand_right", 65536, dev, "(metric) Vector Allegro, batch 65536"),
        frames_config("offline/shadow_hand_right", 65536, dev, "3 Position Shadow n=30, batch 65536"),
        frames_config("teleop/leap_hand_right_dexpilot", 65536, dev, "DexPilot LEAP independent frames, batch 65536"),
        sequences_config("teleop/leap_hand_right_dexpilot", 2048, 300, dev, "4 DexPilot LEAP 2048 streams x 300 (whole job on 1 GPU)"),
        sequences_config("teleop/leap_hand_right_dexpilot", 256, 300, dev, "4 DexPilot LEAP 256 streams x 300 (1/8 shard)"),
        mixed_config(dev),
    ]
    for r in rows:
        print(json.dumps(r), flush=True)
    if args.out:
        lines = ["# BASELINE.json configurations on one B200 (device-resident inputs, CUDA events)", "",
                 "| config | frames/s | ms/launch | mean LM iterations | flagged | algorithmic GB/s |", "|---|---|---|---|---|---|"]
        for r in rows:
            lines.append("| %s | %.3e | %.3f | %s | %s | %s |" % (
                r["config"], r["frames_per_s"], r["ms_per_launch"],
                ("%.2f" % r["mean_iterations"]) if "mean_iterations" in r else "-", r.get("flagged", "-"),
                ("%.1f" % r["hbm_gbs"]) if "hbm_gbs" in r else "-"))
        Path(args.out).write_text("\n".join(lines) + "\n\n```\n" + "\n".join(json.dumps(r) for r in rows) + "\n```\n")


if __name__ == "__main__":
    main()
