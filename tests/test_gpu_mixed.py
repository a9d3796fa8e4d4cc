"""Mixed-robot launch (`retarget_batch_mixed` -> `dexr_solve_frames_multi`): BASELINE.json config 5 -- six robots, one
persistent launch -- must give bit-identical results to one `retarget_batch` launch per robot (the reference would build one
optimizer per robot and run them one after the other, retargeting_config.py:167-257), on every solver instantiation
(16-lane block / dense, 32-lane arrow / dense), with ragged group sizes, DexPilot flags and fixed joints in the mix."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))
import workloads as W  # noqa: E402

pytestmark = pytest.mark.gpu


def _jobs(keys, sizes, seed0=400):
    import torch

    dev = torch.device("cuda", 0)
    jobs = []
    for i, (key, n) in enumerate(zip(keys, sizes)):
        seq = W.build(key, device=0)
        opt = seq.optimizer
        kp, x0, fixed, _ = W.frames(seq, n, seed0 + i)
        kw = dict(keypoints=torch.from_numpy(kp).to(dev), last_qpos=torch.from_numpy(x0).to(dev),
                  fixed_qpos=torch.from_numpy(fixed).to(dev) if fixed is not None else None,
                  status_out=torch.zeros((n,), dtype=torch.int32, device=dev))
        if opt.retargeting_type == "DEXPILOT":
            kw["projected"] = torch.zeros((n, opt._objective_spec().len_proj), dtype=torch.uint8, device=dev)
        jobs.append((opt, kw))
    return jobs


def _clone(kw):
    return {k: (v.clone() if hasattr(v, "clone") else v) for k, v in kw.items()}


@pytest.mark.parametrize("keys,sizes", [
    (W.MIXED_KEYS, [2048] * 6),                                                        # config 5 in small
    (W.MIXED_KEYS, [777, 3, 1030, 64, 129, 2000]),                                      # ragged, smaller than one tile per CTA
    ([W.LEAP_DEXPILOT_KEY, W.SHADOW_POS_KEY, "offline/schunk_svh_hand_right", W.METRIC_KEY, "teleop/panda_gripper"],
     [500, 300, 200, 4096, 50]),                                                        # dexpilot flags, free-flying bases, prismatic
    ([W.METRIC_KEY], [5000]),                                                           # a single group
    (W.MIXED_KEYS, [9000] * 6),                                                         # enough work per SM: one persistent CTA per SM and group
], ids=["six-robots", "ragged", "loss-families", "single", "six-robots-large"])
@pytest.mark.parametrize("mode", ["streams", "streams-spread", "persistent"])
def test_mixed_launch_equals_per_robot_launches(keys, sizes, mode, monkeypatch):
    """Every implementation behind dexr_solve_frames_multi: fork-join launches on side streams with the CTAs of all groups
    sized together (default: one-round tiles of exact size, shrunk tiles, whole-round CTAs or persistent CTAs depending on
    the total work, slowest solver first), the same with every group sized as a lone launch (DEXR_MULTI_SLOTS=spread), and
    the single persistent kernel whose CTAs walk the groups (DEXR_MULTI_MODE=persistent); all switches are read per call."""
    import torch

    monkeypatch.setenv("DEXR_MULTI_MODE", "persistent" if mode == "persistent" else "streams")
    if mode == "streams-spread":
        monkeypatch.setenv("DEXR_MULTI_SLOTS", "spread")

    from dex_retargeting_b200.optimizer import retarget_batch_mixed

    jobs = _jobs(keys, sizes)
    ref = []
    for opt, kw in jobs:
        k2 = _clone(kw)
        q = opt.retarget_batch(**k2)
        ref.append((q, k2))
    torch.cuda.synchronize()
    mixed = [(opt, _clone(kw)) for opt, kw in jobs]
    outs = retarget_batch_mixed(mixed)
    torch.cuda.synchronize()
    for (q_ref, k_ref), q_mix, (_, k_mix) in zip(ref, outs, mixed):
        assert torch.equal(q_ref, q_mix)
        assert torch.equal(k_ref["status_out"], k_mix["status_out"])
        if "projected" in k_ref:
            assert torch.equal(k_ref["projected"], k_mix["projected"])
        assert int((k_mix["status_out"] >> 24).max()) == 0


def test_mixed_launch_argument_errors():
    from dex_retargeting_b200.optimizer import retarget_batch_mixed

    jobs = _jobs([W.METRIC_KEY], [16])
    assert retarget_batch_mixed([]) == []
    with pytest.raises(ValueError, match="at most"):
        retarget_batch_mixed(jobs * 17)
    opt, kw = jobs[0]
    bad = dict(kw)
    bad["last_qpos"] = kw["last_qpos"][:, :5].contiguous()
    with pytest.raises(ValueError):
        retarget_batch_mixed([(opt, bad)])
