"""Rows f2 / f4 of SURVEY.md section 8 on the GPU: the output adapters applied to what the kernels produce -- joint-order
remapping of a whole device batch by name (reference README.md:84-106, show_realtime_retargeting.py:113-118) and the
reference's pickle trajectory layout (detect_from_video.py:60-69) written from a device stream -- and a stream started from the
batched analytic warm start (seq_retarget.py:45-110)."""
import pickle

import numpy as np
import pytest

from helpers import build_product, keypoint_trajectory
from dex_retargeting_b200.adapters import joint_order_map, load_trajectory, remap_qpos, save_trajectory
from dex_retargeting_b200.constants import HandType

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def test_device_batch_remap_and_trajectory_file(tmp_path):
    seq = build_product("teleop/schunk_svh_hand_right", device=0)  # 20 DoF, 11 mimic joints: full qpos != optimised qpos
    dev = torch.device("cuda", 0)
    kp = torch.from_numpy(keypoint_trajectory()[None, :120].astype(np.float32)).to(dev)
    rq, _ = seq.retarget_sequences(kp.contiguous())                      # [1, 120, 20] pinocchio order, on the device
    names = seq.joint_names
    sim_order = sorted(names, reverse=True)                              # a simulator's own joint order
    idx = joint_order_map(names, sim_order)
    on_dev = remap_qpos(rq, idx)                                         # device gather, no host round trip
    assert on_dev.is_cuda and on_dev.shape == rq.shape
    torch.cuda.synchronize()
    host = rq.cpu().numpy()
    np.testing.assert_array_equal(on_dev.cpu().numpy(), host[..., idx])
    for j, n in enumerate(sim_order):
        np.testing.assert_array_equal(on_dev[0, :, j].cpu().numpy(), host[0, :, names.index(n)])
    path = save_trajectory(tmp_path / "traj.pkl", rq[0], names, config_path="teleop/schunk_svh_hand_right.yml")
    with open(path, "rb") as f:
        raw = pickle.load(f)
    assert set(raw) == {"data", "meta_data"} and set(raw["meta_data"]) == {"config_path", "dof", "joint_names"}
    data, meta = load_trajectory(path)
    np.testing.assert_array_equal(data, host[0])
    assert meta["dof"] == 20 and meta["joint_names"] == names


def test_streams_started_from_batched_warm_start():
    """Position retargeting with a free-flying base: the six dummy joints of S streams are initialised analytically on the
    device (warm_start_batch on StreamState.last_qpos), then the streams run; a stream started that way must reach the same
    first-frame solution as the single-stream API started with warm_start()."""
    key = "offline/allegro_hand_right"
    seq = build_product(key, device=0)
    dev = torch.device("cuda", 0)
    S = 4
    rng = np.random.RandomState(1)
    kp = keypoint_trajectory()[:20].astype(np.float32)
    pos = np.tile(kp[0, 0], (S, 1)) + rng.randn(S, 3) * 0.0
    quat = np.tile(np.array([1.0, 0, 0, 0]), (S, 1))
    state = seq.make_stream_state(S)
    seq.warm_start_batch(state.last_qpos, torch.from_numpy(pos).to(dev), torch.from_numpy(quat).to(dev), HandType.right)
    tk = torch.from_numpy(np.tile(kp[None], (S, 1, 1, 1))).to(dev).contiguous()
    rq, state = seq.retarget_sequences(tk, state=state)
    torch.cuda.synchronize()
    one = build_product(key, device=0)
    one.warm_start(pos[0], quat[0], HandType.right)
    idx = np.asarray(one.optimizer.target_link_human_indices)
    first = one.retarget(kp[0][idx])
    np.testing.assert_allclose(rq[0, 0].cpu().numpy(), first, atol=2e-4)
    assert torch.equal(rq[0], rq[1])  # identical streams, identical results
