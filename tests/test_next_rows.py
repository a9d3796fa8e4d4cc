"""Rows SURVEY.md section 8(f) marks "next": batched analytic warm start and the output adapters (host logic, no GPU)."""
import numpy as np
import pytest

from helpers import build_product
from dex_retargeting_b200.adapters import joint_order_map, load_trajectory, remap_qpos, save_trajectory
from dex_retargeting_b200.constants import HandType

torch = pytest.importorskip("torch")


@pytest.mark.parametrize("mano", [False, True])
def test_warm_start_batch_matches_single(mano):
    seq = build_product("offline/shadow_hand_right")
    rng = np.random.RandomState(4)
    S = 9
    quat = rng.randn(S, 4)
    pos = rng.randn(S, 3) * 0.4
    want = []
    for s in range(S):
        one = build_product("offline/shadow_hand_right")
        one.warm_start(pos[s], quat[s], HandType.right, is_mano_convention=mano)
        want.append(one.last_qpos.copy())
    last = torch.from_numpy(np.tile(seq.last_qpos, (S, 1)))
    seq.warm_start_batch(last, torch.from_numpy(pos), torch.from_numpy(quat), HandType.right, is_mano_convention=mano)
    np.testing.assert_allclose(last.numpy(), np.array(want), atol=2e-6)
    assert seq.is_warm_started
    with pytest.raises(ValueError):
        build_product("teleop/allegro_hand_right").warm_start_batch(torch.zeros(2, 16), torch.zeros(2, 3), torch.zeros(2, 4))


def test_joint_order_remap_numpy_and_torch():
    seq = build_product("teleop/allegro_hand_right")
    names = seq.joint_names
    sim_order = sorted(names)  # e.g. a simulator that sorts joints by name
    idx = joint_order_map(names, sim_order)
    q = np.arange(3 * len(names), dtype=np.float64).reshape(3, len(names))
    out = remap_qpos(q, idx)
    for j, n in enumerate(sim_order):
        np.testing.assert_array_equal(out[:, j], q[:, names.index(n)])
    t = remap_qpos(torch.from_numpy(q), idx)
    np.testing.assert_array_equal(t.numpy(), out)
    with pytest.raises(ValueError):
        joint_order_map(names, ["not_a_joint"])


def test_trajectory_pickle_roundtrip(tmp_path):
    seq = build_product("teleop/leap_hand_right")
    q = np.random.RandomState(0).randn(12, seq.optimizer.robot.dof)
    p = save_trajectory(tmp_path / "out" / "traj.pkl", q, seq.joint_names, config_path="teleop/leap_hand_right.yml")
    data, meta = load_trajectory(p)
    np.testing.assert_array_equal(data, q)
    assert meta["dof"] == 16 and meta["joint_names"] == seq.joint_names and meta["config_path"].endswith("leap_hand_right.yml")
    with pytest.raises(ValueError):
        save_trajectory(tmp_path / "bad.pkl", q[:, :3], seq.joint_names)


def test_stream_state_checkpoint_round_trip():
    """StreamState.state_dict / from_state_dict: host copies of every tensor, None kept, shapes checked, and a checkpoint from
    before the carried damping existed resumes at the solver's default (zeros)."""
    import torch

    from dex_retargeting_b200.seq_retarget import StreamState

    S = 5
    st = StreamState(last_qpos=torch.randn(S, 16), filter_state=torch.randn(S, 16), filter_init=torch.ones(S, dtype=torch.uint8),
                     projected=torch.zeros(S, 6, dtype=torch.uint8), damping=torch.full((S,), 0.3))
    sd = st.state_dict()
    assert set(sd) == {"last_qpos", "filter_state", "filter_init", "projected", "damping"}
    sd["last_qpos"][0, 0] = 99.0  # a copy, not a view
    assert float(st.last_qpos[0, 0]) != 99.0
    back = StreamState.from_state_dict(st.state_dict())
    for k in sd:
        assert torch.equal(getattr(back, k), getattr(st, k)) and getattr(back, k).is_contiguous()
    old = {k: v for k, v in st.state_dict().items() if k != "damping"}
    old["projected"] = None
    back = StreamState.from_state_dict(old)
    assert back.projected is None and torch.equal(back.damping, torch.zeros(S))
    with pytest.raises(ValueError, match="lacks"):
        StreamState.from_state_dict({"last_qpos": st.last_qpos})
    with pytest.raises(ValueError, match="streams"):
        StreamState.from_state_dict({**st.state_dict(), "filter_init": torch.ones(S + 1, dtype=torch.uint8)})
