"""`SeqRetargeting.warm_start` / `warm_start_batch` against tests/golden/reference_warm_start.npz: the reference's own
seq_retarget.py:45-110 executed on seeded wrist poses (tests/tools/gen_reference_warm_start.py; pytransform3d's quaternion and
intrinsic-xyz Euler conventions served by scipy's Rotation, pinocchio by the oracle's FK, itself pinned to the reference's tree
FK).  Row f2 of SURVEY.md section 8: the analytic initialisation of the six dummy free joints, single and batched."""
import numpy as np
import pytest

from helpers import GOLDEN, build_product
from dex_retargeting_b200.constants import HandType

torch = pytest.importorskip("torch")
VEC = np.load(GOLDEN / "reference_warm_start.npz")
KEYS = [str(k) for k in VEC["keys"]]


def test_inventory():
    assert len(KEYS) >= 13 and "offline/shadow_hand_right" in KEYS and "offline/panda_gripper" in KEYS


@pytest.mark.parametrize("key", KEYS)
@pytest.mark.parametrize("hand", ["right", "left"])
@pytest.mark.parametrize("mano", [False, True])
def test_single_warm_start_matches_reference(key, hand, mano):
    pos, quat, want = VEC[f"{key}/pos"], VEC[f"{key}/quat"], VEC[f"{key}/{hand}/{int(mano)}"]
    for s in range(len(pos)):
        seq = build_product(key)
        seq.warm_start(pos[s], quat[s], HandType[hand], is_mano_convention=mano)
        np.testing.assert_allclose(seq.last_qpos, want[s], atol=2e-6)  # last_qpos is float32 on both sides
        assert seq.is_warm_started


def batched(key, hand, mano, device):
    pos, quat = VEC[f"{key}/pos"], VEC[f"{key}/quat"]
    seq = build_product(key)
    S = len(pos)
    last = torch.from_numpy(np.tile(seq.last_qpos, (S, 1))).to(device)
    seq.warm_start_batch(last, torch.from_numpy(pos).to(device), torch.from_numpy(quat).to(device), HandType[hand],
                         is_mano_convention=mano)
    return last.cpu().numpy()


@pytest.mark.parametrize("key", KEYS)
def test_batched_warm_start_matches_reference_cpu(key):
    for hand in ("right", "left"):
        for mano in (False, True):
            np.testing.assert_allclose(batched(key, hand, mano, "cpu"), VEC[f"{key}/{hand}/{int(mano)}"], atol=2e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("key", KEYS)
def test_batched_warm_start_matches_reference_gpu(key):
    """The same tensor algebra on CUDA tensors (StreamState.last_qpos lives on the device)."""
    for hand in ("right", "left"):
        for mano in (False, True):
            np.testing.assert_allclose(batched(key, hand, mano, "cuda:0"), VEC[f"{key}/{hand}/{int(mano)}"], atol=2e-6)
