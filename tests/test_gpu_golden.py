"""GPU vs the committed golden vectors (no CPU solve at test time): |dq|_inf < 1e-4 rad against the oracle's
converged minimiser, through the C ABI, for every loss / robot family in tests/golden/oracle_vectors.npz."""
import numpy as np
import pytest

from helpers import GOLDEN, build_product

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

VEC = np.load(GOLDEN / "oracle_vectors.npz")
CASES = sorted({k.split("/")[0] for k in VEC.files if k.endswith("/qpos_converged")})
TOL = 1e-4


@pytest.mark.parametrize("case", CASES)
def test_frames_against_golden(case):
    dev = torch.device("cuda", 0)
    seq = build_product(str(VEC[f"{case}/key"]))
    opt = seq.optimizer
    refs, fixed, x0 = VEC[f"{case}/ref_value"], VEC[f"{case}/fixed_qpos"], VEC[f"{case}/last_qpos"]
    B = refs.shape[0]
    status = torch.zeros(B, dtype=torch.int32, device=dev)
    cost = torch.zeros(B, dtype=torch.float32, device=dev)
    proj = torch.zeros((B, opt._objective_spec().len_proj), dtype=torch.uint8, device=dev) if opt.retargeting_type == "DEXPILOT" else None
    q = opt.retarget_batch(torch.from_numpy(refs).to(dev), torch.from_numpy(fixed).to(dev) if fixed.shape[1] else None,
                           torch.from_numpy(x0).to(dev), status_out=status, cost_out=cost, projected=proj)
    torch.cuda.synchronize()
    q, cost = q.cpu().numpy(), cost.cpu().numpy()
    assert int((status.cpu().numpy() >> 24).max()) == 0
    dq = np.abs(q - VEC[f"{case}/qpos_converged"]).max(1)
    same = dq < TOL
    assert same.mean() >= 0.9, f"{case}: {same.mean():.2f} within {TOL}, worst {dq.max():.2e}"
    assert np.median(dq) < 1e-5
    fb = VEC[f"{case}/cost_converged"]
    np.testing.assert_allclose(cost[same], fb[same], rtol=3e-5, atol=1e-8)


def test_stream_against_golden():
    dev = torch.device("cuda", 0)
    seq = build_product("teleop/allegro_hand_right")
    kp = torch.from_numpy(VEC["allegro_stream/keypoints"][None].astype(np.float32)).to(dev)
    out, _ = seq.retarget_sequences(kp.contiguous())
    torch.cuda.synchronize()
    err = np.abs(out.cpu().numpy()[0] - VEC["allegro_stream/robot_qpos"]).max(1)
    assert (err < TOL).mean() >= 0.95 and np.median(err) < 1e-5
