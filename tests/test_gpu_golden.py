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
    # the kernel reports F in fp32: link positions of ~1 m carry ~6e-8 m of rounding, a 5 mm residual therefore ~1e-5 relative
    # and F (quadratic in the residuals) ~2e-5 per term -- measured spread on the dummy-base hands +-2.6e-5, outliers 7e-5
    np.testing.assert_allclose(cost[same], fb[same], rtol=1.5e-4, atol=1e-8)


def test_stream_against_golden():
    dev = torch.device("cuda", 0)
    seq = build_product("teleop/allegro_hand_right")
    kp = torch.from_numpy(VEC["allegro_stream/keypoints"][None].astype(np.float32)).to(dev)
    out, _ = seq.retarget_sequences(kp.contiguous())
    torch.cuda.synchronize()
    err = np.abs(out.cpu().numpy()[0] - VEC["allegro_stream/robot_qpos"]).max(1)
    assert (err < TOL).mean() >= 0.95 and np.median(err) < 1e-5


# ---------------------------------------------------------------------------------------------------------------
# tests/golden/reference_vectors.npz: numbers computed by the reference's OWN optimizer.py / seq_retarget.py code
# (tests/tools/gen_reference_vectors.py; pinocchio and nlopt shimmed).  The reference stops SLSQP early (ftol_abs
# 1e-5 / 1e-6), so its joint vector is not a fixed point anything else can reproduce to 1e-4 rad; what can be -- and
# is -- asserted is that on the same inputs the CUDA path ends at or below the reference's own objective value, close
# to its joint vector, and with bit-identical DexPilot hysteresis flags.
RVEC = np.load(GOLDEN / "reference_vectors.npz")
RCASES = sorted({k.split("/")[0] for k in RVEC.files if k.endswith("/retarget_cost")})
RSTREAMS = sorted({k.split("/")[0] for k in RVEC.files if k.startswith("stream_") and k.endswith("/robot_qpos")})


@pytest.mark.parametrize("case", RCASES)
def test_frames_never_worse_than_reference_retarget(case):
    dev = torch.device("cuda", 0)
    seq = build_product(str(RVEC[f"{case}/key"]))
    opt = seq.optimizer
    refs, fixed, x0 = RVEC[f"{case}/ref_value"], RVEC[f"{case}/fixed_qpos"], RVEC[f"{case}/last_qpos"]
    B = refs.shape[0]
    status = torch.zeros(B, dtype=torch.int32, device=dev)
    cost = torch.zeros(B, dtype=torch.float32, device=dev)
    dexpilot = opt.retargeting_type == "DEXPILOT"
    proj = torch.zeros((B, opt._objective_spec().len_proj), dtype=torch.uint8, device=dev) if dexpilot else None
    q = opt.retarget_batch(torch.from_numpy(refs).to(dev), torch.from_numpy(fixed).to(dev) if fixed.shape[1] else None,
                           torch.from_numpy(x0).to(dev), status_out=status, cost_out=cost, projected=proj)
    torch.cuda.synchronize()
    q, cost = q.cpu().numpy(), cost.cpu().numpy().astype(np.float64)
    assert int((status.cpu().numpy() >> 25).max()) == 0  # finite everywhere
    ref_cost = RVEC[f"{case}/retarget_cost"]
    # consistent objective at the GPU's answer (fp32) <= at the reference's early-stopped answer (reference closure, f64)
    # The problem is non-convex: from the same start two correct solvers can settle in different local minima.  It shows
    # in the DexPilot cases, where every other problem carries a synthetic 200x-weighted pinch target far from the warm
    # start (measured: ability 4 of 12 end above and 1 far below the reference's objective, leap 1 of 12; none in the other
    # five families).  Count those; everywhere else the CUDA path must end at or below the reference's own objective.
    worse = cost > ref_cost * (1 + 2e-5) + 1e-7
    allowed = B // 3 if dexpilot else 0
    assert worse.sum() <= allowed, f"{case}: {worse.sum()}/{B} end above the reference's objective, worst excess {(cost - ref_cost).max():.3e}"
    assert np.median(cost - ref_cost) <= 1e-7
    if dexpilot:  # the flags depend on the inputs only (optimizer.py:466-476): must be identical
        np.testing.assert_array_equal(proj.cpu().numpy().astype(bool), RVEC[f"{case}/projected"])
    dq = np.abs(q - RVEC[f"{case}/retarget"]).max(1)
    assert np.median(dq) < 0.15  # same basin as the early-stopped SLSQP iterate; tight parity is vs the converged minimiser


@pytest.mark.parametrize("stream", RSTREAMS)
def test_sequences_track_reference_stream(stream):
    dev = torch.device("cuda", 0)
    seq = build_product(str(RVEC[f"{stream}/key"]))
    kp = torch.from_numpy(RVEC[f"{stream}/keypoints"][None].astype(np.float32)).to(dev)
    out, state = seq.retarget_sequences(kp.contiguous())
    torch.cuda.synchronize()
    want, got = RVEC[f"{stream}/robot_qpos"], out.cpu().numpy()[0].astype(np.float64)
    # The early-stopped SLSQP iterate leaves the weakly determined joint directions wherever they were (up to ~1 rad on
    # the 24-DoF Shadow hand), so the streams are compared where the objective lives: the task vectors of the two filtered
    # joint trajectories, through the host float64 FK.  Bar: the reference's own 1e-2 m (tests/test_optimizer.py:141, a MEAN
    # over problems and vectors) on the mean, 5e-3 on the median, and 2e-2 on the single worst vector of the worst frame
    # (measured: 1.57e-2 on the LEAP DexPilot stream, 1.1e-2 or less elsewhere -- a DexPilot frame whose pinch flags just
    # switched has minimisers that differ by more than a centimetre in one finger-pair vector at the same loss).
    opt, robot = seq.optimizer, seq.optimizer.robot
    links = list(opt.computed_link_indices)
    o_sel, t_sel = np.asarray(opt.origin_link_indices), np.asarray(opt.task_link_indices)

    def task_vectors(q):
        robot.compute_forward_kinematics(q)
        pos = np.stack([robot.get_link_pose(i)[:3, 3] for i in links])
        return pos[t_sel] - pos[o_sel]

    err = np.array([np.linalg.norm(task_vectors(got[t]) - task_vectors(want[t]), axis=1).max() for t in range(want.shape[0])])
    assert np.mean(err) < 1e-2 and np.median(err) < 5e-3 and err.max() < 2e-2, (np.mean(err), np.median(err), err.max())
    if seq.optimizer.retargeting_type == "DEXPILOT":
        np.testing.assert_array_equal(state.projected.cpu().numpy()[0].astype(bool), RVEC[f"{stream}/projected"][-1])
