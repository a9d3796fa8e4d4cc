"""Reference-executed vectors (tests/golden/reference_vectors.npz, made by tests/tools/gen_reference_vectors.py by running
the reference's OWN optimizer.py / kinematics_adaptor.py / seq_retarget.py / optimizer_utils.py with pinocchio and nlopt
shimmed): the oracle's restatement of the objective, gradient, DexPilot state machine, mimic adaptor, solver driver and
sequence wrapper must reproduce what the reference's code computed on the same inputs.

What these do NOT pin: pinocchio's FK/Jacobian arithmetic (the shim robot uses the oracle's pure-Python FK) and nlopt's
SLSQP internals (scipy's SLSQP stands in on both sides)."""
import numpy as np
import pytest

from helpers import GOLDEN, build_oracle
from oracle.solvers import OracleSeqRetargeting, solve_reference

VEC = np.load(GOLDEN / "reference_vectors.npz")
CASES = sorted({k.split("/")[0] for k in VEC.files if k.endswith("/values")})
STREAMS = sorted({k.split("/")[0] for k in VEC.files if k.startswith("stream_") and k.endswith("/robot_qpos")})


def test_inventory():
    assert len(CASES) == 7 and len(STREAMS) == 4
    assert {"leap_dexpilot", "ability_dexpilot_mimic"} <= set(CASES)


@pytest.mark.parametrize("case", CASES)
def test_index_maps_match_reference(case):
    """optimizer.py:27-40 (idx_pin2target / idx_pin2fixed) and :66-75 (mimic joints leave the fixed set)."""
    o = build_oracle(str(VEC[f"{case}/key"]))
    np.testing.assert_array_equal(o.idx_pin2target, VEC[f"{case}/idx_pin2target"])
    np.testing.assert_array_equal(o.idx_pin2fixed, VEC[f"{case}/idx_pin2fixed"])


@pytest.mark.parametrize("case", CASES)
def test_objective_value_and_gradient_match_reference_closure(case):
    """The closure returned by get_objective_function (optimizer.py:138-200, :241-306, :456-577), called as nlopt calls it."""
    o = build_oracle(str(VEC[f"{case}/key"]))
    refs, fixed, x0 = VEC[f"{case}/ref_value"], VEC[f"{case}/fixed_qpos"], VEC[f"{case}/last_qpos"]
    pts, vals, grads = VEC[f"{case}/points"], VEC[f"{case}/values"], VEC[f"{case}/grads"]
    for i in range(refs.shape[0]):
        if o.type == "dexpilot":
            o.projected[:] = False
        obj = o.make_objective(refs[i], fixed[i], x0[i])
        for p in range(pts.shape[1]):
            v, g = obj.value_and_grad(pts[i, p])
            # the reference evaluates the loss in float64 with float32-rounded targets / weights; so does the oracle
            assert v == pytest.approx(float(vals[i, p]), rel=1e-9, abs=1e-12)
            np.testing.assert_allclose(g, grads[i, p], rtol=1e-8, atol=1e-11)
        if o.type == "dexpilot":
            np.testing.assert_array_equal(o.projected, VEC[f"{case}/projected"][i])


def test_dexpilot_cases_exercise_the_projection():
    assert VEC["leap_dexpilot/projected"].any() and not VEC["leap_dexpilot/projected"].all()
    assert VEC["ability_dexpilot_mimic/projected"].any()


@pytest.mark.parametrize("case", CASES)
def test_reference_mode_solver_matches_reference_retarget(case):
    """Optimizer.retarget (optimizer.py:77-102) through the SLSQP stand-in == oracle mode A on the same inputs."""
    o = build_oracle(str(VEC[f"{case}/key"]))
    refs, fixed, x0 = VEC[f"{case}/ref_value"], VEC[f"{case}/fixed_qpos"], VEC[f"{case}/last_qpos"]
    for i in range(refs.shape[0]):
        if o.type == "dexpilot":
            o.projected[:] = False
        xa, _ = solve_reference(o, refs[i], fixed[i], x0[i])
        np.testing.assert_allclose(xa, VEC[f"{case}/retarget"][i], atol=2e-6)


@pytest.mark.parametrize("stream", STREAMS)
def test_sequence_wrapper_matches_reference_stream(stream):
    """SeqRetargeting.retarget (seq_retarget.py:112-134) + LPFilter (optimizer_utils.py:7-13) + the DexPilot hysteresis
    carried from frame to frame, over the recorded keypoint trajectory."""
    o = build_oracle(str(VEC[f"{stream}/key"]))
    oseq = OracleSeqRetargeting(o, mode="reference")
    kp = VEC[f"{stream}/keypoints"].astype(np.float64)
    want = VEC[f"{stream}/robot_qpos"]
    for t in range(kp.shape[0]):
        got = oseq.retarget(o.ref_from_keypoints(kp[t]), fixed_qpos=np.zeros(len(o.idx_pin2fixed)))
        np.testing.assert_allclose(got, want[t], atol=5e-6, err_msg=f"frame {t}")
        if o.type == "dexpilot":
            np.testing.assert_array_equal(o.projected, VEC[f"{stream}/projected"][t])
