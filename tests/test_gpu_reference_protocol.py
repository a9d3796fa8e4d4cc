"""The reference's own test suite for this path, run against the drop-in through the UNCHANGED public API.

tests/test_optimizer.py of the reference builds every default config (7 robots x 2 hands x position / vector / dexpilot = 40
parametrisations), draws 100 seeded reachable targets per config (np.random.seed(1); target pose uniform in the joint
limits, cold start = pose + 0.5 N(0,1) clipped), calls `retargeting.set_qpos(init)` / `retargeting.retarget(target,
fixed_qpos=...)` (position) or `optimizer.retarget(target, fixed_qpos, last_qpos)` (vector, dexpilot) ONE FRAME AT A TIME, and asserts a mean task-space error < 1e-2 m (:141, :209, :278).  This file is that
protocol, written against the same names (`RetargetingConfig.load_from_file`, `get_default_config_path`, `ROBOT_NAMES`,
`optimizer.idx_pin2target`, `robot.get_link_pose`, `robot.model.nq` ...), so it reads like the reference's test and shows
what a user switching packages would see.  Every call goes numpy -> C ABI -> CUDA kernel -> numpy (B = 1)."""
import numpy as np
import pytest

from helpers import ROBOTS

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from dex_retargeting_b200.constants import ROBOT_NAMES, HandType, RetargetingType, RobotName, get_default_config_path  # noqa: E402
from dex_retargeting_b200.optimizer import DexPilotOptimizer, PositionOptimizer, VectorOptimizer  # noqa: E402
from dex_retargeting_b200.retargeting_config import RetargetingConfig  # noqa: E402

NUM_OPTIMIZATION = 100
DEXPILOT_ROBOT_NAMES = [r for r in ROBOT_NAMES if r is not RobotName.ability]  # tests/test_optimizer.py:24-25


def sample_qpos(optimizer):
    robot, adaptor = optimizer.robot, optimizer.adaptor
    limit = robot.joint_limits
    random_qpos = np.random.uniform(limit[:, 0], limit[:, 1])
    if adaptor is not None:
        random_qpos = adaptor.forward_qpos(random_qpos)
    init_qpos = np.clip(random_qpos + np.random.randn(robot.dof) * 0.5, limit[:, 0] + 1e-5, limit[:, 1] - 1e-5)
    return random_qpos, init_qpos


def pin_qpos(optimizer, qpos, fixed_qpos):
    full = np.zeros(optimizer.robot.model.nq)
    full[optimizer.idx_pin2target] = qpos
    full[optimizer.idx_pin2fixed] = fixed_qpos
    return optimizer.adaptor.forward_qpos(full) if optimizer.adaptor is not None else full


def link_positions(robot, qpos, link_indices):
    robot.compute_forward_kinematics(qpos)
    return np.array([robot.get_link_pose(i)[:3, 3] for i in link_indices])


def run_protocol(retargeting, vector: bool):
    optimizer = retargeting.optimizer
    robot = optimizer.robot
    errors = []
    np.random.seed(1)
    for _ in range(NUM_OPTIMIZATION):
        random_qpos, init_qpos = sample_qpos(optimizer)
        if vector:
            pos = link_positions(robot, random_qpos, optimizer.computed_link_indices)
            target = pos[optimizer.task_link_indices] - pos[optimizer.origin_link_indices]
        else:
            target = link_positions(robot, random_qpos, optimizer.target_link_indices)
        fixed_qpos = random_qpos[optimizer.idx_pin2fixed]
        if vector:
            # the reference's vector / dexpilot tests call the optimizer directly (:179-185): with low_pass_alpha = 0 the
            # sequence wrapper's filter would hold its first output for ever (optimizer_utils.py:12), here as there
            computed = optimizer.retarget(target, fixed_qpos=fixed_qpos, last_qpos=init_qpos[optimizer.idx_pin2target])
        else:
            retargeting.set_qpos(init_qpos)
            computed = retargeting.retarget(target, fixed_qpos=fixed_qpos)[optimizer.idx_pin2target]
        full = pin_qpos(optimizer, computed, fixed_qpos)
        if vector:
            pos = link_positions(robot, full, optimizer.computed_link_indices)
            got = pos[optimizer.task_link_indices] - pos[optimizer.origin_link_indices]
        else:
            got = link_positions(robot, full, optimizer.target_link_indices)
        errors.append(np.mean(np.linalg.norm(got - target, axis=-1)))
    return float(np.mean(errors))


@pytest.fixture(autouse=True)
def _urdf_dir():
    RetargetingConfig.set_default_urdf_dir(str(RetargetingConfig.packaged_urdf_dir()))


@pytest.mark.parametrize("robot_name", ROBOT_NAMES, ids=lambda r: r.name)
@pytest.mark.parametrize("hand_type", list(HandType), ids=lambda h: h.name)
def test_position_optimizer(robot_name, hand_type):
    config_path = get_default_config_path(robot_name, RetargetingType.position, hand_type)
    retargeting = RetargetingConfig.load_from_file(config_path, dict(normal_delta=0)).build()
    assert isinstance(retargeting.optimizer, PositionOptimizer)
    assert run_protocol(retargeting, vector=False) < 1e-2


@pytest.mark.parametrize("robot_name", ROBOT_NAMES, ids=lambda r: r.name)
@pytest.mark.parametrize("hand_type", list(HandType), ids=lambda h: h.name)
def test_vector_optimizer(robot_name, hand_type):
    config_path = get_default_config_path(robot_name, RetargetingType.vector, hand_type)
    retargeting = RetargetingConfig.load_from_file(config_path, dict(low_pass_alpha=0, scaling_factor=1.0, normal_delta=0)).build()
    assert isinstance(retargeting.optimizer, VectorOptimizer) and retargeting.optimizer.retargeting_type == "VECTOR"
    assert run_protocol(retargeting, vector=True) < 1e-2


@pytest.mark.parametrize("robot_name", DEXPILOT_ROBOT_NAMES, ids=lambda r: r.name)
@pytest.mark.parametrize("hand_type", list(HandType), ids=lambda h: h.name)
def test_dexpilot_optimizer(robot_name, hand_type):
    config_path = get_default_config_path(robot_name, RetargetingType.dexpilot, hand_type)
    retargeting = RetargetingConfig.load_from_file(config_path, dict(low_pass_alpha=0, scaling_factor=1.0, normal_delta=0)).build()
    assert isinstance(retargeting.optimizer, DexPilotOptimizer) and retargeting.optimizer.retargeting_type == "DEXPILOT"
    assert run_protocol(retargeting, vector=True) < 1e-2


def test_alpha_zero_filter_holds_first_output_like_the_reference():
    """optimizer_utils.py:7-13 with alpha = 0: y <- y + 0 * (x - y).  The drop-in keeps that (surprising) behaviour."""
    config_path = get_default_config_path(RobotName.allegro, RetargetingType.vector, HandType.right)
    retargeting = RetargetingConfig.load_from_file(config_path, dict(low_pass_alpha=0, scaling_factor=1.0, normal_delta=0)).build()
    optimizer = retargeting.optimizer
    np.random.seed(3)
    outs = []
    for _ in range(3):
        random_qpos, init_qpos = sample_qpos(optimizer)
        pos = link_positions(optimizer.robot, random_qpos, optimizer.computed_link_indices)
        retargeting.set_qpos(init_qpos)
        outs.append(retargeting.retarget(pos[optimizer.task_link_indices] - pos[optimizer.origin_link_indices]))
    np.testing.assert_array_equal(outs[0], outs[1])
    np.testing.assert_array_equal(outs[0], outs[2])
    assert np.abs(retargeting.last_qpos - outs[0][optimizer.idx_pin2target]).max() > 1e-3  # the solver itself did move
