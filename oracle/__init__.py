"""CPU oracle for the retargeting hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs may
import this package.  Nothing under `dex_retargeting_b200/` imports it, and the product path fails
loudly when the CUDA library is missing instead of falling back to anything here.

What it restates (reference file:line, relative to /root/reference):
  * oracle/robot.py      (+ oracle/c/orc.c, the same arithmetic in C for speed, cross-checked in
                         tests/test_oracle_c.py)
                         robot_wrapper.py:8-95 + the pinocchio calls it wraps (FK, frame placement,
                         LOCAL frame Jacobian rotated to world axes), retargeting_config.py:167-257
                         (model construction order, dummy joints), kinematics_adaptor.py:46-113
  * oracle/objectives.py optimizer.py:116-200 (position), :203-306 (vector), :309-577 (dexpilot)
  * oracle/solvers.py    optimizer.py:77-102 (nlopt LD_SLSQP driver; scipy's SLSQP -- the same Kraft
                         code -- is the stand-in), seq_retarget.py:112-134, optimizer_utils.py:1-17

PARITY UNPINNED at the third-party boundary, pinned above it.  pinocchio (pin>=3.3.1) and nlopt
(nlopt>=2.8.0) are not vendored, not installed in the build container and not installable offline
(pyproject.toml:30-38), and the reference's own tests hold no golden joint vectors -- only the bar
"mean task-space error < 1e-2 m over 100 seeded problems" (tests/test_optimizer.py:141,209,278) and
two docstring examples (optimizer.py:411-412, :434-438).  So:
  * everything ABOVE those two packages is pinned by executing the reference's own optimizer.py /
    kinematics_adaptor.py / seq_retarget.py / optimizer_utils.py on seeded inputs with the two packages
    shimmed (tests/tools/gen_reference_vectors.py -> tests/golden/reference_vectors.npz): objective
    values, gradients, DexPilot flags, retarget() results and SeqRetargeting streams, reproduced by
    this oracle in tests/test_reference_vectors.py;
  * the robot model INPUT (joint tree, origins, axes, limits, mimic lists, dummy free joints) is pinned
    by executing the reference's own yourdfpy.py parser on every hand URDF
    (tests/tools/gen_reference_urdf_vectors.py -> tests/golden/reference_urdf_vectors.npz,
    tests/test_reference_urdf_vectors.py);
  * pinocchio's FK/Jacobian VALUES and nlopt's SLSQP ITERATES stay unpinned: the oracle's kinematics
    are checked by internal consistency only (two independent FK implementations, analytic vs
    finite-difference Jacobians) and scipy's SLSQP stands in for nlopt's.

Two solver modes:
  mode A "reference-faithful": objective value WITHOUT the norm_delta term, gradient WITH it
         (optimizer.py:166-167 vs :194), SLSQP stopped at ftol 1e-5 / 1e-6 (optimizer.py:136,239,397).
  mode B "converged": the minimiser of the consistent objective L(x) + norm_delta*|x-x_last|^2 inside
         the widened bounds, polished in float64 until the projected gradient vanishes.  This is
         the joint-space parity target (|dq|_inf < 1e-4 rad) for the CUDA solver.
"""
