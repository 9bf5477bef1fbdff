"""CPU tests of the oracle's restatement of ceres::TrustRegionMinimizer's termination tests (oracle/ho_window.h
`Termination`, `window_optimize`): the checker the GPU termination tests compare accept_kernel with.  The reference leaves
the tolerances at Ceres' defaults (reference internal/hyper/optimizers/ceres/optimizer.cpp:38-54)."""
import numpy as np
import pytest

import oracle_lib as ol
from hyperslam_b200 import synthetic


def window(order=4, seed_off=0):
    return synthetic.make_window(order=order, num_knots=18, num_landmarks=120, num_imu=300, seed=synthetic.SEED_BASE + 3000 + order + seed_off, constant_knots=2)


def test_no_tolerance_runs_every_iteration():
    win = window()
    o = ol.OracleWindow(win).optimize(12, function_tolerance=0.0, gradient_tolerance=0.0, parameter_tolerance=0.0, min_radius=1e-300)
    assert o["iterations"] == 12 and o["type"] == 0
    accepted_costs = [c for c, a in zip(o["cost_new"], o["accepted"]) if a]
    assert all(b <= a * (1 + 1e-12) for a, b in zip(accepted_costs, accepted_costs[1:]))   # accepted steps never raise the cost
    assert o["cost"][0] > o["cost"][-1]


@pytest.mark.parametrize("order", [4, 6])
def test_ceres_defaults_end_by_function_tolerance(order):
    win = window(order)
    ow = ol.OracleWindow(win)
    o = ow.optimize(40)
    assert 0 < o["iterations"] < 40
    assert o["type"] == 1                                           # Ceres: CONVERGENCE by function tolerance on this problem
    # `iterations` counts the COMPLETED iterations: the one whose trial step met the tolerance ends the solve without being
    # applied or recorded -- so no recorded iteration may satisfy the test
    for i in range(o["iterations"]):
        assert abs(o["cost"][i] - o["cost_new"][i]) > 1e-6 * o["cost"][i], i


def test_each_tolerance_fires_with_its_own_type_and_leaves_the_trial_step_unapplied():
    win = window()
    free = ol.OracleWindow(win).optimize(40, function_tolerance=0.0, gradient_tolerance=0.0, parameter_tolerance=0.0, min_radius=1e-300)
    for kw, code in ((dict(function_tolerance=1e-3, gradient_tolerance=0.0, parameter_tolerance=0.0), 1),
                     (dict(function_tolerance=0.0, gradient_tolerance=0.0, parameter_tolerance=3e-2), 2),
                     (dict(function_tolerance=0.0, gradient_tolerance=20.0, parameter_tolerance=0.0), 3)):
        ow = ol.OracleWindow(win)
        o = ow.optimize(40, min_radius=1e-32, **kw)
        assert o["type"] == code, (kw, o["type"])
        n = o["iterations"]
        assert 0 < n < 40
        # the completed iterations are the unconstrained trajectory's first n
        assert np.allclose(o["cost"], free["cost"][:n], rtol=1e-12, atol=0)
        assert np.array_equal(o["accepted"], free["accepted"][:n])
        if code == 2:
            assert o["step_norm"] <= 3e-2 * (o["x_norm"] + 3e-2)
        if code == 3:
            assert o["gradient_max_norm"] <= 20.0


def test_minimum_radius_ends_the_solve():
    win = window()
    o = ol.OracleWindow(win, radius=1e-3).optimize(40, function_tolerance=0.0, gradient_tolerance=0.0, parameter_tolerance=0.0, min_radius=1e-2)
    assert o["type"] == 4 and o["iterations"] <= 2
