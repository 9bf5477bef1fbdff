"""ceres::Solve termination semantics on the device: function / gradient / parameter
tolerance, evaluated in accept_kernel where ceres::TrustRegionMinimizer evaluates them, against the oracle's
restatement of the same loop -- iteration counts, termination type, per-iteration records and the final state.
The reference leaves the tolerances at Ceres' defaults (reference optimizer.cpp:38-54)."""
import numpy as np
import pytest

import oracle_lib as ol
from hyperslam_b200 import runtime, synthetic

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / (np.abs(np.asarray(b)).max() + 1e-300))


CASES = {
    "ceres_defaults": dict(function_tolerance=1e-6, gradient_tolerance=1e-10, parameter_tolerance=1e-8),
    "function": dict(function_tolerance=1e-3, gradient_tolerance=0.0, parameter_tolerance=0.0),
    "parameter": dict(function_tolerance=0.0, gradient_tolerance=0.0, parameter_tolerance=3e-2),
    "gradient": dict(function_tolerance=0.0, gradient_tolerance=20.0, parameter_tolerance=0.0),
}


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("order", [4, 6])
def test_termination_matches_oracle(built, order, name):
    tol = CASES[name]
    win = synthetic.make_window(order=order, num_knots=18, num_landmarks=120, num_imu=300, seed=synthetic.SEED_BASE + 3000 + order, constant_knots=2)
    max_iter = 40
    o = ol.OracleWindow(win).optimize(max_iter, min_radius=1e-32, **tol)
    ctx = runtime.Context(0)
    ctx.load_window(win)
    ctx.set_termination(min_trust_region_radius=1e-32, **tol)
    recs = ctx.iterate(max_iter)
    t = ctx.termination()
    assert t["type"] == o["type"], (t, o["type"], o["iterations"])
    assert t["iterations"] == o["iterations"], (t, o["iterations"])
    assert 0 < t["iterations"] < max_iter            # a tolerance ended the solve, not the iteration limit
    expected_type = {"function": 1, "parameter": 2, "gradient": 3}.get(name)
    if expected_type:
        assert t["type"] == expected_type
    for i in range(t["iterations"]):
        assert abs(recs[i]["cost"] - o["cost"][i]) <= 1e-7 * o["cost"][i]
        assert abs(recs[i]["cost_new"] - o["cost_new"][i]) <= 1e-6 * o["cost_new"][i]
        assert recs[i]["accepted"] == o["accepted"][i]
    for i in range(t["iterations"], max_iter):       # iterations after the termination are no-ops with zero records
        assert recs[i]["cost"] == 0.0 and recs[i]["accepted"] == 0
    assert rel_err(t["gradient_max_norm"], o["gradient_max_norm"]) < 1e-4
    if o["step_norm"] > 0 and t["type"] in (1, 2):   # (the gradient test fires before the iteration's step exists)
        assert rel_err(t["step_norm"], o["step_norm"]) < 1e-4 and rel_err(t["x_norm"], o["x_norm"]) < 1e-8
    ow = ol.OracleWindow(win)
    ow.optimize(max_iter, min_radius=1e-32, **tol)
    st, so = ctx.state(), ow.state()
    for key in so:
        assert rel_err(st[key], so[key]) < 1e-6, key
    # a new call starts with a clean termination state and continues from the converged point
    recs2 = ctx.iterate(3)
    assert ctx.termination()["iterations"] <= 3
    ctx.close()


def test_termination_off_runs_every_iteration(built):
    win = synthetic.make_window(order=4, num_knots=14, num_landmarks=60, num_imu=150, seed=synthetic.SEED_BASE + 3100, constant_knots=2)
    ctx = runtime.Context(0)
    ctx.load_window(win)
    recs = ctx.iterate(12)
    assert all(r["cost"] > 0 for r in recs)
    assert ctx.termination()["type"] == 0 and ctx.termination()["iterations"] == 12
    ctx.close()


def test_optimize_restarts_the_trust_region_like_ceres_solve(built):
    """hb200_optimize == CeresOptimizer::optimize(): every call starts at initial_trust_region_radius."""
    win = synthetic.make_window(order=4, num_knots=14, num_landmarks=60, num_imu=150, seed=synthetic.SEED_BASE + 3101, constant_knots=2)
    ctx = runtime.Context(0)
    ctx.load_window(win)
    st = ctx.state()
    recs = ctx.optimize(2, st["knots"], st["gyro_bias"], st["accel_bias"], st["gravity"], st["landmarks"])
    assert recs[-1]["radius"] > 1e4                  # accepted steps grew the radius
    ow = ol.OracleWindow(win)
    ow.iterate(apply=True); ow.iterate(apply=True)
    so = ow.state()
    ow2 = ol.OracleWindow(type(win)(**{**win.__dict__, "knots": so["knots"], "gyro_bias": so["gyro_bias"], "accel_bias": so["accel_bias"],
                                      "gravity": so["gravity"], "landmarks": so["landmarks"]}))   # fresh solve: radius 1e4 again
    o = ow2.iterate(apply=True)
    recs = ctx.optimize(1, st["knots"], st["gyro_bias"], st["accel_bias"], st["gravity"], st["landmarks"])
    assert abs(recs[0]["cost"] - o["cost"]) <= 1e-7 * o["cost"]
    assert abs(recs[0]["cost_new"] - o["cost_new"]) <= 1e-6 * o["cost_new"]
    ctx.close()
