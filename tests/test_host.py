"""CPU-only checks of the host side: the C-ABI library loads and exports exactly what
include/hyperb200.h declares, fails loudly without a GPU, and the window generator / sharding
logic behaves."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from hyperslam_b200 import runtime, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    with open(os.path.join(ROOT, "include", "hyperb200.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hb200_[a-z_0-9]+)\s*\(", text)) - {"hb200_allreduce_fn"})


def test_library_exports_every_declared_symbol(built):
    lib = runtime.load_library()
    declared = header_functions()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/hyperb200.h but not exported"
    assert sorted(runtime.EXPORTS) == declared, "runtime.EXPORTS out of sync with the header"


def test_no_cpu_fallback(built):
    """Without a CUDA device the product path refuses to run (no oracle / CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(runtime.HB200Error) as e:
        runtime.Context(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_package_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "hyperslam_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert "oracle_lib" not in src and "hyper_oracle" not in src and "ho_window" not in src, fn


def test_synthetic_configs_match_baseline_json():
    w = synthetic.make_config(1)
    assert w.order == 4 and w.knots.shape == (50, 8) and w.v_stamp.size == 10000 and w.i_stamp.size == 2000 and w.landmarks.shape == (1000, 3)
    assert w.reduced_size() == 326
    assert np.all(np.diff(w.v_stamp) >= 0) and np.all(np.diff(w.i_stamp) > 0)
    # every landmark: 5 frames x 2 cameras inside the image
    counts = np.bincount(w.v_lm, minlength=1000)
    assert np.all(counts == 10)
    px, _ = synthetic.pixel_model(w.truth["knots"], w.order, w.cameras, w.truth["landmarks"], w.v_stamp, w.v_cam, w.v_lm)
    assert np.all((px[:, 0] > 0) & (px[:, 0] < 752) & (px[:, 1] > 0) & (px[:, 1] < 480))
    assert np.abs(px - w.v_pixel).std() < 1.0           # 0.5 px noise
    t_lo, t_hi = w.knots[1, 7], w.knots[48, 7]
    assert w.v_stamp.min() >= t_lo and w.v_stamp.max() < t_hi and w.i_stamp.min() >= t_lo and w.i_stamp.max() < t_hi
    w2 = synthetic.make_config(1)
    assert np.array_equal(w.v_pixel, w2.v_pixel) and np.array_equal(w.knots, w2.knots)   # seeded
    w0 = synthetic.make_config(0)
    assert w0.knots.shape[0] == 8 and w0.i_stamp.size == 200 and w0.v_stamp.size == 0
    small = synthetic.make_config(2, scale=0.01)
    assert small.order == 6 and small.knots.shape[0] == 200


def test_shards_partition_factors_by_landmark_owner():
    w = synthetic.make_window(order=4, num_knots=20, num_landmarks=101, num_imu=333, seed=4)
    for world in (1, 2, 4, 8):
        shards = [w.shard(r, world) for r in range(world)]
        assert sum(s.v_stamp.size for s in shards) == w.v_stamp.size
        assert sum(s.i_stamp.size for s in shards) == w.i_stamp.size
        owners = [set(s.v_lm.tolist()) for s in shards]
        for a in range(world):
            for b in range(a + 1, world):
                assert not owners[a] & owners[b]
        for s in shards:
            assert s.knots is w.knots and s.landmarks is w.landmarks       # state replicated, not copied
            assert np.all(np.diff(s.v_stamp) >= 0)


def test_numpy_forward_model_is_consistent():
    """Truth windows generated without noise have (numerically) zero residual under the generator's
    own forward model, and the body rates integrate the rotation."""
    w = synthetic.make_window(order=6, num_knots=16, num_landmarks=20, num_imu=50, seed=8, perturb=False, noise=False, generic_calibration=True)
    px, _ = synthetic.pixel_model(w.knots, w.order, w.cameras, w.landmarks, w.v_stamp, w.v_cam, w.v_lm)
    assert np.abs(px - w.v_pixel).max() < 1e-10
    pred = synthetic.inertial_model(w.knots, w.order, w.imu, w.gyro_bias, w.accel_bias, w.bias_order, w.gravity, w.i_stamp)
    assert np.abs(pred - w.i_meas).max() < 1e-12
    t = w.i_stamp[10:12].mean() + np.array([-1e-6, 0, 1e-6])
    R, p, om, al, pdd, pd = synthetic.spline_eval(w.knots, w.order, t)
    wx = R[1].T @ (R[2] - R[0]) / 2e-6
    assert np.allclose([wx[2, 1], wx[0, 2], wx[1, 0]], om[1], atol=1e-6)
    assert np.allclose((p[2] - p[0]) / 2e-6, pd[1], atol=1e-6)
