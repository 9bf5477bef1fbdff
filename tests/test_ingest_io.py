"""Ingest (pixels -> bearings, stereo triangulation) and trajectory output around the hot path
(reference abstract.cpp:186-264, main.cpp:56-83, evaluation/conversions.py)."""
import os

import numpy as np
import pytest

import oracle_lib as ol
from hyperslam_b200 import synthetic, trajectory


def stereo_frame(win, stamp, n, rng, noise=0.0):
    """n landmarks seen by cameras 0 and 1 at `stamp`: pixels from the numpy forward model."""
    k = win.order
    truth = win.truth
    base = int(np.searchsorted(truth["knots"][:, 7], stamp, side="right") - 1 - (k - 1) // 2)
    lm_idx = rng.choice(win.landmarks.shape[0], size=n, replace=False)
    st = np.full(n, stamp)
    px0, ps0 = synthetic.pixel_model(truth["knots"], k, win.cameras, truth["landmarks"], st, np.zeros(n, dtype=int), lm_idx)
    px1, ps1 = synthetic.pixel_model(truth["knots"], k, win.cameras, truth["landmarks"], st, np.ones(n, dtype=int), lm_idx)
    ok = (ps0[:, 2] > 0.3) & (ps1[:, 2] > 0.3)
    return base, lm_idx[ok], px0[ok] + rng.normal(0, noise, px0[ok].shape), px1[ok] + rng.normal(0, noise, px1[ok].shape), ps0[ok], ps1[ok]


def make_truth_window(order):
    return synthetic.make_window(order=order, num_knots=16, num_landmarks=200, num_imu=0, seed=synthetic.SEED_BASE + 700 + order, perturb=False)


@pytest.mark.parametrize("order", [4, 6])
def test_oracle_ingest_is_exact_on_noise_free_tracks(order):
    """distort(undistort(pixel)) = pixel, bearings are the unit rays of the landmark in each sensor frame, and the
    midpoint triangulation returns the landmark itself when the rays intersect."""
    win = make_truth_window(order)
    rng = np.random.default_rng(5)
    stamp = 0.731
    base, lm, px0, px1, ps0, ps1 = stereo_frame(win, stamp, 60, rng)
    assert lm.size > 20
    B0, B1, L = ol.ingest_stereo_frame(win.knots[base:base + order], stamp, win.cameras[0], win.cameras[1], px0, px1)
    assert np.abs(B0 - ps0 / np.linalg.norm(ps0, axis=1, keepdims=True)).max() < 1e-12
    assert np.abs(B1 - ps1 / np.linalg.norm(ps1, axis=1, keepdims=True)).max() < 1e-12
    assert np.abs(np.linalg.norm(B0, axis=1) - 1).max() < 1e-15
    assert np.abs(L - win.truth["landmarks"][lm]).max() < 1e-9
    # re-projecting the bearing through the forward camera model gives the pixel back
    back = synthetic.project(win.cameras, np.zeros(lm.size, dtype=int), B0)
    assert np.abs(back - px0).max() < 1e-9


def test_trajectory_format_and_tum_conversion(tmp_path):
    """One line per sample: stamp, qx, qy, qz, qw, px, py, pz in %.20e separated by ', ' (reference main.cpp:57-79);
    TUM conversion reorders to stamp, position, quaternion (evaluation/conversions.py:7)."""
    lower, upper = trajectory.state_range(np.arange(10) * 0.1, 4)
    assert (lower, upper) == (0.1, 0.7000000000000001) or (abs(lower - 0.1) < 1e-15 and abs(upper - 0.7) < 1e-12)
    stamps = trajectory.sample_range(0.1, 0.7, 100)
    assert stamps.size == 60 and abs(stamps[0] - 0.1) < 1e-15 and stamps[-1] < 0.7
    line = trajectory.format_line(1.5, np.array([0.0, 0.0, 0.6, 0.8, 1.0, -2.0, 3.25]))
    fields = line.split(", ")
    assert len(fields) == 8 and fields[0] == "1.50000000000000000000e+00" and fields[5] == "1.00000000000000000000e+00"
    src, dst = tmp_path / "estimation.hyper", tmp_path / "estimation.tum"
    rows = np.random.default_rng(1).normal(size=(5, 8))
    with open(src, "w") as f:
        for r in rows:
            f.write(trajectory.format_line(r[0], r[1:]) + "\n")
    trajectory.convert_hyper_to_tum_format(str(src), str(dst))
    tum = np.loadtxt(dst)
    assert tum.shape == (5, 8) and np.array_equal(tum, rows[:, [0, 5, 6, 7, 1, 2, 3, 4]])
    assert len(open(dst).readline().split(" ")) == 8


@pytest.mark.gpu
@pytest.mark.parametrize("order", [4, 6])
def test_gpu_ingest_matches_oracle(built, order):
    from hyperslam_b200 import runtime
    win = make_truth_window(order)
    rng = np.random.default_rng(9)
    ctx = runtime.Context(0)
    ctx.load_window(win)
    stamps, c0, c1, p0, p1, refs = [], [], [], [], [], []
    for stamp in (0.41, 0.655, 0.93):
        base, lm, px0, px1, _, _ = stereo_frame(win, stamp, 50, rng, noise=0.4)
        B0, B1, L = ol.ingest_stereo_frame(win.knots[base:base + order], stamp, win.cameras[0], win.cameras[1], px0, px1)
        stamps += [stamp] * lm.size; c0 += [0] * lm.size; c1 += [1] * lm.size
        p0.append(px0); p1.append(px1); refs.append((B0, B1, L))
    b0, b1, lm_w, bad = ctx.ingest_stereo(np.array(stamps), np.array(c0), np.array(c1), np.concatenate(p0), np.concatenate(p1))
    assert bad == 0
    assert np.abs(b0 - np.concatenate([r[0] for r in refs])).max() < 1e-12
    assert np.abs(b1 - np.concatenate([r[1] for r in refs])).max() < 1e-12
    ref_l = np.concatenate([r[2] for r in refs])
    assert np.abs(lm_w - ref_l).max() < 1e-9 * max(1.0, np.abs(ref_l).max())
    # invalid stamp / camera are reported, not silently processed
    _, _, _, bad = ctx.ingest_stereo(np.array([99.0, 0.5]), np.array([0, 7]), np.array([1, 1]), np.zeros((2, 2)), np.zeros((2, 2)))
    assert bad == 2
    ctx.close()


@pytest.mark.gpu
def test_gpu_estimation_dump_matches_oracle_interpolation(built, tmp_path):
    from hyperslam_b200 import runtime
    win = synthetic.make_window(order=4, num_knots=14, num_landmarks=30, num_imu=40, seed=synthetic.SEED_BASE + 720)
    ctx = runtime.Context(0)
    ctx.load_window(win)
    path = tmp_path / "estimation.hyper"
    n = trajectory.write_estimation(ctx, str(path), win.knots[:, 7], win.order, root=100.0)
    data = np.loadtxt(path, delimiter=",")
    lower, upper = trajectory.state_range(win.knots[:, 7], win.order)
    assert data.shape == (n, 8) and n == trajectory.sample_range(lower, upper).size
    left = (win.order - 1) // 2
    stamps = trajectory.sample_range(lower, upper)
    for i in range(0, n, max(1, n // 25)):
        t, row = stamps[i], data[i]
        assert abs(row[0] - (100.0 + t)) < 1e-12
        j = int(np.searchsorted(win.knots[:, 7], t, side="right") - 1) - left
        v, _, _, _ = ol.state_evaluate(win.knots[j:j + win.order], t, 0, False)
        if np.dot(v[:4], row[1:5]) < 0:
            v[:4] = -v[:4]          # q and -q are the same rotation
        assert np.abs(row[1:] - v).max() < 1e-12, (i, t, row[1:], v)
    ctx.close()
