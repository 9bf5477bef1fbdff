"""GPU test of the C++ host plugin surface (hyperslam_b200/host): hyper::Optimizer,
ExteroceptiveCost::Evaluate, Manifold hooks, ContinuousState::evaluate -- driven by the C++ test
binary hyperslam_b200/lib/host_test, checked here against the oracle."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
from hyperslam_b200 import synthetic

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def dump_window(win, path, nconst):
    with open(path, "w") as f:
        w = lambda a: f.write(" ".join(repr(float(x)) for x in np.asarray(a).ravel()) + "\n")
        f.write(f"{win.order} {win.knots.shape[0]} {win.gyro_bias.shape[0]} {win.cameras.shape[0]} {win.landmarks.shape[0]} "
                f"{win.v_stamp.size} {win.i_stamp.size} {nconst} {win.b_stamp.size} {win.m_stamp.size} {win.pose_sensors.shape[0]}\n")
        w(win.knots); w(win.gyro_bias); w(win.accel_bias); w(win.gravity); w(win.cameras); w(win.imu); w(win.landmarks)
        for i in range(win.v_stamp.size):
            f.write(f"{float(win.v_stamp[i])!r} {int(win.v_cam[i])} {int(win.v_lm[i])} {float(win.v_pixel[i, 0])!r} {float(win.v_pixel[i, 1])!r}\n")
        for i in range(win.i_stamp.size):
            f.write(f"{float(win.i_stamp[i])!r} " + " ".join(repr(float(x)) for x in win.i_meas[i]) + "\n")
        for i in range(win.b_stamp.size):
            f.write(f"{float(win.b_stamp[i])!r} {int(win.b_cam[i])} {int(win.b_lm[i])} " + " ".join(repr(float(x)) for x in win.b_bearing[i]) + "\n")
        w(win.pose_sensors) if win.pose_sensors.size else None
        for i in range(win.m_stamp.size):
            f.write(f"{float(win.m_stamp[i])!r} {int(win.m_sensor[i])} " + " ".join(repr(float(x)) for x in win.m_pose[i]) + "\n")


@pytest.mark.parametrize("order,widened", [(4, False), (6, False), (4, True), (6, True)])
def test_cpp_plugin_surface(built, tmp_path, order, widened):
    exe = os.path.join(ROOT, "hyperslam_b200", "lib", "host_test")
    assert os.path.exists(exe), "host_test not built (python -c 'import __graft_entry__ as g; g.build()')"
    win = synthetic.make_window(order=order, num_knots=14, num_landmarks=30, num_imu=40, seed=synthetic.SEED_BASE + 500 + order, constant_knots=2)
    if widened:   # + VisualBearingObservation / ManifoldObservation costs (reference optimizer.cpp:189-251)
        win = synthetic.add_bearing_and_pose_factors(win, num_bearing=40, num_pose=12, seed=synthetic.SEED_BASE + 520 + order)
    src, dst = tmp_path / "window.txt", tmp_path / "out.txt"
    dump_window(win, src, 2)
    res = subprocess.run([exe, str(src), str(dst)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    lines = open(dst).read().split("\n")
    it = iter(lines)
    head = next(it).split()
    assert head[0] == "probe" and int(head[1]) >= (14 if widened else 8) and int(head[2]) == 0          # reference Probe protocol passes
    ow = ol.OracleWindow(win)
    vb, ib, ig, ia = ow.index_maps()
    k, kb = win.order, win.bias_order
    n_costs = 0
    line = next(it)
    while line.startswith("cost"):
        _, kind, index, nr, npar, nb, *sizes = line.split()
        kind, index, nr, nb = int(kind), int(index), int(nr), int(nb)
        sizes = [int(s) for s in sizes]
        r = np.array(next(it).split(), dtype=float)
        J = [np.array(next(it).split(), dtype=float).reshape(nr, s) for s in sizes]
        if kind == 0:
            cam = win.cameras[win.v_cam[index]]
            blocks = [win.knots[vb[index] + m] for m in range(k)] + [cam[:7], cam[7:11], cam[11:15], win.landmarks[win.v_lm[index]]]
            r_o, J_o = ol.cost_evaluate(ol.PIXEL, win.v_stamp[index], win.v_pixel[index], np.concatenate(blocks), k=k)
            assert sizes == [8] * k + [7, 4, 4, 3] and int(npar) == 8 * k + 18
            var = list(range(k)) + [k + 3]
        elif kind == 2:
            base = int(np.searchsorted(win.knots[:, 7], win.b_stamp[index], side="right") - 1 - (k - 1) // 2)
            cam = win.cameras[win.b_cam[index]]
            blocks = [win.knots[base + m] for m in range(k)] + [cam[:7], cam[7:11], cam[11:15], win.landmarks[win.b_lm[index]]]
            r_o, J_o = ol.cost_evaluate(ol.BEARING, win.b_stamp[index], win.b_bearing[index], np.concatenate(blocks), k=k)
            assert nr == 1 and sizes == [8] * k + [7, 4, 4, 3]
            var = list(range(k)) + [k + 3]
        elif kind == 3:
            base = int(np.searchsorted(win.knots[:, 7], win.m_stamp[index], side="right") - 1 - (k - 1) // 2)
            blocks = [win.knots[base + m] for m in range(k)] + [win.pose_sensors[win.m_sensor[index]]]
            r_o, J_o = ol.cost_evaluate(ol.MANIFOLD, win.m_stamp[index], win.m_pose[index], np.concatenate(blocks), k=k)
            assert nr == 6 and sizes == [8] * k + [7]
            var = list(range(k))
        else:
            imu = win.imu
            blocks = ([win.knots[ib[index] + m] for m in range(k)] + [imu[:7], imu[7:13], imu[13:19], imu[19:28], imu[28:37]]
                      + [win.gyro_bias[ig[index] + m] for m in range(kb)] + [win.accel_bias[ia[index] + m] for m in range(kb)] + [win.gravity])
            r_o, J_o = ol.cost_evaluate(ol.INERTIAL, win.i_stamp[index], win.i_meas[index], np.concatenate(blocks), k=k)
            assert sizes == [8] * k + [7, 6, 6, 9, 9] + [4] * 8 + [3]
            var = list(range(k)) + list(range(k + 5, k + 14))
        assert np.abs(r - r_o).max() < 1e-9 * max(1.0, np.abs(r_o).max())
        scale = max(np.abs(J_o[b]).max() for b in var)       # rounding is relative to the factor's Jacobian, not to a tiny edge block
        for b in var:
            mid = ol.M_STATE if b < k else (ol.M_SPHERE if (kind == 1 and b == k + 13) else (ol.M_BIAS if kind == 1 else ol.M_EUCLIDEAN))
            PJ = ol.manifold_plus_jacobian(mid, blocks[b])
            a, o = J[b] @ PJ, J_o[b] @ PJ
            assert np.abs(a - o).max() < 1e-8 * scale, (kind, index, b)
        n_costs += 1
        line = next(it)
    assert n_costs == (8 if widened else 4)
    assert line.startswith("interp")
    left = (k - 1) // 2
    for _ in range(int(line.split()[1])):
        vals = np.array(next(it).split(), dtype=float)
        t = vals[0]
        j = int(np.searchsorted(win.knots[:, 7], t, side="right") - 1) - left
        v, ve, ac, _ = ol.state_evaluate(win.knots[j:j + k], t, 2, False)
        assert np.abs(vals[1:8] - v).max() < 1e-12 and np.abs(vals[8:14] - ve).max() < 1e-9 and np.abs(vals[14:20] - ac).max() < 1e-7
    line = next(it)
    assert line.startswith("optimize")
    for i in range(int(line.split()[1])):
        cost, cost_new, accepted, spd = next(it).split()
        o = ow.iterate(apply=True)
        assert abs(float(cost) - o["cost"]) <= 1e-7 * o["cost"] and int(accepted) == o["accepted"] and int(spd) == 1
    assert next(it).strip() == "knots"
    knots = np.array([next(it).split() for _ in range(win.knots.shape[0])], dtype=float)
    assert np.abs(knots - ow.state()["knots"]).max() < 1e-6
    assert next(it).strip() == "landmarks"
    lms = np.array([next(it).split() for _ in range(win.landmarks.shape[0])], dtype=float)
    assert np.abs(lms - ow.state()["landmarks"]).max() < 1e-6
