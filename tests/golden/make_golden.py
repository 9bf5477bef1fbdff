#!/usr/bin/env python
"""Generates tests/golden/*.json: known-answer vectors for the hot path at 80 significant digits.

Independent of both the C oracle and the CUDA kernels:
  * only VALUE formulas are implemented (cumulative B-spline on SO(3) x R^3,
    pinhole + radial-tangential camera, the IMU model of reference inertial.cpp:62-79);
  * body rates come from finite differences of R(t) in time (omega^ = R^T Rdot,
    alpha^ = R^T Rddot - omega^ omega^), NOT from the Sommer recursion the oracle/kernels use;
  * Jacobians come from central differences of the residual along the manifold retractions
    (R <- Exp(theta) R, p <- p + rho, Ceres SphereManifold Plus for gravity), h = 1e-20.
Run:  python tests/golden/make_golden.py      (takes ~1-2 minutes)
"""
import json
import os
import sys

import mpmath as mp
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hyperslam_b200 import synthetic  # noqa: E402

mp.mp.dps = 80
H_T = mp.mpf(10) ** -18    # time step for body-rate finite differences
H_X = mp.mpf(10) ** -20    # state perturbation


def hat(v):
    return mp.matrix([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


def so3_exp(w):
    t = mp.sqrt(w[0] ** 2 + w[1] ** 2 + w[2] ** 2)
    W = hat(w)
    if t == 0:
        return mp.eye(3)
    return mp.eye(3) + (mp.sin(t) / t) * W + ((1 - mp.cos(t)) / t ** 2) * (W * W)


def quat_to_rot(q):
    x, y, z, w = q
    return mp.matrix([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def quat_mul(a, b):
    return [a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1],
            a[3] * b[1] - a[0] * b[2] + a[1] * b[3] + a[2] * b[0],
            a[3] * b[2] + a[0] * b[1] - a[1] * b[0] + a[2] * b[3],
            a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2]]


def quat_exp(d):
    t = mp.sqrt(d[0] ** 2 + d[1] ** 2 + d[2] ** 2)
    if t == 0:
        return [mp.mpf(0)] * 3 + [mp.mpf(1)]
    s = mp.sin(t / 2) / t
    return [s * d[0], s * d[1], s * d[2], mp.cos(t / 2)]


def quat_log(q):
    x, y, z, w = q
    if w < 0:
        x, y, z, w = -x, -y, -z, -w
    n = mp.sqrt(x * x + y * y + z * z)
    if n == 0:
        return [mp.mpf(0)] * 3
    s = 2 * mp.atan2(n, w) / n
    return [s * x, s * y, s * z]


def basis(k, u):
    """lambda_j(u), j = 0..k (cumulative), exact rational blending matrix."""
    Mc = [[mp.mpf(0)] * k for _ in range(k)]
    M = [[mp.mpf(0)] * k for _ in range(k)]
    for s in range(k):
        for n in range(k):
            acc = mp.mpf(0)
            for l in range(s, k):
                acc += (-1) ** (l - s) * mp.binomial(k, l - s) * mp.mpf(k - 1 - l) ** (k - 1 - n)
            M[s][n] = mp.binomial(k - 1, n) / mp.factorial(k - 1) * acc
    for j in range(k):
        for n in range(k):
            Mc[j][n] = sum(M[r][n] for r in range(j, k))
    lam = [sum(Mc[j][n] * u ** n for n in range(k)) for j in range(k)] + [mp.mpf(0)]
    lam[0] = mp.mpf(1)
    return lam


def segment(stamps, k, t):
    j = max(i for i in range(len(stamps)) if stamps[i] <= t)
    return j - (k - 1) // 2, (t - stamps[j]) / (stamps[j + 1] - stamps[j])


def pose(knots, k, t):
    """R(t), p(t) of the split cumulative spline; knots: list of [q(4) p(3) stamp]."""
    stamps = [kn[7] for kn in knots]
    base, u = segment(stamps, k, t)
    lam = basis(k, u)
    q = list(knots[base][:4])
    p = mp.matrix(knots[base][4:7])
    for j in range(1, k):
        qa, qb = knots[base + j - 1][:4], knots[base + j][:4]
        d = quat_log(quat_mul([-qa[0], -qa[1], -qa[2], qa[3]], qb))
        q = quat_mul(q, quat_exp([lam[j] * c for c in d]))
        p = p + lam[j] * (mp.matrix(knots[base + j][4:7]) - mp.matrix(knots[base + j - 1][4:7]))
    return quat_to_rot(q), p


def bias(bknots, kb, t):
    stamps = [b[3] for b in bknots]
    base, u = segment(stamps, kb, t)
    lam = basis(kb, u)
    return sum(((lam[m] - lam[m + 1]) * mp.matrix(bknots[base + m][:3]) for m in range(kb)), mp.matrix([0, 0, 0]))


def vee(M):
    return mp.matrix([(M[2, 1] - M[1, 2]) / 2, (M[0, 2] - M[2, 0]) / 2, (M[1, 0] - M[0, 1]) / 2])


def pixel_residual(st, f):
    R, p = pose(st["knots"], st["k"], f["stamp"])
    cam = st["cams"][f["cam"]]
    R_bs = quat_to_rot(cam[:4]); t_bs = mp.matrix(cam[4:7])
    p_b = R.T * (mp.matrix(st["landmarks"][f["lm"]]) - p)
    p_s = R_bs.T * (p_b - t_bs)
    x, y = p_s[0] / p_s[2], p_s[1] / p_s[2]
    cx, cy, fx, fy, k1, k2, p1, p2 = cam[7:15]
    r2 = x * x + y * y
    rad = 1 + k1 * r2 + k2 * r2 * r2
    dx = x * rad + 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
    dy = y * rad + p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
    return [fx * dx + cx - f["pixel"][0], fy * dy + cy - f["pixel"][1]]


def bearing_residual(st, f):
    """AngularMetric<Bearing>: angle between p_s (reference bearing.cpp:58-60) and the measured direction."""
    R, p = pose(st["knots"], st["k"], f["stamp"])
    cam = st["cams"][f["cam"]]
    R_bs = quat_to_rot(cam[:4]); t_bs = mp.matrix(cam[4:7])
    p_s = R_bs.T * (R.T * (mp.matrix(st["landmarks"][f["lm"]]) - p) - t_bs)
    b = mp.matrix(f["bearing"])
    cr = mp.matrix([p_s[1] * b[2] - p_s[2] * b[1], p_s[2] * b[0] - p_s[0] * b[2], p_s[0] * b[1] - p_s[1] * b[0]])
    return [mp.atan2(mp.sqrt(cr[0] ** 2 + cr[1] ** 2 + cr[2] ** 2), p_s[0] * b[0] + p_s[1] * b[1] + p_s[2] * b[2])]


def rot_log(R):
    c = (R[0, 0] + R[1, 1] + R[2, 2] - 1) / 2
    t = mp.acos(c)
    v = vee(R - R.T)         # vee() averages the two off-diagonal entries: vee(R - R^T) = 2 sin(t) axis
    return v * (t / (2 * mp.sin(t))) if t != 0 else mp.matrix([0, 0, 0])


def pose_residual(st, f):
    """ManifoldMetric<SE3> of T_ws = T_wb (+) T_bs (reference manifold.cpp:38) against the measured pose."""
    R, p = pose(st["knots"], st["k"], f["stamp"])
    T_bs = st["pose_sensors"][f["sensor"]]
    R_ws = R * quat_to_rot(T_bs[:4])
    p_ws = p + R * mp.matrix(T_bs[4:7])
    th = rot_log(R_ws * quat_to_rot(f["pose"][:4]).T)
    d = p_ws - mp.matrix(f["pose"][4:7])
    return [th[0], th[1], th[2], d[0], d[1], d[2]]


def imu_matrix(c):
    return mp.matrix([[c[0], 0, 0], [c[3], c[1], 0], [c[4], c[5], c[2]]])


def inertial_residual(st, f):
    t = f["stamp"]
    R0, p0 = pose(st["knots"], st["k"], t)
    Rp, pp = pose(st["knots"], st["k"], t + H_T)
    Rm, pm = pose(st["knots"], st["k"], t - H_T)
    Rd = (Rp - Rm) / (2 * H_T)
    Rdd = (Rp - 2 * R0 + Rm) / H_T ** 2
    pdd = (pp - 2 * p0 + pm) / H_T ** 2
    wx = R0.T * Rd
    w = vee(wx)
    al = vee(R0.T * Rdd - hat(w) * hat(w))
    imu = st["imu"]
    R_sb = quat_to_rot(imu[:4]).T; t_bs = mp.matrix(imu[4:7])
    I_g, I_a = imu_matrix(imu[7:13]), imu_matrix(imu[13:19])
    S_g = mp.matrix(3, 3); X_a = mp.matrix(3, 3)
    for i in range(3):
        for j in range(3):
            S_g[i, j] = imu[19 + i + 3 * j]; X_a[i, j] = imu[28 + i + 3 * j]
    a_i = R0.T * (pdd - mp.matrix(st["gravity"]))
    F = hat(w) * hat(w) + hat(al)
    a_m = mp.matrix([a_i[r] + sum(F[r, c] * (X_a[c, r] + t_bs[c]) for c in range(3)) for r in range(3)])
    gyro = I_g * R_sb * w + S_g * a_m + bias(st["bg"], st["kb"], t)
    acc = I_a * R_sb * a_m + bias(st["ba"], st["kb"], t)
    z = f["meas"]
    return [gyro[i] - z[i] for i in range(3)] + [acc[i] - z[3 + i] for i in range(3)]


def sphere_plus(x, d):
    x = [mp.mpf(v) for v in x]
    nd = mp.sqrt(d[0] ** 2 + d[1] ** 2)
    if nd == 0:
        return x
    sigma = x[0] ** 2 + x[1] ** 2
    v = [x[0], x[1], mp.mpf(1)]
    beta = mp.mpf(0)
    if sigma == 0:
        if x[2] < 0:
            beta = mp.mpf(2)
    else:
        mu = mp.sqrt(x[2] ** 2 + sigma)
        vp = x[2] - mu if x[2] <= 0 else -sigma / (x[2] + mu)
        beta = 2 * vp * vp / (sigma + vp * vp)
        v[0] /= vp; v[1] /= vp
    nx = mp.sqrt(sigma + x[2] ** 2)
    y = [mp.sin(nd) / nd * d[0], mp.sin(nd) / nd * d[1], mp.cos(nd)]
    vy = sum(a * b for a, b in zip(v, y))
    return [nx * (y[i] - v[i] * beta * vy) for i in range(3)]


def perturbed(st, kind, index, comp, h):
    """Copy of st with one tangent coordinate moved by h."""
    out = dict(st)
    if kind == "knot":
        knots = [list(kn) for kn in st["knots"]]
        kn = knots[index]
        if comp < 3:
            d = [mp.mpf(0)] * 3; d[comp] = h
            kn[:4] = quat_mul(quat_exp(d), kn[:4])
        else:
            kn[4 + comp - 3] += h
        out["knots"] = knots
    elif kind == "landmark":
        lm = [list(v) for v in st["landmarks"]]
        lm[index][comp] += h
        out["landmarks"] = lm
    elif kind in ("bg", "ba"):
        b = [list(v) for v in st[kind]]
        b[index][comp] += h
        out[kind] = b
    elif kind == "gravity":
        d = [mp.mpf(0)] * 2; d[comp] = h
        out["gravity"] = sphere_plus(st["gravity"], d)
    return out


def numdiff(fn, st, f, kind, index, comp):
    rp = fn(perturbed(st, kind, index, comp, H_X), f)
    rm = fn(perturbed(st, kind, index, comp, -H_X), f)
    return [(a - b) / (2 * H_X) for a, b in zip(rp, rm)]


def to_mp_state(w):
    f = lambda a: [[mp.mpf(float(x)) for x in row] for row in np.asarray(a)]
    return dict(k=w.order, kb=w.bias_order, knots=f(w.knots), cams=f(w.cameras), imu=[mp.mpf(float(x)) for x in w.imu],
                landmarks=f(w.landmarks), bg=f(w.gyro_bias), ba=f(w.accel_bias), gravity=[mp.mpf(float(x)) for x in w.gravity])


def make_case(name, n_pix=3, n_imu=3, **kw):
    w = synthetic.make_window(**kw)
    st = to_mp_state(w)
    k, kb = w.order, w.bias_order
    stamps = [kn[7] for kn in st["knots"]]
    bst = [b[3] for b in st["bg"]]
    rng = np.random.default_rng(7)
    pix_idx = sorted(rng.choice(w.v_stamp.size, n_pix, replace=False).tolist()) if w.v_stamp.size else []
    imu_idx = sorted(rng.choice(w.i_stamp.size, n_imu, replace=False).tolist())
    case = dict(name=name, window=dict(order=k, bias_order=kb, knots=w.knots.tolist(), gyro_bias=w.gyro_bias.tolist(), accel_bias=w.accel_bias.tolist(),
                                       gravity=w.gravity.tolist(), cameras=w.cameras.tolist(), imu=w.imu.tolist(), landmarks=w.landmarks.tolist()),
                pixel=[], inertial=[], digits=mp.mp.dps)
    fl = lambda v: [float(x) for x in v]
    for f_ in pix_idx:
        f = dict(stamp=mp.mpf(float(w.v_stamp[f_])), cam=int(w.v_cam[f_]), lm=int(w.v_lm[f_]), pixel=[mp.mpf(float(x)) for x in w.v_pixel[f_]])
        base, _ = segment(stamps, k, f["stamp"])
        r = pixel_residual(st, f)
        Jp = [numdiff(pixel_residual, st, f, "knot", base + m, c) for m in range(k) for c in range(6)]   # columns
        Jl = [numdiff(pixel_residual, st, f, "landmark", f["lm"], c) for c in range(3)]
        case["pixel"].append(dict(stamp=float(w.v_stamp[f_]), cam=f["cam"], lm=f["lm"], pixel=w.v_pixel[f_].tolist(), base=base, r=fl(r),
                                  Jp=[[float(Jp[c][row]) for c in range(6 * k)] for row in range(2)],
                                  Jl=[[float(Jl[c][row]) for c in range(3)] for row in range(2)]))
        print(name, "pixel", f_, fl(r))
    for f_ in imu_idx:
        f = dict(stamp=mp.mpf(float(w.i_stamp[f_])), meas=[mp.mpf(float(x)) for x in w.i_meas[f_]])
        base, _ = segment(stamps, k, f["stamp"])
        gb, _ = segment(bst, kb, f["stamp"])
        r = inertial_residual(st, f)
        Jp = [numdiff(inertial_residual, st, f, "knot", base + m, c) for m in range(k) for c in range(6)]
        wg = [numdiff(inertial_residual, st, f, "bg", gb + m, 0)[0] for m in range(kb)]
        wa = [numdiff(inertial_residual, st, f, "ba", gb + m, 0)[3] for m in range(kb)]
        Jg = [numdiff(inertial_residual, st, f, "gravity", 0, c) for c in range(2)]
        case["inertial"].append(dict(stamp=float(w.i_stamp[f_]), meas=w.i_meas[f_].tolist(), base=base, bias_base=gb, r=fl(r),
                                     Jp=[[float(Jp[c][row]) for c in range(6 * k)] for row in range(6)], wg=fl(wg), wa=fl(wa),
                                     Jg=[[float(Jg[c][row]) for c in range(2)] for row in range(6)]))
        print(name, "inertial", f_, fl(r)[:3])
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), f"{name}.json")
    with open(out, "w") as fh:
        json.dump(case, fh)
    print("wrote", out)


def make_widened_case(name, n_bearing=3, n_pose=3, **kw):
    """Bearing + pose factors (the two other residual families of reference optimizer.cpp:189-251)."""
    w = synthetic.add_bearing_and_pose_factors(synthetic.make_window(**kw), num_bearing=8, num_pose=8, seed=kw["seed"] + 50)
    st = to_mp_state(w)
    st["pose_sensors"] = [[mp.mpf(float(x)) for x in row] for row in w.pose_sensors]
    k = w.order
    stamps = [kn[7] for kn in st["knots"]]
    case = dict(name=name, window=dict(order=k, bias_order=w.bias_order, knots=w.knots.tolist(), gyro_bias=w.gyro_bias.tolist(), accel_bias=w.accel_bias.tolist(),
                                       gravity=w.gravity.tolist(), cameras=w.cameras.tolist(), imu=w.imu.tolist(), landmarks=w.landmarks.tolist(),
                                       pose_sensors=w.pose_sensors.tolist()),
                pixel=[], inertial=[], bearing=[], pose=[], digits=mp.mp.dps)
    fl = lambda v: [float(x) for x in v]
    for f_ in range(n_bearing):
        f = dict(stamp=mp.mpf(float(w.b_stamp[f_])), cam=int(w.b_cam[f_]), lm=int(w.b_lm[f_]), bearing=[mp.mpf(float(x)) for x in w.b_bearing[f_]])
        base, _ = segment(stamps, k, f["stamp"])
        r = bearing_residual(st, f)
        Jp = [numdiff(bearing_residual, st, f, "knot", base + m, c) for m in range(k) for c in range(6)]
        Jl = [numdiff(bearing_residual, st, f, "landmark", f["lm"], c) for c in range(3)]
        case["bearing"].append(dict(stamp=float(w.b_stamp[f_]), cam=f["cam"], lm=f["lm"], bearing=w.b_bearing[f_].tolist(), base=base, r=fl(r),
                                    Jp=[float(Jp[c][0]) for c in range(6 * k)], Jl=[float(Jl[c][0]) for c in range(3)]))
        print(name, "bearing", f_, fl(r))
    for f_ in range(n_pose):
        f = dict(stamp=mp.mpf(float(w.m_stamp[f_])), sensor=int(w.m_sensor[f_]), pose=[mp.mpf(float(x)) for x in w.m_pose[f_]])
        base, _ = segment(stamps, k, f["stamp"])
        r = pose_residual(st, f)
        Jp = [numdiff(pose_residual, st, f, "knot", base + m, c) for m in range(k) for c in range(6)]
        case["pose"].append(dict(stamp=float(w.m_stamp[f_]), sensor=f["sensor"], pose=w.m_pose[f_].tolist(), base=base, r=fl(r),
                                 Jp=[[float(Jp[c][row]) for c in range(6 * k)] for row in range(6)]))
        print(name, "pose", f_, fl(r)[:3])
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), f"{name}.json")
    with open(out, "w") as fh:
        json.dump(case, fh)
    print("wrote", out)


if __name__ == "__main__":
    if "--widened-only" in sys.argv:
        make_widened_case("k4_bearing_pose", order=4, num_knots=10, num_landmarks=12, num_imu=0, seed=synthetic.SEED_BASE + 905)
        make_widened_case("k6_bearing_pose", order=6, num_knots=12, num_landmarks=12, num_imu=0, seed=synthetic.SEED_BASE + 906, n_bearing=2, n_pose=2)
        sys.exit(0)
    make_case("k4_euroc", order=4, num_knots=10, num_landmarks=12, num_imu=20, seed=synthetic.SEED_BASE + 901)
    make_case("k6_euroc", order=6, num_knots=12, num_landmarks=12, num_imu=20, seed=synthetic.SEED_BASE + 902)
    make_case("k4_generic_calibration", order=4, num_knots=10, num_landmarks=12, num_imu=20, generic_calibration=True, seed=synthetic.SEED_BASE + 903)
    make_case("k6_generic_calibration", order=6, num_knots=12, num_landmarks=12, num_imu=20, generic_calibration=True, seed=synthetic.SEED_BASE + 904, n_pix=2, n_imu=2)
    make_widened_case("k4_bearing_pose", order=4, num_knots=10, num_landmarks=12, num_imu=0, seed=synthetic.SEED_BASE + 905)
    make_widened_case("k6_bearing_pose", order=6, num_knots=12, num_landmarks=12, num_imu=0, seed=synthetic.SEED_BASE + 906, n_bearing=2, n_pose=2)
