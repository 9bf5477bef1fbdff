"""ctypes binding of the CPU oracle (oracle/libhyper_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (hyperslam_b200/) never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "libhyper_oracle.so")

PIXEL, INERTIAL, BEARING, MANIFOLD = 0, 1, 2, 3
NUM_RESIDUALS = {0: 2, 1: 6, 2: 1, 3: 6}
M_STATE, M_SE3, M_EUCLIDEAN, M_CONSTANT, M_BIAS, M_SPHERE = range(6)
QUIRKS_ALL = 15

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(_ROOT, "oracle")])


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.ho_window_create.restype = C.c_void_p
        _lib.ho_window_cost.restype = C.c_double
        _lib.ho_window_cost.argtypes = [C.c_void_p]
        _lib.ho_window_destroy.argtypes = [C.c_void_p]
    return _lib


def basis(k):
    M = np.zeros((k, k))
    assert lib().ho_basis(k, _d(M)) == 0
    return M


def basis_eval(k, u, inv_dt):
    lam = np.zeros((3, k))
    lib().ho_basis_eval(k, C.c_double(u), C.c_double(inv_dt), _d(lam))
    return lam


def state_evaluate(cps, stamp, derivative=2, jac=True):
    cps = np.ascontiguousarray(cps, dtype=np.float64)
    k = cps.shape[0]
    value, vel, acc = np.zeros(7), np.zeros(6), np.zeros(6)
    J = np.zeros((3, 6, 8 * k))
    lib().ho_state_evaluate(k, _d(cps), C.c_double(stamp), derivative, int(jac), _d(value), _d(vel), _d(acc), _d(J))
    return value, vel, acc, J


def layout(kind, k=4, k_bg=4, k_ba=4):
    out = np.zeros(6, dtype=np.int32)
    offs = np.zeros(32, dtype=np.int32)
    sizes = np.zeros(32, dtype=np.int32)
    lib().ho_layout(kind, k, k_bg, k_ba, _i(out), _i(offs), _i(sizes))
    nb = int(out[0])
    return dict(num_blocks=nb, num_parameters=int(out[1]), state_idx=int(out[2]), sensor_static_idx=int(out[3]),
                sensor_dynamic_idx=int(out[4]), observation_idx=int(out[5]), offsets=offs[:nb].copy(), sizes=sizes[:nb].copy())


def cost_evaluate(kind, stamp, meas, params, k=4, k_bg=4, k_ba=4, jac=True, jac_mask=None, quirks=0):
    """Per-factor reference-shaped Evaluate.  Returns (residuals, [per-block row-major jacobians])."""
    L = layout(kind, k, k_bg, k_ba)
    nr = NUM_RESIDUALS[kind]
    params = np.ascontiguousarray(params, dtype=np.float64)
    assert params.size == L["num_parameters"]
    meas = np.ascontiguousarray(meas, dtype=np.float64)
    r = np.zeros(nr)
    J = np.zeros(nr * L["num_parameters"]) if jac else None
    mask = None if jac_mask is None else np.ascontiguousarray(jac_mask, dtype=np.int32)
    rc = lib().ho_cost_evaluate(kind, C.c_double(stamp), _d(meas), k, k_bg, k_ba, _d(params), _i(mask), _d(r), _d(J), quirks)
    assert rc == 0
    if not jac:
        return r, None
    blocks = []
    for b in range(L["num_blocks"]):
        o, s = int(L["offsets"][b]), int(L["sizes"][b])
        blocks.append(J[nr * o: nr * (o + s)].reshape(nr, s).copy())
    return r, blocks


def probe(kind, stamp, meas, params, manifold_ids, k=4, k_bg=4, k_ba=4, tol=1e-5, quirks=0, richardson=False):
    L = layout(kind, k, k_bg, k_ba)
    params = np.ascontiguousarray(params, dtype=np.float64)
    meas = np.ascontiguousarray(meas, dtype=np.float64)
    ids = np.ascontiguousarray(manifold_ids, dtype=np.int32)
    res = np.zeros(4)
    per = np.zeros((L["num_blocks"], 2))
    rc = lib().ho_probe(kind, C.c_double(stamp), _d(meas), k, k_bg, k_ba, _d(params), _i(ids), C.c_double(tol), quirks, _d(res), _d(per), int(richardson))
    return rc == 0, res, per


def manifold_plus(mid, x, delta):
    x = np.ascontiguousarray(x, dtype=np.float64)
    delta = np.ascontiguousarray(delta, dtype=np.float64)
    out = np.zeros_like(x)
    lib().ho_manifold_plus(mid, x.size, _d(x), _d(delta), _d(out))
    return out


def manifold_minus(mid, y, x):
    y = np.ascontiguousarray(y, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    a, t = C.c_int(), C.c_int()
    lib().ho_manifold_sizes(mid, x.size, C.byref(a), C.byref(t))
    out = np.zeros(t.value)
    lib().ho_manifold_minus(mid, x.size, _d(y), _d(x), _d(out))
    return out


def manifold_plus_jacobian(mid, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    a, t = C.c_int(), C.c_int()
    lib().ho_manifold_sizes(mid, x.size, C.byref(a), C.byref(t))
    J = np.zeros((a.value, t.value))
    lib().ho_manifold_plus_jacobian(mid, x.size, _d(x), _d(J))
    return J


class OracleWindow:
    """Flat sliding-window problem evaluated by the oracle (mirrors hyperslam_b200.Window)."""

    def __init__(self, win, quirks=0, radius=1e4):
        L = lib()
        self.h = C.c_void_p(L.ho_window_create())
        self.win = win
        w = win
        L.ho_window_set_spline(self.h, w.order, w.knots.shape[0], _d(w.knots))
        L.ho_window_set_bias(self.h, w.bias_order, w.gyro_bias.shape[0], _d(w.gyro_bias), w.accel_bias.shape[0], _d(w.accel_bias))
        L.ho_window_set_gravity(self.h, _d(w.gravity))
        L.ho_window_set_cameras(self.h, w.cameras.shape[0], _d(w.cameras))
        L.ho_window_set_imu(self.h, _d(w.imu))
        L.ho_window_set_landmarks(self.h, w.landmarks.shape[0], _d(w.landmarks))
        L.ho_window_set_pixel_factors(self.h, w.v_stamp.size, _d(w.v_stamp), _i(w.v_cam), _i(w.v_lm), _d(w.v_pixel))
        L.ho_window_set_inertial_factors(self.h, w.i_stamp.size, _d(w.i_stamp), _d(w.i_meas))
        if w.b_stamp.size:
            L.ho_window_set_bearing_factors(self.h, w.b_stamp.size, _d(w.b_stamp), _i(w.b_cam), _i(w.b_lm), _d(w.b_bearing))
            L.ho_window_set_bearing_loss(self.h, C.c_double(w.huber_bearing))
        if w.m_stamp.size:
            L.ho_window_set_pose_sensors(self.h, w.pose_sensors.shape[0], _d(w.pose_sensors))
            L.ho_window_set_manifold_factors(self.h, w.m_stamp.size, _d(w.m_stamp), _i(w.m_sensor), _d(w.m_pose))
        kc = np.ascontiguousarray(w.knot_const, dtype=np.uint8)
        L.ho_window_set_constant(self.h, kc.ctypes.data_as(C.POINTER(C.c_ubyte)), int(w.gravity_const), int(w.bias_const))
        L.ho_window_set_options(self.h, C.c_double(w.huber_pixel), C.c_double(w.imu_loss_scale), quirks, C.c_double(radius))
        self.bad = L.ho_window_bind(self.h)
        self.n = L.ho_window_reduced_size(self.h)

    def __del__(self):
        try:
            lib().ho_window_destroy(self.h)
        except Exception:
            pass

    def index_maps(self):
        w = self.win
        vb = np.zeros(w.v_stamp.size, dtype=np.int32)
        ib = np.zeros(w.i_stamp.size, dtype=np.int32)
        ig = np.zeros(w.i_stamp.size, dtype=np.int32)
        ia = np.zeros(w.i_stamp.size, dtype=np.int32)
        lib().ho_window_get_index_maps(self.h, _i(vb), _i(ib), _i(ig), _i(ia))
        return vb, ib, ig, ia

    def evaluate(self, want_J=True, nthreads=0, v_count=-1, i_count=-1, outputs=True):
        w = self.win
        k, kb = w.order, w.bias_order
        np_, nb, nm = w.v_stamp.size, w.b_stamp.size, w.m_stamp.size
        nv, ni = np_ + nb, w.i_stamp.size     # visual list = pixel factors followed by bearing factors
        if not outputs:
            lib().ho_window_evaluate(self.h, int(want_J), None, None, None, None, None, None, None, None, nthreads, v_count, i_count)
            return None
        out = dict(v_r=np.zeros((nv, 2)), v_Jp=np.zeros((nv, 2, 6 * k)), v_Jl=np.zeros((nv, 2, 3)), i_r=np.zeros((ni, 6)),
                   i_Jp=np.zeros((ni, 6, 6 * k)), i_wg=np.zeros((ni, kb)), i_wa=np.zeros((ni, kb)), i_Jg=np.zeros((ni, 6, 2)))
        lib().ho_window_evaluate(self.h, int(want_J), _d(out["v_r"]), _d(out["v_Jp"]), _d(out["v_Jl"]), _d(out["i_r"]), _d(out["i_Jp"]),
                                 _d(out["i_wg"]), _d(out["i_wa"]), _d(out["i_Jg"]), nthreads, v_count, i_count)
        if nb:
            assert np.all(out["v_r"][np_:, 1] == 0) and np.all(out["v_Jp"][np_:, 1] == 0) and np.all(out["v_Jl"][np_:, 1] == 0)
            out.update(b_r=out["v_r"][np_:, 0].copy(), b_Jp=out["v_Jp"][np_:, 0].copy(), b_Jl=out["v_Jl"][np_:, 0].copy())
            out.update(v_r=out["v_r"][:np_], v_Jp=out["v_Jp"][:np_], v_Jl=out["v_Jl"][:np_])
        if nm:
            out.update(m_r=np.zeros((nm, 6)), m_Jp=np.zeros((nm, 6, 6 * k)))
            lib().ho_window_evaluate_manifold(self.h, int(want_J), _d(out["m_r"]), _d(out["m_Jp"]))
        return out

    def cost(self):
        return lib().ho_window_cost(self.h)

    def iterate(self, apply=True, nthreads=0, outputs=True):
        n, L = self.n, self.win.landmarks.shape[0]
        stats = np.zeros(8)
        if not outputs:
            lib().ho_window_iterate(self.h, int(apply), None, None, None, None, _d(stats), nthreads)
            return dict(stats=stats)
        S, b, dp, dl = np.zeros((n, n)), np.zeros(n), np.zeros(n), np.zeros((L, 3))
        lib().ho_window_iterate(self.h, int(apply), _d(S), _d(b), _d(dp), _d(dl), _d(stats), nthreads)
        return dict(S=S, b=b, delta_p=dp, delta_l=dl, cost=stats[0], cost_new=stats[1], model_change=stats[2], rho=stats[3],
                    radius=stats[4], accepted=int(stats[5]), spd=int(stats[6]), stats=stats)

    def optimize(self, max_iter, function_tolerance=1e-6, gradient_tolerance=1e-10, parameter_tolerance=1e-8, min_radius=1e-32, nthreads=0):
        """ceres::Solve restated: iterations until one of Ceres' termination tests fires (or max_iter)."""
        stats, per = np.zeros(6), np.zeros((max_iter, 4))
        n = lib().ho_window_optimize(self.h, int(max_iter), C.c_double(function_tolerance), C.c_double(gradient_tolerance), C.c_double(parameter_tolerance),
                                     C.c_double(min_radius), _d(stats), _d(per), nthreads)
        return dict(type=int(stats[0]), iterations=int(n), gradient_max_norm=stats[2], step_norm=stats[3], x_norm=stats[4], radius=stats[5],
                    cost=per[:n, 0].copy(), cost_new=per[:n, 1].copy(), accepted=per[:n, 2].astype(int), radii=per[:n, 3].copy())

    def build_packed(self):
        n = self.n
        packed = np.zeros(n * n + 3 * n + 2)
        lib().ho_window_build_packed(self.h, _d(packed))
        return packed

    def finalize_packed(self, packed):
        packed = np.ascontiguousarray(packed, dtype=np.float64).copy()
        lib().ho_window_finalize_packed(self.h, _d(packed))
        n = self.n
        return packed[: n * n].reshape(n, n), packed[n * n: n * n + n]

    def state(self):
        w = self.win
        knots, bg, ba = np.zeros_like(w.knots), np.zeros_like(w.gyro_bias), np.zeros_like(w.accel_bias)
        g, lm = np.zeros(3), np.zeros_like(w.landmarks)
        lib().ho_window_get_state(self.h, _d(knots), _d(bg), _d(ba), _d(g), _d(lm))
        return dict(knots=knots, gyro_bias=bg, accel_bias=ba, gravity=g, landmarks=lm)


def ingest_stereo_frame(cps, stamp, cam0, cam1, px0, px1):
    """Oracle restatement of the stereo-frame ingest (oracle/ho_ingest.h): returns (B0, B1, landmarks)."""
    cps = np.ascontiguousarray(cps, dtype=np.float64)
    px0 = np.ascontiguousarray(px0, dtype=np.float64); px1 = np.ascontiguousarray(px1, dtype=np.float64)
    n = px0.shape[0]
    B0, B1, L = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3))
    rc = lib().ho_ingest_stereo_frame(cps.shape[0], _d(cps), C.c_double(stamp), _d(np.ascontiguousarray(cam0, dtype=np.float64)),
                                      _d(np.ascontiguousarray(cam1, dtype=np.float64)), n, _d(px0), _d(px1), _d(B0), _d(B1), _d(L))
    assert rc == 0
    return B0, B1, L


def num_threads():
    return lib().ho_num_threads()
