"""Calibration-block Jacobians and the HYPER_REFERENCE_QUIRKS switch on the device.

The reference's gradient tests run with EVERY sensor manifold non-constant (reference
tests/internal/tests/optimizers/evaluators/pixel.cpp:57, inertial.cpp:56-61): extrinsics, intrinsics, distortion,
i_g, i_a, S_g, X_a all get analytic-vs-numeric checks.  Here the CUDA path's Ceres-shaped copy-out
(hb200_factor_evaluate) is compared block by block with the oracle's restatement of
pixel.cpp:91-141 / inertial.cpp:155-194 / bearing.cpp:76 / manifold.cpp:57, in the consistent model and with the
in-tree formulas verbatim (quirks = 15), and the reference's own Probe protocol (evaluator.hpp:26-65: central
differences through Manifold::Plus, step 1e-6, relative OR normalised-absolute error <= 1e-5) is run directly
against the GPU residuals for the calibration blocks.
"""
import numpy as np
import pytest

import oracle_lib as ol
from hyperslam_b200 import runtime, synthetic

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / (np.abs(np.asarray(b)).max() + 1e-300))


def make_window(order, seed_off=0):
    win = synthetic.make_window(order=order, num_knots=16, num_landmarks=40, num_imu=60, generic_calibration=True,
                                seed=synthetic.SEED_BASE + 900 + seed_off)
    return synthetic.add_bearing_and_pose_factors(win, num_bearing=30, num_pose=8, seed=synthetic.SEED_BASE + 901 + seed_off)


def make_ctx(win, quirks=0):
    ctx = runtime.Context(0)
    ctx.load_window(win)
    ctx.set_reference_quirks(quirks)
    ctx.evaluate()
    return ctx


def blocks_of(win, kind, f, maps):
    vb, ib, ig, ia = maps
    k, kb = win.order, win.bias_order
    if kind == runtime.PIXEL:
        cam = win.cameras[win.v_cam[f]]
        return [win.knots[vb[f] + m] for m in range(k)] + [cam[:7], cam[7:11], cam[11:15], win.landmarks[win.v_lm[f]]], [ol.M_SE3, ol.M_EUCLIDEAN, ol.M_EUCLIDEAN]
    if kind == runtime.INERTIAL:
        imu = win.imu
        return ([win.knots[ib[f] + m] for m in range(k)] + [imu[:7], imu[7:13], imu[13:19], imu[19:28], imu[28:37]]
                + [win.gyro_bias[ig[f] + m] for m in range(kb)] + [win.accel_bias[ia[f] + m] for m in range(kb)] + [win.gravity]), \
            [ol.M_SE3] + [ol.M_EUCLIDEAN] * 4
    raise ValueError(kind)


def knot_base(win, stamp):
    st = win.knots[:, 7]
    j = int(np.searchsorted(st, stamp, side="right") - 1)
    return j - (win.order - 1) // 2


@pytest.mark.parametrize("quirks", [0, 15, 1, 2, 4, 8])
@pytest.mark.parametrize("order", [4, 6])
def test_calibration_blocks_match_oracle(built, order, quirks):
    win = make_window(order)
    ctx = make_ctx(win, quirks)
    maps = ctx.index_maps()
    k = win.order
    for f in (0, 5, win.v_stamp.size - 1):
        blocks, mids = blocks_of(win, runtime.PIXEL, f, maps)
        r, jac = ctx.factor_evaluate(runtime.PIXEL, f, blocks)
        r_o, jac_o = ol.cost_evaluate(ol.PIXEL, win.v_stamp[f], win.v_pixel[f], np.concatenate(blocks), k=k, quirks=quirks)
        assert rel_err(r, r_o) < 1e-9
        PJ = ol.manifold_plus_jacobian(ol.M_SE3, blocks[k])
        assert rel_err(jac[k] @ PJ, jac_o[k] @ PJ) < 1e-8          # extrinsics (pixel.cpp:141)
        assert rel_err(jac[k + 1], jac_o[k + 1]) < 1e-9            # intrinsics (pixel.cpp:99-102)
        assert rel_err(jac[k + 2], jac_o[k + 2]) < 1e-9            # distortion (pixel.cpp:95,114)
    for f in (0, 7, win.i_stamp.size - 1):
        blocks, mids = blocks_of(win, runtime.INERTIAL, f, maps)
        r, jac = ctx.factor_evaluate(runtime.INERTIAL, f, blocks)
        r_o, jac_o = ol.cost_evaluate(ol.INERTIAL, win.i_stamp[f], win.i_meas[f], np.concatenate(blocks), k=k, quirks=quirks)
        assert rel_err(r, r_o) < 1e-9
        for m in range(k):   # the control-point blocks follow the selected variant too
            PJ = ol.manifold_plus_jacobian(ol.M_STATE, blocks[m])
            assert rel_err(jac[m] @ PJ, jac_o[m] @ PJ) < 1e-8
        PJ = ol.manifold_plus_jacobian(ol.M_SE3, blocks[k])
        assert rel_err(jac[k] @ PJ, jac_o[k] @ PJ) < 1e-8          # extrinsics (inertial.cpp:155-162)
        for b in range(1, 5):                                       # i_g, i_a, S_g, X_a (inertial.cpp:164-194)
            scale = np.abs(jac_o[k + b]).max()
            assert np.abs(jac[k + b] - jac_o[k + b]).max() <= 1e-9 * max(scale, 1.0), (b, quirks)
        PJ = ol.manifold_plus_jacobian(ol.M_SPHERE, win.gravity)
        assert rel_err(jac[-1] @ PJ, jac_o[-1] @ PJ) < 1e-8
    # bearing (bearing.cpp:76) and manifold (manifold.cpp:57) extrinsics
    for f in (0, win.b_stamp.size - 1):
        cam = win.cameras[win.b_cam[f]]
        base = knot_base(win, win.b_stamp[f])
        blocks = [win.knots[base + m] for m in range(k)] + [cam[:7], cam[7:11], cam[11:15], win.landmarks[win.b_lm[f]]]
        r, jac = ctx.factor_evaluate(runtime.BEARING, f, blocks)
        r_o, jac_o = ol.cost_evaluate(ol.BEARING, win.b_stamp[f], win.b_bearing[f], np.concatenate(blocks), k=k)
        PJ = ol.manifold_plus_jacobian(ol.M_SE3, blocks[k])
        assert rel_err(r, r_o) < 1e-9
        assert rel_err(jac[k] @ PJ, jac_o[k] @ PJ) < 1e-8
        assert np.abs(jac[k + 1]).max() == 0.0 and np.abs(jac[k + 2]).max() == 0.0   # the bearing evaluator has no intrinsics / distortion block
    for f in (0, win.m_stamp.size - 1):
        base = knot_base(win, win.m_stamp[f])
        blocks = [win.knots[base + m] for m in range(k)] + [win.pose_sensors[win.m_sensor[f]]]
        r, jac = ctx.factor_evaluate(runtime.MANIFOLD, f, blocks)
        r_o, jac_o = ol.cost_evaluate(ol.MANIFOLD, win.m_stamp[f], win.m_pose[f], np.concatenate(blocks), k=k)
        PJ = ol.manifold_plus_jacobian(ol.M_SE3, blocks[k])
        assert rel_err(r, r_o) < 1e-9
        assert rel_err(jac[k] @ PJ, jac_o[k] @ PJ) < 1e-8
    ctx.close()


@pytest.mark.parametrize("order", [4, 6])
def test_quirk_window_outputs_match_oracle_quirk_mode(built, order):
    """Whole-window compact outputs (what the iteration consumes) under quirks = 15 vs the oracle's verbatim mode."""
    win = synthetic.make_window(order=order, num_knots=16, num_landmarks=40, num_imu=200, generic_calibration=True, seed=synthetic.SEED_BASE + 910)
    ref = ol.OracleWindow(win, quirks=15).evaluate()
    ctx = runtime.Context(0)
    ctx.load_window(win)
    ctx.set_reference_quirks(15)
    ctx.evaluate()
    got = ctx.outputs()
    for key in ("i_r", "i_Jp", "i_Jg", "i_wg", "i_wa", "v_r", "v_Jp", "v_Jl"):
        assert rel_err(got[key], ref[key]) < 1e-8, key
    # and the two variants really differ on a general calibration (the switch is live)
    ctx.set_reference_quirks(0)
    ctx.evaluate()
    assert rel_err(ctx.outputs()["i_Jp"], ref["i_Jp"]) > 1e-4
    ctx.close()


def test_quirk_variants_coincide_on_reference_calibration(built):
    """I_g = I_a = I, S_g = X_a = 0, R_bs = I (reference settings.yaml:83-106): every variant gives the same Jacobians."""
    win = synthetic.make_window(order=4, num_knots=16, num_landmarks=20, num_imu=200, seed=synthetic.SEED_BASE + 911)
    ctx = runtime.Context(0)
    ctx.load_window(win)
    ctx.evaluate()
    base = ctx.outputs()
    for q in (1, 2, 8, 11):   # (4 only touches the extrinsics block)
        ctx.set_reference_quirks(q)
        ctx.evaluate()
        got = ctx.outputs()
        for key in ("i_r", "i_Jp", "i_Jg"):
            assert rel_err(got[key], base[key]) < 1e-12, (q, key)
    ctx.close()


def _probe_block(residual_fn, x, mid, analytic, nr):
    """Reference Probe (evaluator.hpp:38-65) of one block: residual_fn(x_block) -> residual vector."""
    PJ = ol.manifold_plus_jacobian(mid, x)
    Ja = analytic @ PJ
    nt = PJ.shape[1]
    Jn = np.zeros((nr, nt))
    h = 1e-6
    for c in range(nt):
        d = np.zeros(nt); d[c] = h
        rp = residual_fn(ol.manifold_plus(mid, x, d))
        rm = residual_fn(ol.manifold_plus(mid, x, -d))
        Jn[:, c] = (rp - rm) / (2 * h)
    ae = np.abs(Ja - Jn)
    den = np.maximum(np.abs(Ja), np.abs(Jn))
    rel = np.where((Ja == 0) | (Jn == 0), ae, ae / np.where(den == 0, 1, den)).max()
    na, nn = np.linalg.norm(Ja), np.linalg.norm(Jn)
    ab = np.abs(Ja / (na if na > 0 else 1) - Jn / (nn if nn > 0 else 1)).max()
    return rel, ab


@pytest.mark.parametrize("order", [4, 6])
def test_gpu_calibration_jacobians_pass_reference_probe(built, order):
    """The reference's own gradient protocol, with the GPU as the function under test: perturb a calibration block
    through its manifold, re-upload, re-evaluate the residuals on the device, compare with the device's analytic
    block.  Consistent model (quirks = 0): the verbatim extrinsics block of inertial.cpp:157-158 does not pass its
    own probe for a rotated IMU (tests/test_oracle.py::test_reference_quirks_agree_on_reference_fixtures)."""
    win = make_window(order, seed_off=7)
    ctx = make_ctx(win)
    maps = ctx.index_maps()
    k = win.order
    tol = 1e-5

    f = 3
    blocks, mids = blocks_of(win, runtime.PIXEL, f, maps)
    _, jac = ctx.factor_evaluate(runtime.PIXEL, f, blocks)
    ci = int(win.v_cam[f])
    for b, (lo, hi) in enumerate([(0, 7), (7, 11), (11, 15)]):
        def residual(xb, lo=lo, hi=hi):
            cams = win.cameras.copy()
            cams[ci, lo:hi] = xb
            ctx.set_cameras(cams)
            ctx.evaluate()   # (the cost-only sweep does not materialise residuals)
            return ctx.outputs(jacobians=False)["v_r"][f].copy()
        rel, ab = _probe_block(residual, win.cameras[ci, lo:hi].copy(), mids[b], jac[k + b], 2)
        assert rel <= tol or ab <= tol, ("pixel", b, rel, ab)
    ctx.set_cameras(win.cameras)

    f = 11
    blocks, mids = blocks_of(win, runtime.INERTIAL, f, maps)
    ctx.evaluate()
    _, jac = ctx.factor_evaluate(runtime.INERTIAL, f, blocks)
    for b, (lo, hi) in enumerate([(0, 7), (7, 13), (13, 19), (19, 28), (28, 37)]):
        def residual(xb, lo=lo, hi=hi):
            imu = win.imu.copy()
            imu[lo:hi] = xb
            ctx.set_imu(imu)
            ctx.evaluate()
            return ctx.outputs(jacobians=False)["i_r"][f].copy()
        rel, ab = _probe_block(residual, win.imu[lo:hi].copy(), mids[b], jac[k + b], 6)
        assert rel <= tol or ab <= tol, ("inertial", b, rel, ab)
    ctx.close()
