"""Pins the CPU oracle (oracle/, test infrastructure) against everything the reference offers for this path:
independent OpenCV projections (golden fixture), the constants and tolerances of the reference's own unit tests, the Ceres
iteration log stored in its demo notebook, and the acceptance criteria of its integration test. Runs without a GPU."""
import json
import os

import numpy as np
import pytest

from calico_b200 import spline as sp
from calico_b200 import synthetic

HERE = os.path.dirname(os.path.abspath(__file__))


def test_camera_models_match_opencv_golden(oracle):
    """OpenCv5 / OpenCv8 / KannalaBrandt ProjectPoint (camera_models.h:105-141, 257-298, 420-462) against cv2.projectPoints /
    cv2.fisheye.projectPoints outputs committed in tests/golden/cv2_projection.json (generator beside it)."""
    with open(os.path.join(HERE, "golden", "cv2_projection.json")) as f:
        g = json.load(f)
    pts = np.array(g["points"])
    for model, d in g["models"].items():
        intr = np.array(d["intrinsics"])
        want = np.array(d["pixels"])
        for p, w in zip(pts, want):
            ok, px = oracle.project_point(int(model), intr, p)
            assert ok
            np.testing.assert_allclose(px, w, rtol=0, atol=1e-9)


def test_camera_models_reject_points_like_reference(oracle):
    """camera_test.cpp:113-237 behaviour: a point behind the camera does not project (z <= 0, camera_models.h:108-111)."""
    for model, intr in synthetic.CAMERA_TRUTH.items():
        ok, _ = oracle.project_point(model, intr, np.array([0.1, 0.1, 1.0]))
        assert ok
        if model in (1, 2, 3, 5):
            ok, _ = oracle.project_point(model, intr, np.array([0.1, 0.1, -1.0]))
            assert not ok


def test_imu_models_known_answers(oracle):
    """accelerometer_models.h:80-85,129-141,208-235: s*w, s*w+b, S*A*w+b with the constants of batch_optimizer_test.cpp:90."""
    w = np.array([0.3, -0.2, 0.5])
    ok, out = oracle.imu_project(1, np.array([1.3]), w)
    np.testing.assert_allclose(out, 1.3 * w)
    ok, out = oracle.imu_project(2, np.array([1.3, 0.01, -0.01, 0.01]), w)
    np.testing.assert_allclose(out, 1.3 * w + np.array([0.01, -0.01, 0.01]))
    p = synthetic.IMU_MODEL_TRUTH[3]
    S = np.diag(p[:3])
    A = np.array([[1, p[3], p[4]], [p[5], 1, p[6]], [p[7], p[8], 1]])
    ok, out = oracle.imu_project(3, p, w)
    np.testing.assert_allclose(out, S @ A @ w + p[9:12])


def test_uniform_basis_matrix(oracle):
    """bspline.hpp:192-244 on uniform knots gives the order-6 uniform B-spline matrix."""
    want = np.array([[1, 26, 66, 26, 1, 0], [-5, -50, 0, 50, 5, 0], [10, 20, -60, 20, 10, 0], [-10, 20, 0, -20, 10, 0],
                     [5, -20, 30, -20, 5, 0], [-1, 5, -10, 10, -5, 1]]) / 120.0
    knots = np.arange(24, dtype=np.float64) * 0.1
    np.testing.assert_allclose(oracle.basis_matrix(6, knots, 12), want, atol=1e-12)
    np.testing.assert_allclose(sp.basis_matrix(knots, 6, 12), want, atol=1e-12)


def _bspline_fixture():
    t = 0.1 * np.arange(101)
    data = np.stack([np.cos(t), np.sin(1.5 * t), t * np.cos(t)], axis=1)
    return t, data


def test_spline_interpolation_precision(oracle):
    """bspline_test.cpp:52-94 InterpolationPrecision3DOF: order 6, 5 Hz knots fitted to (cos t, sin 1.5t, t cos t); derivatives
    0..3 within 1e-6 / 1e-5 / 1e-4 / 1e-2 of the analytic values, evaluated by the oracle's BSpline::Evaluate restatement."""
    t, data = _bspline_fixture()
    data6 = np.concatenate([data, np.zeros_like(data)], axis=1)
    from oracle import spline_fit as ofit
    knots, _, _, ctrl = ofit.fit_spline(t, data6, 6, 5.0)     # the oracle's restatement of BSpline::FitSpline (dense normal equations)
    spl = sp.Spline(6, knots, ctrl)
    api = oracle.oracle_api()
    api.set_trajectory(6, spl.knots, spl.ctrl)
    ti = (t[-1] - t[0]) / 201 * np.arange(201)
    want = [np.stack([np.cos(ti), np.sin(1.5 * ti), ti * np.cos(ti)], 1),
            np.stack([-np.sin(ti), 1.5 * np.cos(1.5 * ti), np.cos(ti) - ti * np.sin(ti)], 1),
            np.stack([-np.cos(ti), -2.25 * np.sin(1.5 * ti), -2 * np.sin(ti) - ti * np.cos(ti)], 1),
            np.stack([np.sin(ti), -3.375 * np.cos(1.5 * ti), ti * np.sin(ti) - 3 * np.cos(ti)], 1)]
    for d, tol in zip(range(4), (1e-6, 1e-5, 1e-4, 1e-2)):
        got = oracle.spline_interpolate(api, ti, d)[:, :3]
        assert np.abs(got - want[d]).max() < tol
        np.testing.assert_allclose(spl.evaluate(ti, d)[:, :3], got, atol=1e-9)   # host-side spline agrees with the oracle


def test_spline_invalid_arguments(oracle):
    """bspline_test.cpp:34-50: derivative -1 or >= order, or a time outside the valid knots -> kInvalidArgument (3)."""
    from calico_b200 import _capi
    t, data = _bspline_fixture()
    from oracle import spline_fit as ofit
    knots, _, _, ctrl = ofit.fit_spline(t, np.concatenate([data, data], axis=1), 6, 5.0)
    spl = sp.Spline(6, knots, ctrl)
    api = oracle.oracle_api()
    api.set_trajectory(6, spl.knots, spl.ctrl)
    for bad in (-1, 6):
        with pytest.raises(_capi.CalicoError) as e:
            oracle.spline_interpolate(api, [0.0], bad)
        assert e.value.code == _capi.INVALID_ARGUMENT
    with pytest.raises(_capi.CalicoError) as e:
        oracle.spline_interpolate(api, [-1.0], 0)
    assert e.value.code == _capi.INVALID_ARGUMENT


def _rot(oracle, phi):
    return oracle.exp_so3(np.asarray(phi, dtype=float))


def test_exp_so3_jacobian_against_finite_differences(oracle):
    """geometry_test.cpp:44-161: ExpSO3Jacobian maps d(phi) to the left-perturbation of the rotation:
    R(phi + d) R(phi)^T ~ Exp(J d)."""
    rng = np.random.default_rng(0)
    for _ in range(20):
        phi = rng.uniform(-2.5, 2.5, 3)
        J = oracle.exp_so3_jacobian(phi)
        R = _rot(oracle, phi)
        for k in range(3):
            d = np.zeros(3); d[k] = 1e-6
            dR = (_rot(oracle, phi + d) @ R.T - _rot(oracle, phi - d) @ R.T) / 2e-6
            w = np.array([dR[2, 1] - dR[1, 2], dR[0, 2] - dR[2, 0], dR[1, 0] - dR[0, 1]]) / 2
            np.testing.assert_allclose(w, J[:, k], atol=1e-8)
    np.testing.assert_allclose(oracle.exp_so3_jacobian(np.zeros(3)), np.eye(3))


def test_quaternion_conventions(oracle):
    """ceres::AngleAxisToQuaternion and EigenQuaternionManifold::Plus (Ceres external): x,y,z,w storage, left-multiplicative
    plus with a half-angle tangent (SURVEY §8 trap 5)."""
    q = oracle.angle_axis_to_quaternion(np.array([0.0, 0.0, np.pi / 2]))
    np.testing.assert_allclose(q, [0, 0, np.sin(np.pi / 4), np.cos(np.pi / 4)], atol=1e-15)
    q0 = np.array([0.0, 0.0, 0.0, 1.0])
    d = np.array([0.1, 0.0, 0.0])
    np.testing.assert_allclose(oracle.quaternion_plus(q0, d), [np.sin(0.1), 0, 0, np.cos(0.1)], atol=1e-15)
    qa = oracle.angle_axis_to_quaternion(np.array([0.3, -0.2, 0.5]))
    got = oracle.quaternion_plus(qa, d)
    want = sp.quat_mul_xyzw(np.array([np.sin(0.1), 0, 0, np.cos(0.1)]), qa)
    np.testing.assert_allclose(got, want, atol=1e-15)


def test_loss_functions(oracle):
    """ceres::HuberLoss / CauchyLoss (Ceres external), chosen by optimization_utils.h:31-47."""
    np.testing.assert_allclose(oracle.loss(1, 2.0, 1.0), [1.0, 1.0, 0.0])
    rho = oracle.loss(1, 2.0, 9.0)
    np.testing.assert_allclose(rho, [2 * 2 * 3 - 4, 2.0 / 3.0, -(2.0 / 3.0) / 18.0])
    rho = oracle.loss(2, 2.0, 9.0)
    np.testing.assert_allclose(rho, [4 * np.log(1 + 9 / 4), 1 / (1 + 9 / 4), -0.25 / (1 + 9 / 4) ** 2])


def test_trust_region_radius_schedule_matches_stored_ceres_log(oracle):
    """demos/imu_camera_calibration.ipynb:350-363 — Ceres's own iteration table for this problem family. Four steps that fail
    to evaluate take the radius 1e4 -> 5e3 -> 1.25e3 -> 1.56e2 -> 9.77 (divide by 2, 4, 8, 16); every accepted step then
    follows radius / max(1/3, 1 - (2 rho - 1)^3). The table prints 3 significant digits, hence the interval check."""
    r, f = 1e4, 2.0
    seen = [r]
    for _ in range(4):
        r /= f
        f *= 2
        seen.append(r)
    np.testing.assert_allclose(seen, [1e4, 5e3, 1.25e3, 1.5625e2, 9.765625])
    # (tr_ratio, tr_radius after the step) rows 5..12 of the stored log
    rows = [(9.13e-01, 2.25e+01), (9.51e-01, 6.74e+01), (8.17e-01, 9.04e+01), (3.71e-01, 8.89e+01), (9.61e-01, 2.67e+02),
            (5.28e-01, 2.67e+02), (5.97e-01, 2.69e+02), (5.83e-01, 2.70e+02)]
    radius = 9.765625
    for ratio, printed in rows:
        lo = oracle.radius_after_accept(radius, ratio - 5e-4)
        hi = oracle.radius_after_accept(radius, ratio + 5e-4)
        lo, hi = min(lo, hi), max(lo, hi)
        assert lo * (1 - 6e-3) <= printed <= hi * (1 + 6e-3), (ratio, printed, lo, hi)
        radius = oracle.radius_after_accept(radius, ratio)


def test_gyro_and_accel_kinematics_against_finite_differences(oracle):
    """gyroscope_test.cpp:106-157 / accelerometer_test.cpp:106-177: the analytic angular velocity and specific force of the
    functors agree with finite differences of the spline pose. Identity intrinsics (scale 1), identity extrinsics."""
    truth = synthetic.build_truth(synthetic.CONFIGS["tiny"], fit=oracle.oracle_api.fit_trajectory)
    spl = truth.spline
    api = oracle.oracle_api()
    api.set_trajectory(6, spl.knots, spl.ctrl)
    api.set_gravity(truth.gravity)
    q_id, z3 = np.array([0.0, 0, 0, 1]), np.zeros(3)
    gid = api.add_sensor(1, 1, "g", np.array([1.0]), q_id, z3, 0.0, 1.0, 0, 1.0, 0, 0, 0)
    aid = api.add_sensor(2, 1, "a", np.array([1.0]), q_id, z3, 0.0, 1.0, 0, 1.0, 0, 0, 0)
    times = np.linspace(0.3, 1.5, 25)
    _, omega = oracle.project_imu(api, gid, times)
    _, acc = oracle.project_imu(api, aid, times)
    h = 1e-5

    def R_rw(t):   # rotation world -> rig = Exp(-phi_wr)
        return np.stack([oracle.exp_so3(-p[:3]) for p in spl.evaluate(t, 0)])
    Rp, Rm, R0 = R_rw(times + h), R_rw(times - h), R_rw(times)
    for i in range(times.size):
        dR = (Rp[i] - Rm[i]) / (2 * h) @ R0[i].T
        w_fd = np.array([dR[2, 1] - dR[1, 2], dR[0, 2] - dR[2, 0], dR[1, 0] - dR[0, 1]]) / 2
        # gyroscope_cost_functor.h:110: omega_gyro = -(q_rg^-1 * J(phi_rw) phidot_rw)
        assert np.sum((omega[i] + w_fd) ** 2) < 1e-5
    tdd = spl.evaluate(times, 2)[:, 3:]
    for i in range(times.size):
        want = R0[i] @ (tdd[i] - truth.gravity)   # zero lever arm: specific force = R_rw (tdd - g)
        assert np.sum((acc[i] - want) ** 2) < 1e-3


@pytest.mark.timeout(300)
def test_reference_integration_test_acceptance(oracle):
    """ToyStereoCameraAndImuCalibration (batch_optimizer_test.cpp:32-213) restated: DefaultSyntheticTest trajectory and
    chart (test_utils.h:11-116), two OpenCv5 cameras + gyroscope + accelerometer, perfect data, the test's initial guess.
    Acceptance as the reference asserts (:185-210): CONVERGENCE, final_cost < 1e-7, every parameter within 1e-7 of the truth."""
    from calico_b200 import _capi
    truth, prob = synthetic.toy_stereo_imu_problem(oracle.oracle_api, seed=3)
    assert prob.spline.ctrl.shape[0] == 185 and len(truth.bodies[0].pts) == 36   # 240 poses -> 185 control points
    o = oracle.oracle_api()
    ids = prob.push(o)
    summ, log = o.optimize(oracle.OracleOptions(linear_solver=1, num_threads=os.cpu_count() or 1))
    assert summ.termination_type == _capi.CONVERGENCE
    assert summ.final_cost < 1e-7
    prob.pull(o, ids)
    for s, t in zip(prob.sensors, truth.sensors):
        np.testing.assert_allclose(s.intr, t.intr, rtol=0, atol=1e-7)
        np.testing.assert_allclose(s.t, t.t, rtol=0, atol=1e-7)
        np.testing.assert_allclose(s.q_xyzw, t.q_xyzw, rtol=0, atol=1e-7)
        assert abs(s.latency - t.latency) < 1e-7
    # the first trial steps are rejected and the radius follows Ceres's stored schedule (previous test): /2, /4, /8
    radii = [it.trust_region_radius for it in log[:4]]
    np.testing.assert_allclose(radii, [1e4, 5e3, 1.25e3, 1.5625e2])
    assert [it.step_is_successful for it in log[1:4]] == [0, 0, 0]


def test_linear_solvers_agree(oracle):
    """The oracle's three linear solvers (dense normal equations, banded Schur, Ceres-ordered dense Schur) take the same
    LM path: the Schur variants are re-orderings of the same elimination."""
    truth, prob = synthetic.generate("micro", oracle.oracle_api, noise=True)
    costs = []
    for ls in (0, 1, 2):
        o = oracle.oracle_api()
        prob.clone().push(o)
        summ, log = o.optimize(oracle.OracleOptions(linear_solver=ls, max_num_iterations=6))
        costs.append([it.cost for it in log])
    np.testing.assert_allclose(costs[1], costs[0], rtol=1e-9)
    np.testing.assert_allclose(costs[2], costs[0], rtol=1e-9)


def test_wide_angle_models_round_trip_like_the_reference_test(oracle):
    """The reference's own test of DoubleSphere / FieldOfView / UnifiedCamera / ExtendedUnifiedCamera (camera_models_test.cpp:170-253):
    project the 61 x 61 ground grid seen from 1 m above its centre (fixture :71-101), invert with the model's closed-form inverse and compare
    the bearing vectors — tolerance 1e-12, and 2e-2 for ExtendedUnifiedCamera whose projection (camera_models.h:995, `beta * norm`, not
    squared) is not the inverse of its own UnprojectPixel. The inverses below are the published closed forms (Usenko et al. 2018 for the
    double sphere / unified / extended unified models, Devernay & Faugeras 2001 for the field-of-view model) written independently in
    numpy: this pins the oracle's ProjectPoint of the four models that have no OpenCV counterpart."""
    R_wc = np.diag([1.0, -1.0, -1.0])
    t_wc = np.array([0.75, 0.75, 1.0])
    xs = np.arange(61) * 0.025
    pts_w = np.array([[x, y, 0.0] for x in xs for y in xs])
    pts_c = (pts_w - t_wc) @ R_wc          # R_wc^T (p - t), R_wc symmetric
    bearing = pts_c / np.linalg.norm(pts_c, axis=1, keepdims=True)

    def inv_double_sphere(intr, px):
        f, cx, cy, xi, al = intr
        mx, my = (px[0] - cx) / f, (px[1] - cy) / f
        r2 = mx * mx + my * my
        mz = (1 - al * al * r2) / (al * np.sqrt(1 - (2 * al - 1) * r2) + 1 - al)
        k = (mz * xi + np.sqrt(mz * mz + (1 - xi * xi) * r2)) / (mz * mz + r2)
        return np.array([k * mx, k * my, k * mz - xi])

    def inv_fov(intr, px):
        f, cx, cy, w = intr
        mx, my = (px[0] - cx) / f, (px[1] - cy) / f
        rd = np.hypot(mx, my)
        eta = np.sin(rd * w) / (2 * rd * np.tan(w / 2)) if rd > 1e-8 else w / (2 * np.tan(w / 2))
        return np.array([eta * mx, eta * my, np.cos(rd * w)])

    def inv_unified(intr, px):
        f, cx, cy, al = intr
        mx, my = (1 - al) * (px[0] - cx) / f, (1 - al) * (px[1] - cy) / f
        r2 = mx * mx + my * my
        xi = al / (1 - al)
        k = (xi + np.sqrt(1 + (1 - xi * xi) * r2)) / (1 + r2)
        return np.array([k * mx, k * my, k - xi])

    def inv_extended_unified(intr, px):
        f, cx, cy, al, be = intr
        mx, my = (px[0] - cx) / f, (px[1] - cy) / f
        r2 = mx * mx + my * my
        mz = (1 - be * al * al * r2) / (al * np.sqrt(1 - (2 * al - 1) * be * r2) + 1 - al)
        return np.array([mx, my, mz])

    cases = [(4, [785.0, 640, 400, 0.5, 0.5], inv_double_sphere, 1e-12), (5, [785.0, 640, 400, 0.05], inv_fov, 1e-12),
             (6, [785.0, 640, 400, 0.5], inv_unified, 1e-12), (7, [785.0, 640, 400, 0.5, 0.5], inv_extended_unified, 2e-2)]
    for model, intr, inverse, tol in cases:
        worst = 0.0
        for p, b in zip(pts_c, bearing):
            ok, px = oracle.project_point(model, intr, p)
            assert ok
            v = inverse(intr, px)
            worst = max(worst, np.abs(v / np.linalg.norm(v) - b).max())
        assert worst <= tol, (model, worst)
