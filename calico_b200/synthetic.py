"""Seeded synthetic calibration problems of the BASELINE.json shapes (recipe: SURVEY.md §8d).

Measurements are the reference's `Camera::Project` / `Gyroscope::Project` / `Accelerometer::Project`
(calico/sensors/camera.cpp:155-208, gyroscope.cpp:56-82, accelerometer.cpp:76-123) evaluated at the ground truth: the
projection at time t is produced by a *backend* (any handle exposing the calico_b200 C-ABI shape) as the residual of a
zero measurement with unit sigma and zero latency, `proj = -r`, and then stamped t + latency as the reference does
(camera.cpp:179,198). bench.py uses the CUDA library itself as the backend (batched projection on the GPU, SURVEY §8f
rank 3); the CPU tests pass the oracle.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List

import numpy as np

from . import spline as sp
from .spec import ACCELEROMETER, CAMERA, GYROSCOPE, ProblemSpec, RigidBodySpec, SensorSpec

SEED = 20261017

# camera_models_test.cpp:107-108 / :153-154; batch_optimizer_test.cpp:90,95
OPENCV5_TRUTH = np.array([785.0, 640.0, 400.0, -3.149e-1, 1.069e-1, 1.616e-4, 1.141e-4, -1.853e-2])
KB_TRUTH = np.array([785.0, 640.0, 400.0, -3.149e-1, 1.069e-1, 1.616e-4, 1.141e-4])
IMU_TRUTH = np.array([1.3, 0.01, -0.01, 0.01])
# Intrinsics the reference's own model tests use (camera_models_test.cpp:107-108,128-130,153-154,174-175,195-196,216-217,237-238).
CAMERA_TRUTH = {
    1: OPENCV5_TRUTH,
    2: np.array([785.0, 640.0, 400.0, -3.149e-1, 1.069e-1, 1.616e-4, 1.141e-4, -1.853e-2, 1.225e-1, -5.26e-2, 8.58e-3]),
    3: KB_TRUTH,
    4: np.array([785.0, 640.0, 400.0, 0.5, 0.5]),
    5: np.array([785.0, 640.0, 400.0, 0.05]),
    6: np.array([785.0, 640.0, 400.0, 0.5]),
    7: np.array([785.0, 640.0, 400.0, 0.5, 0.5]),
}
IMU_MODEL_TRUTH = {
    1: np.array([1.3]),
    2: IMU_TRUTH,
    3: np.array([1.3, 1.2, 0.9, 0.01, -0.02, 0.015, 0.03, -0.01, 0.02, 0.01, -0.01, 0.01]),
}


def aprilgrid_points(tag_rows=6, tag_cols=6, tag_size=0.088, tag_spacing=0.3):
    """Chart geometry of AprilGridDetector::SetupDetector, aprilgrid_detector.cpp:28-50: feature id = tag*4 + corner."""
    ids, pts = [], []
    w = tag_size * (1.0 + tag_spacing)
    for row in range(tag_rows):
        for col in range(tag_cols):
            tag = row * tag_cols + col
            for k in range(4):
                ids.append(tag * 4 + k)
                pts.append([w * col + tag_size * (k in (1, 2)), w * row + tag_size * (k in (2, 3)), 0.0])
    return np.array(ids, dtype=np.int32), np.array(pts)


@dataclass
class Config:
    name: str
    n_cameras: int
    camera_model: int        # CameraIntrinsicsModel value
    n_imus: int
    n_frames: int
    corners_per_image: int   # 0 = all visible corners
    huber: bool = False
    cauchy: bool = False       # ceres::CauchyLoss(1.0) on the cameras (optimization_utils.h:40-41; the reference's Kalibr demo uses it)
    outlier_fraction: float = 0.0
    frame_rate: float = 20.0
    imu_rate: float = 200.0
    knot_frequency: float = 10.0
    imu_is_rig: bool = False   # IMU frame = rig frame: IMU extrinsics/latency fixed, every camera block free
    camera_models: tuple = ()  # per-camera model override (cycled); empty = camera_model for all
    imu_models: tuple = ()     # per-IMU (gyro, accel) model override (cycled); empty = ScaleAndBias


CONFIGS = {
    # BASELINE.json configs[0..4]
    "C1": Config("C1", 1, 1, 0, 50, 0),
    "C2": Config("C2", 1, 1, 1, 500, 0, imu_is_rig=True),
    "C3": Config("C3", 4, 3, 0, 2000, 25),
    "C4": Config("C4", 8, 1, 1, 5000, 25),
    "C5": Config("C5", 16, 1, 2, 10000, 25, huber=True, outlier_fraction=0.02),
    # small shapes for tests / smoke
    "tiny": Config("tiny", 2, 1, 1, 40, 12),
    "tiny_kb": Config("tiny_kb", 2, 3, 0, 40, 12),
    "small": Config("small", 2, 1, 1, 160, 10),
    "small_huber": Config("small_huber", 2, 1, 1, 160, 10, huber=True, outlier_fraction=0.03),
    "micro": Config("micro", 1, 1, 1, 24, 6),
    "small_cauchy": Config("small_cauchy", 2, 1, 1, 160, 10, cauchy=True, outlier_fraction=0.03),
    # the C5 shape (16 cameras + 2 IMUs, Huber, 2 % outliers: N_c = 279 calibration unknowns) on a short trajectory
    "C5_like": Config("C5_like", 16, 1, 2, 300, 25, huber=True, outlier_fraction=0.02),
    # every camera and IMU intrinsics model once (Jacobian parity of the "next" models)
    "tiny_models": Config("tiny_models", 7, 1, 3, 30, 10, camera_models=(1, 2, 3, 4, 5, 6, 7), imu_models=(1, 2, 3)),
}


def truth_trajectory(cfg: Config, rng: np.random.Generator, fit=None) -> sp.Spline:
    """Camera-facing-chart base pose (test_utils.h:16-20: Rz(pi)*Rx(pi), 1 m stand-off) + smooth excitation: sum of 3
    sinusoids per axis (+-20 deg, +-0.3 m, periods 3-11 s), sampled at frame times and fitted like Trajectory::FitSpline."""
    t = np.arange(cfg.n_frames) / cfg.frame_rate
    amp_r, amp_t = np.deg2rad(20.0) / 3.0, 0.3 / 3.0
    rot = np.zeros((t.size, 3))
    pos = np.zeros((t.size, 3))
    for axis in range(3):
        for _ in range(3):
            period, phase = rng.uniform(3.0, 11.0), rng.uniform(0, 2 * np.pi)
            rot[:, axis] += amp_r * np.sin(2 * np.pi * t / period + phase)
            period, phase = rng.uniform(3.0, 11.0), rng.uniform(0, 2 * np.pi)
            pos[:, axis] += amp_t * np.sin(2 * np.pi * t / period + phase)
    q0 = sp.quat_mul_xyzw(sp.angle_axis_to_quat_xyzw([0, 0, np.pi])[0], sp.angle_axis_to_quat_xyzw([np.pi, 0, 0])[0])
    q = sp.quat_mul_xyzw(q0[None, :], sp.angle_axis_to_quat_xyzw(rot))
    chart_center = np.array([0.343, 0.343, 0.0])  # look at the middle of the 6x6 AprilGrid
    pos = pos + np.array([0.0, 0.0, 1.0]) + chart_center
    return (fit or sp.fit_trajectory)(t, q, pos, cfg.knot_frequency, 6)


def _camera_extrinsics(i: int, n: int):
    """Cameras on a ring: baseline 0.1 m, +-10 deg yaw."""
    if n == 1:
        return np.array([0.0, 0.0, 0.0, 1.0]), np.zeros(3)
    ang = 2 * np.pi * i / n
    yaw = np.deg2rad(10.0) * np.cos(ang)
    pitch = np.deg2rad(10.0) * np.sin(ang)
    q = sp.quat_mul_xyzw(sp.angle_axis_to_quat_xyzw([0, yaw, 0])[0], sp.angle_axis_to_quat_xyzw([pitch, 0, 0])[0])
    if i == 0:
        return np.array([0.0, 0.0, 0.0, 1.0]), np.zeros(3)
    return q, 0.05 * np.array([np.cos(ang), np.sin(ang), 0.0])


def build_truth(cfg: Config, seed: int = SEED, fit=None) -> ProblemSpec:
    rng = np.random.default_rng(seed)
    spl = truth_trajectory(cfg, rng, fit)
    ids, pts = aprilgrid_points()
    spec = ProblemSpec(spline=spl)
    spec.bodies.append(RigidBodySpec(0, np.array([0.0, 0, 0, 1]), np.zeros(3), ids, pts, True, True))
    for i in range(cfg.n_cameras):
        q, t = _camera_extrinsics(i, cfg.n_cameras)
        model = cfg.camera_models[i % len(cfg.camera_models)] if cfg.camera_models else cfg.camera_model
        intr = CAMERA_TRUTH[model].copy()
        if cfg.name == "C1":
            intr[3:] = 0.0   # "pinhole": true distortion = 0
        spec.sensors.append(SensorSpec(CAMERA, model, f"cam{i}", intr, q, t, latency=0.0 if i == 0 else 0.01, sigma=0.1))
    for i in range(cfg.n_imus):
        qg = sp.angle_axis_to_quat_xyzw(np.deg2rad(2.0) * _unit(rng))[0]
        qa = sp.angle_axis_to_quat_xyzw(np.deg2rad(2.0) * _unit(rng))[0]
        if cfg.imu_is_rig:
            qg = qa = np.array([0.0, 0.0, 0.0, 1.0])
        im = cfg.imu_models[i % len(cfg.imu_models)] if cfg.imu_models else 2
        spec.sensors.append(SensorSpec(GYROSCOPE, im, f"gyro{i}", IMU_MODEL_TRUTH[im].copy(), qg, np.zeros(3), latency=0.02, sigma=1e-3))
        spec.sensors.append(SensorSpec(ACCELEROMETER, im, f"accel{i}", IMU_MODEL_TRUTH[im].copy(), qa, 0.02 * rng.standard_normal(3) if i else np.zeros(3),
                                       latency=0.02, sigma=1e-2))
    return spec


def _unit(rng):
    v = rng.standard_normal(3)
    return v / np.linalg.norm(v)


def project_sensor(api_factory: Callable, truth: ProblemSpec, s_idx: int, times: np.ndarray, feature_ids=None):
    """Reference `*::Project` at ground truth through a backend: residual of a zero measurement, unit sigma, zero latency.

    camera → (proj [n_t, n_feat, 2], visible [n_t, n_feat]); imu → proj [n_t, 3].
    """
    s = truth.sensors[s_idx]
    probe = ProblemSpec(spline=truth.spline, gravity=truth.gravity, bodies=truth.bodies)
    ps = SensorSpec(s.kind, s.model, s.name, s.intr, s.q_xyzw, s.t, latency=0.0, sigma=1.0)
    times = np.asarray(times, dtype=np.float64)
    if s.kind == CAMERA:
        body = truth.bodies[0]
        fids = body.feature_ids if feature_ids is None else np.asarray(feature_ids)
        nt, nf = times.size, fids.size
        ps.stamp = np.repeat(times, nf)
        ps.image_id = np.repeat(np.arange(nt), nf).astype(np.int32)
        ps.model_id = np.full(nt * nf, body.id, dtype=np.int32)
        ps.feature_id = np.tile(fids, nt).astype(np.int32)
        ps.meas = np.zeros((nt * nf, 2))
    else:
        ps.stamp = times
        ps.meas = np.zeros((times.size, 3))
    probe.sensors.append(ps)
    api = api_factory()
    try:
        (sid,) = probe.push(api)
        r, _, valid = api.evaluate_sensor(sid, want_jac=False)
    finally:
        api.close()
    if s.kind == CAMERA:
        return (-r).reshape(nt, nf, 2), valid.reshape(nt, nf)
    return -r


def generate(cfg_name: str, api_factory: Callable, seed: int = SEED, noise: bool = True, chunk_frames: int = 512):
    """Returns (truth, problem): `truth` holds the ground-truth state, `problem` the measurements + initial guess."""
    cfg = CONFIGS[cfg_name] if isinstance(cfg_name, str) else cfg_name
    # The trajectory fit runs where the projections run: cb2_fit_trajectory on the device for a C-ABI handle factory, the CPU oracle's
    # restatement when the factory is the oracle's (it carries a `fit_trajectory` attribute; tests only).
    truth = build_truth(cfg, seed, getattr(api_factory, "fit_trajectory", None))
    rng = np.random.default_rng(seed + 1)
    frame_t = np.arange(cfg.n_frames) / cfg.frame_rate
    t_end = frame_t[-1]
    imu_t = np.arange(0.0, t_end + 1e-9, 1.0 / cfg.imu_rate)
    # keep stamps (t + latency) inside the valid knots
    last_valid = truth.spline.valid_knots[-1]
    body = truth.bodies[0]
    problem = truth.clone()
    for si, s in enumerate(truth.sensors):
        ps = problem.sensors[si]
        if s.kind == CAMERA:
            times = frame_t[frame_t + s.latency < last_valid]
            stamps, img, fid, pix = [], [], [], []
            for c0 in range(0, times.size, chunk_frames):
                tt = times[c0:c0 + chunk_frames]
                proj, vis = project_sensor(api_factory, truth, si, tt)
                for j in range(tt.size):
                    cand = np.nonzero(vis[j])[0]
                    if cfg.corners_per_image and cand.size > cfg.corners_per_image:
                        cand = np.sort(rng.choice(cand, cfg.corners_per_image, replace=False))
                    stamps.append(np.full(cand.size, tt[j] + s.latency))
                    img.append(np.full(cand.size, c0 + j, dtype=np.int32))
                    fid.append(body.feature_ids[cand])
                    pix.append(proj[j, cand])
            ps.stamp = np.concatenate(stamps)
            ps.image_id = np.concatenate(img)
            ps.feature_id = np.concatenate(fid).astype(np.int32)
            ps.model_id = np.full(ps.stamp.size, body.id, dtype=np.int32)
            ps.meas = np.concatenate(pix)
            if noise:
                ps.meas = ps.meas + s.sigma * rng.standard_normal(ps.meas.shape)
            if cfg.outlier_fraction > 0:
                bad = rng.random(ps.stamp.size) < cfg.outlier_fraction
                ps.meas[bad] += rng.uniform(-20, 20, size=(int(bad.sum()), 2))
            if cfg.huber:
                ps.loss_type, ps.loss_scale = 1, 1.0
            if cfg.cauchy:
                ps.loss_type, ps.loss_scale = 2, 1.0
        else:
            times = imu_t[imu_t + s.latency < last_valid]
            proj = project_sensor(api_factory, truth, si, times)
            ps.stamp = times + s.latency
            ps.seq = np.arange(times.size, dtype=np.int32)
            ps.meas = proj + (s.sigma * rng.standard_normal(proj.shape) if noise else 0.0)
    # Initial guess as batch_optimizer_test.cpp:125-128,162-163: intrinsics x1.01 with distortion zeroed, extrinsic
    # translations +1 cm noise, latencies 0; sensor 0's extrinsics + latency fixed (gauge).
    first_cam = True
    for si, s in enumerate(problem.sensors):
        if s.kind == CAMERA:
            s.intr = 1.01 * s.intr
            if s.model in (1, 2, 3):
                s.intr[3:] = 0.0
            s.en_intr = True
            if cfg.name == "C1":
                s.en_extr = s.en_lat = False   # intrinsics-only
            elif first_cam and not cfg.imu_is_rig:
                s.en_extr = s.en_lat = False
            else:
                s.en_extr = s.en_lat = True
                s.t = s.t + 0.01 * rng.standard_normal(3)
                s.latency = 0.0
            first_cam = False
        else:
            s.intr = 1.01 * s.intr
            s.en_intr = True
            if cfg.imu_is_rig:
                s.en_extr = s.en_lat = False   # stays at truth
                continue
            s.en_extr = s.en_lat = True
            if s.kind == ACCELEROMETER:
                s.t = s.t + 0.01 * rng.standard_normal(3)
            s.latency = 0.0
    return truth, problem


# ---- the reference's own integration fixture -------------------------------------------------------------------------
def default_synthetic_test():
    """DefaultSyntheticTest (calico/test_utils.h:11-116): base pose Rz(pi)*Rx(pi) at z = 1 m; per axis 4 angular segments
    (+-30 deg) then 4 linear segments (+-0.5 m), 10 cosine-eased samples per segment, dt = 0.075 s -> 240 poses; 6x6 planar
    points with 0.3 m pitch. Returns (stamps, q_xyzw[n,4], t[n,3], points[36,3])."""
    q0 = sp.quat_mul_xyzw(sp.angle_axis_to_quat_xyzw([0, 0, np.pi])[0], sp.angle_axis_to_quat_xyzw([np.pi, 0, 0])[0])
    t0 = np.array([0.0, 0.0, 1.0])
    ang = [0.0, np.deg2rad(30.0), 0.0, -np.deg2rad(30.0), 0.0]
    pos = [0.0, 0.5, 0.0, -0.5, 0.0]
    n_per, seg_dur = 10, 0.75
    dt_interp = 1.0 / n_per
    dt_actual = dt_interp * seg_dur
    interp = [(np.sin(dt_interp * i * np.pi - np.pi / 2) + 1.0) / 2.0 for i in range(n_per)]
    stamps, qs, ts = [], [], []
    cur = 0.0
    for axis in np.eye(3):
        for i in range(1, len(ang)):
            for it in interp:
                theta = (ang[i] - ang[i - 1]) * it + ang[i - 1]
                qs.append(sp.quat_mul_xyzw(q0, sp.angle_axis_to_quat_xyzw(theta * axis)[0]))
                ts.append(t0.copy())
                stamps.append(cur)
                cur += dt_actual
        for i in range(1, len(pos)):
            for it in interp:
                p = (pos[i] - pos[i - 1]) * it + pos[i - 1]
                qs.append(q0.copy())
                ts.append(axis * p + t0)
                stamps.append(cur)
                cur += dt_actual
    pts = np.array([[i * 0.3 - 0.75, j * 0.3 - 0.75, 0.0] for i in range(6) for j in range(6)])
    return np.array(stamps), np.array(qs), np.array(ts), pts


def toy_stereo_imu_problem(api_factory: Callable, seed: int = 1):
    """ToyStereoCameraAndImuCalibration (calico/test/batch_optimizer_test.cpp:32-213): two OpenCv5 cameras + gyroscope +
    accelerometer (ScaleAndBias) on DefaultSyntheticTest, perfect data from the `*::Project` restatements, the test's initial
    guess. The reference draws its perturbations from an unseeded Eigen Random(); here they are seeded. Returns (truth, problem)."""
    rng = np.random.default_rng(seed)
    stamps, q, t, pts = default_synthetic_test()
    spl = (getattr(api_factory, "fit_trajectory", None) or sp.fit_trajectory)(stamps, q, t, 10.0, 6)
    truth = ProblemSpec(spline=spl)
    truth.bodies.append(RigidBodySpec(0, np.array([0.0, 0, 0, 1]), np.zeros(3), np.arange(36, dtype=np.int32), pts, True, True))

    def small_rot(deg):
        return sp.angle_axis_to_quat_xyzw(np.deg2rad(deg) * _unit(rng))[0]
    truth.sensors.append(SensorSpec(CAMERA, 1, "Left", OPENCV5_TRUTH.copy(), sigma=1.0))
    truth.sensors.append(SensorSpec(CAMERA, 1, "Right", OPENCV5_TRUTH.copy(), small_rot(2.0), 0.05 * rng.uniform(-1, 1, 3), latency=0.01, sigma=1.0))
    truth.sensors.append(SensorSpec(GYROSCOPE, 2, "Gyroscope", IMU_TRUTH.copy(), small_rot(2.0), np.zeros(3), latency=0.02, sigma=1.0))
    truth.sensors.append(SensorSpec(ACCELEROMETER, 2, "Accelerometer", IMU_TRUTH.copy(), small_rot(2.0), np.zeros(3), latency=0.02, sigma=1.0))
    problem = truth.clone()
    body = truth.bodies[0]
    for si, s in enumerate(truth.sensors):
        ps = problem.sensors[si]
        if s.kind == CAMERA:
            proj, vis = project_sensor(api_factory, truth, si, stamps)
            ti, fi = np.nonzero(vis)
            ps.stamp = stamps[ti] + s.latency
            ps.image_id = ti.astype(np.int32)
            ps.feature_id = body.feature_ids[fi]
            ps.model_id = np.zeros(ti.size, dtype=np.int32)
            ps.meas = proj[ti, fi]
        else:
            ps.stamp = stamps + s.latency
            ps.seq = np.arange(stamps.size, dtype=np.int32)
            ps.meas = project_sensor(api_factory, truth, si, stamps)
    init_intr = 1.01 * OPENCV5_TRUTH
    init_intr[3:] = 0.0
    left, right, gyro, accel = problem.sensors
    left.intr, left.en_intr = init_intr.copy(), True
    right.intr, right.en_intr, right.en_extr, right.en_lat = init_intr.copy(), True, True, True
    right.t = right.t + 0.01 * rng.uniform(-1, 1, 3)
    right.latency = 0.0
    gyro.intr, gyro.en_intr, gyro.en_extr, gyro.en_lat, gyro.latency = 1.01 * IMU_TRUTH, True, True, True, 0.0
    accel.intr, accel.en_intr, accel.en_extr, accel.en_lat, accel.latency = 1.01 * IMU_TRUTH, True, True, True, 0.0
    accel.t = accel.t + 0.05 * rng.uniform(-1, 1, 3)
    return truth, problem
