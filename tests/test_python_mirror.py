"""calico_b200.api — the Python mirror of the reference's pybind module (calico/calico.cpp) — driven by the reference's integration
test (calico/test/batch_optimizer_test.cpp:32-213, fixture calico/test_utils.h:11-116) written as a notebook user would write it:
on the SIMT emulation (CPU, reduced fixture, a few iterations) and on the GPU (full acceptance)."""
import os
import sys

import numpy as np
import pytest

import calico_b200.api as calico

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emul"))


def _quat_wxyz(angle, axis):
    axis = np.asarray(axis, float) / np.linalg.norm(axis)
    return np.concatenate([[np.cos(0.5 * angle)], np.sin(0.5 * angle) * axis])


def _qmul(a, b):   # w, x, y, z
    aw, ax, ay, az = a; bw, bx, by, bz = b
    return np.array([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx])


def _fixture(samples_per_segment):
    """test_utils.h:11-116."""
    q0 = _qmul(_quat_wxyz(np.pi, [0, 0, 1]), _quat_wxyz(np.pi, [1, 0, 0]))
    t0 = np.array([0.0, 0.0, 1.0])
    ang = np.deg2rad([0.0, 30.0, 0.0, -30.0, 0.0]); pos = [0.0, 0.5, 0.0, -0.5, 0.0]
    dti = 1.0 / samples_per_segment
    interp = [(np.sin(dti * i * np.pi - np.pi / 2) + 1.0) / 2.0 for i in range(samples_per_segment)]
    poses, now = {}, 0.0
    for axis in np.eye(3):
        for i in range(1, 5):
            for u in interp:
                p = calico.Pose3d(); p.rotation = _qmul(q0, _quat_wxyz((ang[i] - ang[i - 1]) * u + ang[i - 1], axis)); p.translation = t0
                poses[now] = p; now += dti * 0.75
        for i in range(1, 5):
            for u in interp:
                p = calico.Pose3d(); p.rotation = q0; p.translation = axis * ((pos[i] - pos[i - 1]) * u + pos[i - 1]) + t0
                poses[now] = p; now += dti * 0.75
    points = [np.array([i * 0.3 - 0.75, j * 0.3 - 0.75, 0.0]) for i in range(6) for j in range(6)]
    return poses, sorted(poses.keys()), points


def _toy_calibration(samples_per_segment, max_iterations, free_world_pose=False):
    rng = np.random.default_rng(5)
    poses, stamps, points = _fixture(samples_per_segment)
    planar_target = calico.RigidBody(world_pose_is_constant=not free_world_pose, model_definition_is_constant=True)
    for i, p in enumerate(points):
        planar_target.model_definition[i] = p
    world_model = calico.WorldModel()
    true_gravity = world_model.GetGravity()
    world_model.AddRigidBody(planar_target)
    with pytest.raises(RuntimeError, match="Error: "):
        world_model.AddRigidBody(planar_target)
    trajectory = calico.Trajectory()
    trajectory.FitSpline(poses)
    assert np.abs(trajectory.Interpolate([stamps[3]])[0].translation - poses[stamps[3]].translation).max() < 1e-3

    k_model = calico.CameraIntrinsicsModel.kOpenCv5
    true_intr = np.array([785, 640, 400, -3.149e-1, 1.069e-1, 1.616e-4, 1.141e-4, -1.853e-2])
    T_right = calico.Pose3d(); T_right.rotation = _quat_wxyz(np.deg2rad(2.0), rng.uniform(-1, 1, 3)); T_right.translation = 0.05 * rng.uniform(-1, 1, 3)
    true_left, true_right = calico.Camera(), calico.Camera()
    with pytest.raises(RuntimeError, match="model has not been set"):
        true_left.SetIntrinsics(true_intr)
    for cam in (true_left, true_right):
        cam.SetModel(k_model); cam.SetIntrinsics(true_intr)
    with pytest.raises(RuntimeError, match="Expected intrinsics size of 8"):
        true_left.SetIntrinsics(true_intr[:7])
    true_right.SetExtrinsics(T_right); true_right.SetLatency(0.01)
    meas_left = true_left.Project(stamps, trajectory, world_model)
    meas_right = true_right.Project(stamps, trajectory, world_model)
    true_gyro_intr = true_acc_intr = np.array([1.3, 0.01, -0.01, 0.01])
    T_gyro = calico.Pose3d(); T_gyro.rotation = _quat_wxyz(np.deg2rad(2.0), rng.uniform(-1, 1, 3))
    T_acc = calico.Pose3d(); T_acc.rotation = _quat_wxyz(np.deg2rad(2.0), rng.uniform(-1, 1, 3))
    true_gyro = calico.Gyroscope(); true_gyro.SetModel(calico.GyroscopeIntrinsicsModel.kGyroscopeScaleAndBias); true_gyro.SetIntrinsics(true_gyro_intr)
    true_gyro.SetExtrinsics(T_gyro); true_gyro.SetLatency(0.02)
    true_acc = calico.Accelerometer(); true_acc.SetModel(calico.AccelerometerIntrinsicsModel.kAccelerometerScaleAndBias); true_acc.SetIntrinsics(true_acc_intr)
    true_acc.SetExtrinsics(T_acc); true_acc.SetLatency(0.02)
    meas_gyro = true_gyro.Project(stamps, trajectory, world_model)
    meas_acc = true_acc.Project(stamps, trajectory, world_model)
    assert len(meas_gyro) == len(stamps) and len(meas_left) > 0

    init_intr = 1.01 * true_intr; init_intr[3:] = 0.0
    T_right0 = calico.Pose3d(T_right); T_right0.translation = T_right.translation + 0.01 * rng.uniform(-1, 1, 3)
    camera_left = calico.Camera(); camera_left.SetName("Left"); camera_left.SetModel(k_model); camera_left.SetIntrinsics(init_intr)
    camera_left.EnableExtrinsicsEstimation(False); camera_left.EnableIntrinsicsEstimation(True); camera_left.EnableLatencyEstimation(False)
    camera_left.AddMeasurements(meas_left)
    with pytest.raises(RuntimeError, match="redundant measurement"):
        camera_left.AddMeasurement(meas_left[0])
    camera_right = calico.Camera(); camera_right.SetName("Right"); camera_right.SetModel(k_model); camera_right.SetIntrinsics(init_intr)
    camera_right.SetExtrinsics(T_right0)
    camera_right.EnableExtrinsicsEstimation(True); camera_right.EnableIntrinsicsEstimation(True); camera_right.EnableLatencyEstimation(True)
    camera_right.AddMeasurements(meas_right)
    gyroscope = calico.Gyroscope(); gyroscope.SetName("Gyroscope"); gyroscope.SetModel(calico.GyroscopeIntrinsicsModel.kGyroscopeScaleAndBias)
    gyroscope.SetIntrinsics(1.01 * true_gyro_intr); gyroscope.SetExtrinsics(T_gyro)
    gyroscope.EnableExtrinsicsEstimation(True); gyroscope.EnableIntrinsicsEstimation(True); gyroscope.EnableLatencyEstimation(True)
    gyroscope.AddMeasurements(meas_gyro)
    T_acc0 = calico.Pose3d(T_acc); T_acc0.translation = T_acc.translation + 0.05 * rng.uniform(-1, 1, 3)
    accelerometer = calico.Accelerometer(); accelerometer.SetName("Accelerometer")
    accelerometer.SetModel(calico.AccelerometerIntrinsicsModel.kAccelerometerScaleAndBias)
    accelerometer.SetIntrinsics(1.01 * true_acc_intr); accelerometer.SetExtrinsics(T_acc0)
    accelerometer.EnableExtrinsicsEstimation(True); accelerometer.EnableIntrinsicsEstimation(True); accelerometer.EnableLatencyEstimation(True)
    accelerometer.AddMeasurements(meas_acc)

    optimizer = calico.BatchOptimizer()
    for s in (camera_left, camera_right, gyroscope, accelerometer):
        optimizer.AddSensor(s)
    optimizer.AddWorldModel(world_model)
    optimizer.AddTrajectory(trajectory)
    options = calico.DefaultSolverOptions()
    options.minimizer_progress_to_stdout = False
    options.max_num_iterations = max_iterations
    summary = optimizer.Optimize(options)
    assert summary.IsSolutionUsable() and "Iterations" in summary.BriefReport()
    truth = dict(intr=true_intr, T_right=T_right, T_gyro=T_gyro, T_acc=T_acc, gyro_intr=true_gyro_intr, acc_intr=true_acc_intr, gravity=true_gravity)
    return summary, truth, (camera_left, camera_right, gyroscope, accelerometer), world_model, meas_right


def _pose_close(a, b, tol):
    qa, qb = a.rotation, b.rotation
    if qa @ qb < 0:
        qb = -qb
    return np.abs(qa - qb).max() < tol and np.abs(a.translation - b.translation).max() < tol


@pytest.mark.timeout(900)
def test_python_mirror_on_emulated_kernels():
    import build as emul_build
    calico.set_library(emul_build.build())
    try:
        summary, *_ = _toy_calibration(samples_per_segment=1, max_iterations=1)      # plumbing only; acceptance is the GPU test below
        assert summary.final_cost <= summary.initial_cost and summary.num_residual_blocks > 0
        # world_pose_is_constant = false (world_model.cpp:66-70): the chart pose is estimated and written back into the RigidBody object
        summary2, _, _, world_model, _ = _toy_calibration(samples_per_segment=1, max_iterations=6, free_world_pose=True)
        body = next(iter(world_model._rigidbodies.values()))
        assert summary2.num_parameters_reduced == summary.num_parameters_reduced + 7
        assert np.abs(body.T_world_rigidbody.translation).max() > 0 or abs(body.T_world_rigidbody.rotation[0] - 1.0) > 0
    finally:
        calico.set_library(None)


@pytest.mark.gpu
def test_python_mirror_reference_integration_test(product_lib):
    summary, truth, (left, right, gyro, acc), world_model, meas_right = _toy_calibration(samples_per_segment=10, max_iterations=50)
    k = 1e-7                                                   # batch_optimizer_test.cpp:185-210
    assert summary.termination_type == 0 and summary.final_cost < k
    assert np.abs(left.GetIntrinsics() - truth["intr"]).max() < k and np.abs(right.GetIntrinsics() - truth["intr"]).max() < k
    assert _pose_close(right.GetExtrinsics(), truth["T_right"], k) and abs(right.GetLatency() - 0.01) < k
    assert np.abs(gyro.GetIntrinsics() - truth["gyro_intr"]).max() < k and _pose_close(gyro.GetExtrinsics(), truth["T_gyro"], k)
    assert abs(gyro.GetLatency() - 0.02) < k
    assert np.abs(acc.GetIntrinsics() - truth["acc_intr"]).max() < k and _pose_close(acc.GetExtrinsics(), truth["T_acc"], k)
    assert abs(acc.GetLatency() - 0.02) < k
    assert np.abs(world_model.GetGravity() - truth["gravity"]).max() < k
    pairs = right.GetMeasurementResidualPairs()                # camera.cpp:258-279 after UpdateResiduals
    assert len(pairs) == len(meas_right) and max(np.abs(r).max() for _, r in pairs) < 1e-3
    # outlier marking -> re-optimise loop of the notebooks (camera.cpp:281-299)
    right.MarkOutliersById([m.id for m, _ in pairs[:10]])
    with pytest.raises(RuntimeError, match="not within the measurement set"):
        right.MarkOutlierById(calico.CameraObservationId(-1.0, 0, 0, 0))


def _every_model_project(n_poses=12):
    """SetModel -> SetIntrinsics -> Project for every camera / IMU model (the reference's parameter counts: camera_models.h:79,231,395,596,716,
    848,961), with a landmark in front of and a rigid-body point behind the camera (camera_test.cpp:113-237: 1 / 0 / 4 / 0 measurements)."""
    from calico_b200 import synthetic
    poses, all_stamps, _ = _fixture(2)
    stamps = all_stamps[:n_poses]
    trajectory = calico.Trajectory()
    trajectory.FitSpline(poses)
    world = calico.WorldModel()
    body = calico.RigidBody()
    body.id = 0
    body.model_definition = {0: np.array([0.0, 0.0, 0.0]), 1: np.array([0.1, 0.0, 0.0]), 2: np.array([0.0, 0.1, 0.0]), 3: np.array([0.1, 0.1, 0.0]),
                             4: np.array([0.0, 0.0, 5.0])}        # id 4 lies behind the camera (the rig looks down at z = 0 from z = 1)
    world.AddRigidBody(body)
    world.AddLandmark(calico.Landmark(np.array([0.05, 0.05, 0.0]), 7))
    world.AddLandmark(calico.Landmark(np.array([0.0, 0.0, 3.0]), 8))      # behind
    times = [t for t in stamps if 0.0 < t < stamps[-1]]
    want = {1: 8, 2: 11, 3: 7, 4: 5, 5: 4, 6: 4, 7: 5}
    for model in calico.CameraIntrinsicsModel:
        if model == calico.CameraIntrinsicsModel.kNone:
            continue
        cam = calico.Camera()
        cam.SetModel(model)
        assert cam.GetIntrinsics().size == want[int(model)]
        with pytest.raises(RuntimeError, match="Expected intrinsics size"):
            cam.SetIntrinsics(np.zeros(want[int(model)] + 1))
        cam.SetIntrinsics(synthetic.CAMERA_TRUTH[int(model)])
        ms = cam.Project(times, trajectory, world)
        ids = {(m.id.model_id, m.id.feature_id) for m in ms}
        # in front: landmark 7 (model id -1) and body points 0..3; behind (z <= 0): landmark 8 and body point 4, skipped for EVERY model
        assert ids == {(-1, 7), (0, 0), (0, 1), (0, 2), (0, 3)}, (model, ids)
        assert len(ms) == 5 * len(times)
        assert all(np.isfinite(m.pixel).all() for m in ms)
    for cls, enum_cls in ((calico.Gyroscope, calico.GyroscopeIntrinsicsModel), (calico.Accelerometer, calico.AccelerometerIntrinsicsModel)):
        for model in enum_cls:
            if int(model) == 0:
                continue
            imu = cls()
            imu.SetModel(model)
            assert imu.GetIntrinsics().size == {1: 1, 2: 4, 3: 12}[int(model)]
            imu.SetIntrinsics(synthetic.IMU_MODEL_TRUTH[int(model)])
            assert len(imu.Project(times, trajectory, world)) == len(times)


@pytest.mark.timeout(900)
def test_every_model_through_the_mirror_on_emulated_kernels():
    import build as emul_build
    calico.set_library(emul_build.build())
    try:
        _every_model_project(n_poses=6)
    finally:
        calico.set_library(None)


@pytest.mark.gpu
def test_every_model_through_the_mirror(product_lib):
    _every_model_project()
