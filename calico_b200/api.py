"""Python mirror of the reference's pybind module `_calico` (calico/calico.cpp) over the C ABI: same class and method names, same
argument meaning, same error behaviour (a non-OK status raises RuntimeError("Error: " + message), calico.cpp:417-421), so the call
sequences of the reference's notebooks (demos/kalibr_multicam_demo.ipynb cells 9-14) run unchanged against the CUDA path:

    import calico_b200.api as calico
    camera = calico.Camera(); camera.SetModel(calico.CameraIntrinsicsModel.kOpenCv5); camera.AddMeasurements(...)
    optimizer = calico.BatchOptimizer(); optimizer.AddSensor(camera); ...; summary = optimizer.Optimize()

Quaternions cross this API in the order w, x, y, z like the reference's Python layer (typedefs.h:69-81); the C ABI uses Eigen's
x, y, z, w. Every numerical step (fit, projection, optimisation, residuals) runs on the device; there is no CPU fallback.
"""
from __future__ import annotations

import enum
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from . import _capi
from . import spline as _sp


class StatusCode(enum.IntEnum):
    kOk = 0
    kInvalidArgument = 3


class LossFunctionType(enum.IntEnum):      # optimization_utils.h:15-22
    kNone = 0
    kHuber = 1
    kCauchy = 2


class CameraIntrinsicsModel(enum.IntEnum):   # camera_models.h:16-33
    kNone = 0
    kOpenCv5 = 1
    kOpenCv8 = 2
    kKannalaBrandt = 3
    kDoubleSphere = 4
    kFieldOfView = 5
    kUnifiedCamera = 6
    kExtendedUnifiedCamera = 7


class GyroscopeIntrinsicsModel(enum.IntEnum):   # gyroscope_models.h:16-25
    kNone = 0
    kGyroscopeScaleOnly = 1
    kGyroscopeScaleAndBias = 2
    kGyroscopeVectorNav = 3


class AccelerometerIntrinsicsModel(enum.IntEnum):   # accelerometer_models.h:16-25
    kNone = 0
    kAccelerometerScaleOnly = 1
    kAccelerometerScaleAndBias = 2
    kAccelerometerVectorNav = 3


class _ParamTable:
    """model -> number of intrinsics, defined ONCE behind the C ABI (cb2_num_intrinsics: camera_models.h:79,231,395,596,716,848,961;
    accelerometer_models.h / gyroscope_models.h). Needs the library, not a GPU."""

    def __init__(self, kind: int):
        self._kind = kind

    def _n(self, model: int) -> int:
        return _capi.num_intrinsics(self._kind, int(model), **({"lib_path": _LIB} if _LIB else {}))

    def __contains__(self, model) -> bool:
        return self._n(model) >= 0

    def __getitem__(self, model) -> int:
        n = self._n(model)
        if n < 0:
            raise KeyError(model)
        return n


_CAMERA_PARAMS = _ParamTable(0)
_IMU_PARAMS = _ParamTable(1)
kLandmarkFrameId = -1                      # camera.h: landmarks are observed with model_id = -1


_LIB = None   # None = the in-tree CUDA library; tests point this at the SIMT-emulation build of the same kernel sources


def set_library(path):
    """Selects the shared library behind the C ABI (default: calico_b200/libcalico_b200.so)."""
    global _LIB
    _LIB = path


def _new_api() -> _capi.CApi:
    return _capi.CApi(_LIB) if _LIB else _capi.CApi()


def _err(msg):
    return RuntimeError("Error: " + msg)


class Pose3d:
    """typedefs.h:39-153; `rotation` is w, x, y, z."""

    def __init__(self, other: Optional["Pose3d"] = None):
        self._q_xyzw = np.array([0.0, 0.0, 0.0, 1.0]) if other is None else other._q_xyzw.copy()
        self._t = np.zeros(3) if other is None else other._t.copy()

    @property
    def rotation(self):
        return np.array([self._q_xyzw[3], self._q_xyzw[0], self._q_xyzw[1], self._q_xyzw[2]])

    @rotation.setter
    def rotation(self, wxyz):
        w, x, y, z = np.asarray(wxyz, dtype=np.float64)
        self._q_xyzw = np.array([x, y, z, w])

    @property
    def translation(self):
        return self._t.copy()

    @translation.setter
    def translation(self, t):
        self._t = np.asarray(t, dtype=np.float64).reshape(3).copy()


@dataclass(frozen=True)
class CameraObservationId:                 # camera.h:24-50
    stamp: float = 0.0
    image_id: int = 0
    model_id: int = 0
    feature_id: int = 0


@dataclass
class CameraMeasurement:
    pixel: np.ndarray = field(default_factory=lambda: np.zeros(2))
    id: CameraObservationId = field(default_factory=CameraObservationId)


@dataclass(frozen=True)
class GyroscopeObservationId:
    stamp: float = 0.0
    sequence: int = 0


AccelerometerObservationId = GyroscopeObservationId


@dataclass
class GyroscopeMeasurement:
    measurement: np.ndarray = field(default_factory=lambda: np.zeros(3))
    id: GyroscopeObservationId = field(default_factory=GyroscopeObservationId)


AccelerometerMeasurement = GyroscopeMeasurement


@dataclass
class Landmark:                            # world_model.h:21-39
    point: np.ndarray = field(default_factory=lambda: np.zeros(3))
    id: int = 0
    point_is_constant: bool = True


@dataclass
class RigidBody:                           # world_model.h:41-69
    model_definition: Dict[int, np.ndarray] = field(default_factory=dict)
    T_world_rigidbody: Pose3d = field(default_factory=Pose3d)
    id: int = 0
    world_pose_is_constant: bool = True
    model_definition_is_constant: bool = True


class WorldModel:                          # world_model.{h,cpp}
    def __init__(self):
        self._landmarks: Dict[int, Landmark] = {}
        self._rigidbodies: Dict[int, RigidBody] = {}
        self._gravity = np.array([0.0, 0.0, -9.80665])

    def AddLandmark(self, landmark: Landmark):
        if landmark.id in self._landmarks:
            raise _err(f"Landmark {landmark.id} already exists in world model.")
        self._landmarks[landmark.id] = landmark

    def AddRigidBody(self, rigidbody: RigidBody):
        if rigidbody.id in self._rigidbodies:
            raise _err(f"Rigid body {rigidbody.id} already exists in world model.")
        self._rigidbodies[rigidbody.id] = rigidbody

    def SetGravity(self, g):
        self._gravity = np.asarray(g, dtype=np.float64).reshape(3).copy()

    def GetGravity(self):
        return self._gravity.copy()

    def EnableGravityEstimation(self, enable: bool):
        pass                                # a no-op in the reference as well (world_model.cpp:79-81)


class Trajectory:                          # trajectory.{h,cpp}
    kDefaultKnotFrequency = 10.0
    kDefaultSplineOrder = 6

    def __init__(self):
        self._spline: Optional[_sp.Spline] = None
        self._poses: Dict[float, Pose3d] = {}

    def FitSpline(self, poses: Dict[float, Pose3d], knot_frequency: float = kDefaultKnotFrequency, spline_order: int = kDefaultSplineOrder):
        """trajectory.cpp:14-49 on the device (cb2_fit_trajectory)."""
        self._poses = dict(poses)
        stamps = np.array(list(poses.keys()), dtype=np.float64)
        q = np.array([p._q_xyzw for p in poses.values()], dtype=np.float64).reshape(-1, 4)
        t = np.array([p._t for p in poses.values()], dtype=np.float64).reshape(-1, 3)
        try:
            knots, ctrl = _capi.fit_trajectory(stamps, q, t, knot_frequency, spline_order, **({"lib_path": _LIB} if _LIB else {}))
        except _capi.CalicoError as e:
            raise RuntimeError(str(e)) from None
        self._spline = _sp.Spline(spline_order, knots, ctrl)

    def Interpolate(self, stamps) -> List[Pose3d]:
        """trajectory.cpp:95-112: poses at `stamps` (axis-angle -> quaternion like ceres::AngleAxisToQuaternion, trajectory.h:98)."""
        if self._spline is None:
            raise _err("Cannot interpolate. Spline has not been fitted.")
        try:
            vals = self._spline.evaluate(np.asarray(stamps, dtype=np.float64), 0)
        except ValueError as e:
            raise _err(str(e)) from None
        out = []
        for v in vals:
            p = Pose3d()
            p._q_xyzw = _sp.angle_axis_to_quat_xyzw(v[:3])[0]
            p._t = v[3:].copy()
            out.append(p)
        return out


class Sensor:
    """sensor_base.h:22-102 as exposed to Python (calico.cpp:54-281)."""
    _KIND = -1
    _PARAMS: Dict[int, int] = {}
    _LOWER = "sensor"

    def __init__(self):
        self._name = ""
        self._T = Pose3d()
        self._model = 0
        self._intrinsics = np.zeros(0)
        self._latency = 0.0
        self._sigma = 1.0
        self._loss, self._loss_scale = LossFunctionType.kNone, 1.0
        self._en_extr = self._en_intr = self._en_lat = False
        self._measurements: Dict[object, object] = {}
        self._residuals: Dict[object, np.ndarray] = {}

    def SetName(self, name: str):
        self._name = name

    def GetName(self) -> str:
        return self._name

    def SetExtrinsics(self, T: Pose3d):
        self._T = Pose3d(T)

    def GetExtrinsics(self) -> Pose3d:
        return Pose3d(self._T)

    def SetIntrinsics(self, intrinsics):
        intrinsics = np.asarray(intrinsics, dtype=np.float64).reshape(-1)
        if self._model == 0:
            raise _err(f"{self._LOWER.capitalize()} model has not been set!")
        want = self._PARAMS[self._model]
        if intrinsics.size != want:
            raise _err(f"Tried to set intrinsics of size {intrinsics.size} for {self._LOWER} {self._name}. Expected intrinsics size of {want}")
        self._intrinsics = intrinsics.copy()

    def GetIntrinsics(self):
        return self._intrinsics.copy()

    def SetLatency(self, latency: float):
        self._latency = float(latency)

    def GetLatency(self) -> float:
        return self._latency

    def EnableExtrinsicsEstimation(self, enable: bool):
        self._en_extr = bool(enable)

    def EnableIntrinsicsEstimation(self, enable: bool):
        self._en_intr = bool(enable)

    def EnableLatencyEstimation(self, enable: bool):
        self._en_lat = bool(enable)

    def SetModel(self, model):
        model = int(model)
        if model not in self._PARAMS:
            raise _err(f"Could not create {self._LOWER} model for type {model}. It is likely not yet implemented.")
        self._model = model
        self._intrinsics = np.zeros(self._PARAMS[model])

    def SetLossFunction(self, loss, scale: float = 1.0):
        self._loss, self._loss_scale = LossFunctionType(int(loss)), float(scale)

    def SetMeasurementNoise(self, sigma: float):
        if sigma <= 0.0:
            raise _err("Sigma must be greater than 0.")
        self._sigma = float(sigma)

    def AddMeasurement(self, m):
        if m.id in self._measurements:
            raise _err(self._redundant(m.id))
        self._measurements[m.id] = m

    def AddMeasurements(self, ms):
        msgs = []
        for m in ms:
            try:
                self.AddMeasurement(m)
            except RuntimeError as e:
                msgs.append(str(e)[len("Error: "):])
        if msgs:
            raise _err("\n".join(msgs) + "\n")

    def NumberOfMeasurements(self) -> int:
        return len(self._measurements)

    def _redundant(self, mid):
        return f"Tried to add redundant measurement - stamp: {mid.stamp}, sequence: {mid.sequence}"

    # ---- C ABI plumbing ----
    def _add_sensor(self, api: _capi.CApi, probe: bool = False):
        if self._model == 0:
            raise _capi.CalicoError(_capi.FAILED_PRECONDITION, "Cannot add sensor parameters. Model is not yet defined.")
        if probe:
            return api.add_sensor(self._KIND, self._model, self._name, self._intrinsics, self._T._q_xyzw, self._T._t, 0.0, 1.0, 0, 1.0, 0, 0, 0)
        return api.add_sensor(self._KIND, self._model, self._name, self._intrinsics, self._T._q_xyzw, self._T._t, self._latency, self._sigma,
                              int(self._loss), self._loss_scale, int(self._en_intr), int(self._en_extr), int(self._en_lat))

    def _read_back(self, api: _capi.CApi, sid: int):
        intr, q, t, lat = api.get_sensor(sid)
        self._intrinsics, self._latency = np.array(intr), float(lat)
        self._T._q_xyzw, self._T._t = np.array(q), np.array(t)
        r, valid = api.get_residuals(sid)
        self._residuals = {mid: r[i].copy() for i, mid in enumerate(self._measurements) if valid[i]}


def _push_world(api: _capi.CApi, trajectory: Trajectory, world_model: WorldModel, landmarks_as_body: bool = False):
    """landmarks_as_body (Camera::Project only, camera.cpp:169-184): the landmarks ride along as a constant pseudo rigid body with identity
    pose and id kLandmarkFrameId. Optimize never pushes them: the reference rejects landmark observations (camera.cpp:125-131) and the
    device path returns the same FailedPrecondition for their model_id."""
    if trajectory._spline is None:
        raise _capi.CalicoError(_capi.FAILED_PRECONDITION, "Trajectory has not been set.")
    api.set_trajectory(trajectory._spline.k, trajectory._spline.knots, trajectory._spline.ctrl)
    api.set_gravity(world_model._gravity)
    for rid, body in world_model._rigidbodies.items():
        ids = np.array(list(body.model_definition.keys()), dtype=np.int32)
        pts = np.array([np.asarray(p, dtype=np.float64) for p in body.model_definition.values()]).reshape(-1, 3)
        api.add_rigid_body(rid, body.T_world_rigidbody._q_xyzw, body.T_world_rigidbody._t, ids, pts, body.world_pose_is_constant,
                           body.model_definition_is_constant)
    if landmarks_as_body and world_model._landmarks:
        if kLandmarkFrameId in world_model._rigidbodies:
            raise _capi.CalicoError(_capi.INVALID_ARGUMENT, "Rigid body id -1 is reserved for landmarks (kLandmarkFrameId).")
        ids = np.array(list(world_model._landmarks.keys()), dtype=np.int32)
        pts = np.array([np.asarray(l.point, dtype=np.float64) for l in world_model._landmarks.values()]).reshape(-1, 3)
        api.add_rigid_body(kLandmarkFrameId, np.array([0.0, 0.0, 0.0, 1.0]), np.zeros(3), ids, pts, True, True)


class Camera(Sensor):                      # camera.{h,cpp}
    _KIND, _PARAMS, _LOWER = 0, _CAMERA_PARAMS, "camera"

    def __init__(self):
        super().__init__()
        self._outliers = set()

    def GetModel(self):
        return CameraIntrinsicsModel(self._model)

    def _redundant(self, mid):
        return f"Tried to add redundant measurement - Image id: {mid.image_id}, model id: {mid.model_id}, feature id: {mid.feature_id}"

    def GetMeasurementIdToMeasurement(self):
        return dict(self._measurements)

    def GetMeasurementResidualPairs(self):   # camera.cpp:258-279
        if not self._measurements:
            raise _err("Measurements are empty. Nothing to return.")
        return [(self._measurements[mid], r) for mid, r in self._residuals.items()]

    def MarkOutlierById(self, mid: CameraObservationId):   # camera.cpp:281-291
        if mid not in self._measurements:
            raise _err("Attempted to add id that is not within the measurement set.")
        self._outliers.add(mid)

    def MarkOutliersById(self, ids):
        for mid in ids:
            self.MarkOutlierById(mid)

    def ClearOutliersList(self):
        self._outliers.clear()

    def _add_to_problem(self, api: _capi.CApi):
        sid = self._add_sensor(api)
        ms = list(self._measurements.values())
        if ms:
            api.add_camera_observations(sid, [m.id.stamp for m in ms], [m.id.image_id for m in ms], [m.id.model_id for m in ms],
                                        [m.id.feature_id for m in ms], np.array([m.pixel for m in ms]).reshape(-1, 2),
                                        np.array([m.id in self._outliers for m in ms], dtype=np.uint8))
        return sid

    def Project(self, interp_times, sensorrig_trajectory: Trajectory, world_model: WorldModel) -> List[CameraMeasurement]:
        """camera.cpp:155-208 through the forward mode of the camera kernel (zero measurement, unit sigma, zero latency)."""
        api = _new_api()
        try:
            _push_world(api, sensorrig_trajectory, world_model, landmarks_as_body=True)
            sid = self._add_sensor(api, probe=True)
            times = np.asarray(interp_times, dtype=np.float64)
            stamp, image_id, model_id, feature_id = [], [], [], []
            for i, t in enumerate(times):
                for lid in world_model._landmarks:                  # camera.cpp:169-184: landmarks first, model_id = kLandmarkFrameId
                    stamp.append(t); image_id.append(i); model_id.append(kLandmarkFrameId); feature_id.append(lid)
                for rid, body in world_model._rigidbodies.items():
                    for pid in body.model_definition:
                        stamp.append(t); image_id.append(i); model_id.append(rid); feature_id.append(pid)
            if not stamp:
                return []
            api.add_camera_observations(sid, stamp, image_id, model_id, feature_id, np.zeros((len(stamp), 2)))
            r, _, flags = api.evaluate_sensor(sid, want_jac=False, raw_flags=True)
        except _capi.CalicoError as e:
            raise RuntimeError(str(e)) from None
        finally:
            api.close()
        # flags == 1: projected and in front of the image plane. Points with z <= 0 are skipped (camera.cpp:172-174,186-188) for EVERY model,
        # also those (DoubleSphere, Unified, ExtendedUnified) whose ProjectPoint accepts them.
        return [CameraMeasurement(-r[i], CameraObservationId(stamp[i] + self._latency, image_id[i], model_id[i], feature_id[i]))
                for i in range(len(stamp)) if flags[i] == 1]


class _Imu(Sensor):
    def _add_to_problem(self, api: _capi.CApi):
        sid = self._add_sensor(api)
        ms = list(self._measurements.values())
        if ms:
            api.add_imu_observations(sid, [m.id.stamp for m in ms], [m.id.sequence for m in ms], np.array([m.measurement for m in ms]).reshape(-1, 3))
        return sid

    def Project(self, interp_times, sensorrig_trajectory: Trajectory, world_model: WorldModel):
        """gyroscope.cpp:56-82 / accelerometer.cpp:76-123 through the forward mode of the IMU kernels."""
        api = _new_api()
        try:
            _push_world(api, sensorrig_trajectory, world_model)
            sid = self._add_sensor(api, probe=True)
            times = np.asarray(interp_times, dtype=np.float64)
            if times.size == 0:
                return []
            api.add_imu_observations(sid, times, np.arange(times.size), np.zeros((times.size, 3)))
            r, _, valid = api.evaluate_sensor(sid, want_jac=False)
        except _capi.CalicoError as e:
            raise RuntimeError(str(e)) from None
        finally:
            api.close()
        if not np.all(valid):
            raise _err(f"Failed to project {self._LOWER} measurement.")
        return [GyroscopeMeasurement(-r[i], GyroscopeObservationId(float(times[i]) + self._latency, i)) for i in range(times.size)]


class Gyroscope(_Imu):                     # gyroscope.{h,cpp}
    _KIND, _PARAMS, _LOWER = 1, _IMU_PARAMS, "gyroscope"

    def GetModel(self):
        return GyroscopeIntrinsicsModel(self._model)


class Accelerometer(_Imu):                 # accelerometer.{h,cpp}
    _KIND, _PARAMS, _LOWER = 2, _IMU_PARAMS, "accelerometer"

    def GetModel(self):
        return AccelerometerIntrinsicsModel(self._model)


class SolverOptions:
    """The ceres::Solver::Options fields exposed at calico.cpp:378-394."""

    def __init__(self):
        self.minimizer_type = "TRUST_REGION"
        self.max_num_iterations = 50
        self.num_threads = 1
        self.function_tolerance = 1e-8            # batch_optimizer.cpp:14
        self.gradient_tolerance = 1e-10
        self.parameter_tolerance = 1e-10          # batch_optimizer.cpp:15
        self.linear_solver_type = "DENSE_SCHUR"   # batch_optimizer.cpp:12; the device path always eliminates the control points
        self.preconditioner_type = "JACOBI"
        self.minimizer_progress_to_stdout = True  # batch_optimizer.cpp:13


def DefaultSolverOptions() -> SolverOptions:
    return SolverOptions()


class Summary:
    """The ceres::Solver::Summary members exposed at calico.cpp:352-375 (+ termination_type, read by batch_optimizer_test.cpp:186)."""

    def __init__(self, s: _capi.Summary):
        for name in ("initial_cost", "final_cost", "num_residual_blocks", "num_residuals", "num_parameter_blocks", "num_parameters",
                     "num_parameter_blocks_reduced", "num_parameters_reduced", "num_effective_parameters_reduced", "num_residual_blocks_reduced",
                     "num_residuals_reduced", "termination_type", "num_successful_steps", "num_unsuccessful_steps", "num_iterations", "total_time"):
            setattr(self, name, getattr(s, name))
        self.message = s.message.decode() if isinstance(s.message, bytes) else str(s.message)

    def IsSolutionUsable(self) -> bool:
        return self.termination_type in (_capi.CONVERGENCE, _capi.NO_CONVERGENCE)

    def BriefReport(self) -> str:
        term = {0: "CONVERGENCE", 1: "NO_CONVERGENCE", 2: "FAILURE"}[self.termination_type]
        return (f"calico_b200 Report: Iterations: {self.num_iterations}, Initial cost: {self.initial_cost:e}, Final cost: {self.final_cost:e}, "
                f"Termination: {term}")

    def FullReport(self) -> str:
        return (f"{self.BriefReport()}\nResidual blocks {self.num_residual_blocks}, residuals {self.num_residuals}, parameter blocks "
                f"{self.num_parameter_blocks} (reduced {self.num_parameter_blocks_reduced}), parameters {self.num_parameters} (reduced "
                f"{self.num_parameters_reduced})\nSuccessful steps {self.num_successful_steps}, unsuccessful steps {self.num_unsuccessful_steps}, "
                f"total time {self.total_time:.6f} s\n{self.message}")


class BatchOptimizer:                      # batch_optimizer.{h,cpp}, calico.cpp:400-424
    def __init__(self):
        self._sensors: List[Sensor] = []
        self._trajectory: Optional[Trajectory] = None
        self._world_model: Optional[WorldModel] = None

    def AddSensor(self, sensor: Sensor):
        self._sensors.append(sensor)

    def AddTrajectory(self, trajectory: Trajectory):
        self._trajectory = trajectory

    def AddWorldModel(self, world_model: WorldModel):
        self._world_model = world_model

    def Optimize(self, options: Optional[SolverOptions] = None) -> Summary:
        """batch_optimizer.cpp:53-81: new problem per call, warm start from the objects, parameters written back in place, residuals refreshed."""
        options = options or DefaultSolverOptions()
        if self._trajectory is None or self._world_model is None:
            raise _err("Trajectory and world model must be added before optimizing.")
        api = _new_api()
        try:
            _push_world(api, self._trajectory, self._world_model)
            ids = []
            for s in self._sensors:
                s._residuals = {}                     # ClearResidualInfo, batch_optimizer.cpp:63
                ids.append(s._add_to_problem(api))
            o = _capi.Options(max_num_iterations=options.max_num_iterations, function_tolerance=options.function_tolerance,
                              gradient_tolerance=options.gradient_tolerance, parameter_tolerance=options.parameter_tolerance,
                              num_threads=options.num_threads, minimizer_progress_to_stdout=int(bool(options.minimizer_progress_to_stdout)))
            failure = None
            try:
                summ, _ = api.optimize(o)
            except _capi.CalicoError as e:            # e.g. kInternal from the residual refresh: Ceres has already mutated the parameters
                failure, summ = e, api.last_summary
            self._trajectory._spline.ctrl[...] = api.get_trajectory()
            for rid, body in self._world_model._rigidbodies.items():      # freed world-model blocks are mutated in place too (world_model.cpp:52-70)
                if body.world_pose_is_constant and body.model_definition_is_constant:
                    continue
                q, t, pts = api.get_rigid_body(rid, len(body.model_definition))
                body.T_world_rigidbody._q_xyzw, body.T_world_rigidbody._t = np.array(q), np.array(t)
                for pid, pt in zip(list(body.model_definition.keys()), pts):
                    body.model_definition[pid] = np.array(pt)
            for s, sid in zip(self._sensors, ids):
                try:
                    s._read_back(api, sid)
                except _capi.CalicoError as e:
                    failure = failure or e
            if failure is not None:
                raise RuntimeError(str(failure))
            return Summary(summ)
        except _capi.CalicoError as e:
            raise RuntimeError(str(e)) from None
        finally:
            api.close()
