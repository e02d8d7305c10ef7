"""Plain-data description of one calibration problem, and its hand-over to a C-ABI handle.

The reference keeps this state spread over `Trajectory`, `WorldModel` and `Sensor` objects that Ceres mutates in place
(calico/batch_optimizer.cpp:53-81); here it is gathered into SoA numpy arrays once, which is what `cb2_*` consumes.
"""
from __future__ import annotations

import copy
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import _capi
from .spline import Spline

CAMERA, GYROSCOPE, ACCELEROMETER = _capi.CAMERA, _capi.GYROSCOPE, _capi.ACCELEROMETER
PARALLEL_OBS = 100_000   # observations from which push() / residuals() use one thread per sensor

# CameraIntrinsicsModel (camera_models.h:16-33) → number of intrinsics.
CAMERA_NUM_PARAMS = {1: 8, 2: 11, 3: 7, 4: 5, 5: 4, 6: 4, 7: 5}
# {Accelerometer,Gyroscope}IntrinsicsModel (accelerometer_models.h:16-25).
IMU_NUM_PARAMS = {1: 1, 2: 4, 3: 12}


@dataclass
class RigidBodySpec:
    id: int
    q_xyzw: np.ndarray
    t: np.ndarray
    feature_ids: np.ndarray
    pts: np.ndarray
    pose_const: bool = True
    model_const: bool = True


@dataclass
class SensorSpec:
    kind: int
    model: int
    name: str
    intr: np.ndarray
    q_xyzw: np.ndarray = field(default_factory=lambda: np.array([0.0, 0.0, 0.0, 1.0]))
    t: np.ndarray = field(default_factory=lambda: np.zeros(3))
    latency: float = 0.0
    sigma: float = 1.0
    loss_type: int = 0
    loss_scale: float = 1.0
    en_intr: bool = False
    en_extr: bool = False
    en_lat: bool = False
    # observations
    stamp: np.ndarray = field(default_factory=lambda: np.zeros(0))
    meas: np.ndarray = field(default_factory=lambda: np.zeros((0, 2)))      # pixel (n,2) or imu (n,3)
    image_id: Optional[np.ndarray] = None
    model_id: Optional[np.ndarray] = None
    feature_id: Optional[np.ndarray] = None
    seq: Optional[np.ndarray] = None
    outlier: Optional[np.ndarray] = None

    @property
    def m(self):
        return 2 if self.kind == CAMERA else 3

    @property
    def n_obs(self):
        return int(np.asarray(self.stamp).size)


@dataclass
class ProblemSpec:
    spline: Spline
    gravity: np.ndarray = field(default_factory=lambda: np.array([0.0, 0.0, -9.80665]))  # world_model.h:78
    bodies: List[RigidBodySpec] = field(default_factory=list)
    sensors: List[SensorSpec] = field(default_factory=list)

    def clone(self) -> "ProblemSpec":
        return copy.deepcopy(self)

    def push(self, api: "_capi.CApi"):
        """Hands the whole problem to a C-ABI handle; returns the sensor ids in order."""
        api.set_trajectory(self.spline.k, self.spline.knots, self.spline.ctrl)
        api.set_gravity(self.gravity)
        for b in self.bodies:
            api.add_rigid_body(b.id, b.q_xyzw, b.t, b.feature_ids, b.pts, b.pose_const, b.model_const)
        ids = []
        for s in self.sensors:
            sid = api.add_sensor(s.kind, s.model, s.name, s.intr, s.q_xyzw, s.t, s.latency, s.sigma, s.loss_type, s.loss_scale,
                                 s.en_intr, s.en_extr, s.en_lat)
            ids.append(sid)

        def add_obs(pair):
            s, sid = pair
            if s.n_obs == 0:
                return
            if s.kind == CAMERA:
                api.add_camera_observations(sid, s.stamp, s.image_id, s.model_id, s.feature_id, s.meas, s.outlier)
            else:
                seq = s.seq if s.seq is not None else np.arange(s.n_obs)
                api.add_imu_observations(sid, s.stamp, seq, s.meas)

        # The observation arrays of DIFFERENT sensors may be handed over concurrently (include/calico_b200.h); the calls are memory-bound
        # copies that release the GIL, so large problems use one thread per sensor.
        pairs = list(zip(self.sensors, ids))
        if len(pairs) > 1 and sum(s.n_obs for s in self.sensors) >= PARALLEL_OBS:
            with ThreadPoolExecutor(max_workers=min(len(pairs), 16)) as ex:
                list(ex.map(add_obs, pairs))
        else:
            for pair in pairs:
                add_obs(pair)
        return ids

    def residuals(self, api: "_capi.CApi", ids=None):
        """Residuals + validity flags of every sensor, in the caller's observation order (Sensor::UpdateResiduals, camera.cpp:70-80);
        the per-sensor copies run concurrently on large problems."""
        ids = ids if ids is not None else list(range(len(self.sensors)))
        if len(ids) > 1 and sum(s.n_obs for s in self.sensors) >= PARALLEL_OBS:
            with ThreadPoolExecutor(max_workers=min(len(ids), 16)) as ex:
                return list(ex.map(api.get_residuals, ids))
        return [api.get_residuals(sid) for sid in ids]

    def pull(self, api: "_capi.CApi", ids=None):
        """Write-back of the optimised state (the reference mutates the user's objects in place, camera.cpp:98-101)."""
        ids = ids if ids is not None else list(range(len(self.sensors)))
        self.spline.ctrl[...] = api.get_trajectory()
        for b in self.bodies:
            if not (b.pose_const and b.model_const):
                b.q_xyzw, b.t, b.pts = api.get_rigid_body(b.id, np.asarray(b.pts).reshape(-1, 3).shape[0])
        for s, sid in zip(self.sensors, ids):
            intr, q, t, lat = api.get_sensor(sid)
            s.intr, s.q_xyzw, s.t, s.latency = intr, q, t, lat

    def counts(self):
        """(residual blocks, scalar residuals) over non-outlier observations."""
        nb = nr = 0
        for s in self.sensors:
            n = s.n_obs - (int(np.count_nonzero(s.outlier)) if s.outlier is not None else 0)
            nb += n
            nr += n * s.m
        return nb, nr

    def jacobian_bytes(self) -> float:
        """Algorithmic bytes of one residual+Jacobian sweep, SURVEY §8(d): per block obs_read + 8*m*w + 8*m, with w the
        tangent columns of the non-constant blocks (control points are always free)."""
        total = 0.0
        for s in self.sensors:
            n = s.n_obs - (int(np.count_nonzero(s.outlier)) if s.outlier is not None else 0)
            w = 6 * self.spline.k
            if s.en_intr:
                w += int(np.asarray(s.intr).size)
            if s.en_extr:
                w += 3 if s.kind == GYROSCOPE else 6   # gyro translation columns are structurally zero
            if s.en_lat:
                w += 1
            obs = 32 if s.kind == CAMERA else 40
            total += n * (obs + 8 * s.m * w + 8 * s.m)
        return total
