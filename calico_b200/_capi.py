"""ctypes binding of the C ABI declared in include/calico_b200.h.

`CApi(lib_path, prefix)` binds every entry point `<prefix><name>`; the product library uses prefix ``cb2_``.
The binder is prefix-generic only so that the test-side CPU oracle (which deliberately exposes the same call
shapes) can be driven by the same problem description; this package never loads anything but its own
``libcalico_b200.so`` and fails loudly when that library is missing (there is no CPU fallback).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcalico_b200.so")

OK, INVALID_ARGUMENT, FAILED_PRECONDITION, UNIMPLEMENTED, INTERNAL = 0, 3, 9, 12, 13
CAMERA, GYROSCOPE, ACCELEROMETER = 0, 1, 2
CONVERGENCE, NO_CONVERGENCE, FAILURE = 0, 1, 2


class Options(C.Structure):
    """cb2_options (include/calico_b200.h); defaults = calico::DefaultSolverOptions (batch_optimizer.cpp:10-17)."""
    _fields_ = [
        ("max_num_iterations", C.c_int32),
        ("function_tolerance", C.c_double),
        ("gradient_tolerance", C.c_double),
        ("parameter_tolerance", C.c_double),
        ("initial_trust_region_radius", C.c_double),
        ("max_trust_region_radius", C.c_double),
        ("min_trust_region_radius", C.c_double),
        ("min_relative_decrease", C.c_double),
        ("min_lm_diagonal", C.c_double),
        ("max_lm_diagonal", C.c_double),
        ("max_num_consecutive_invalid_steps", C.c_int32),
        ("jacobi_scaling", C.c_int32),
        ("num_threads", C.c_int32),
        ("minimizer_progress_to_stdout", C.c_int32),
        ("linear_solver", C.c_int32),
        ("use_cuda_graph", C.c_int32),
    ]

    def __init__(self, **kw):
        super().__init__()
        self.max_num_iterations = 50
        self.function_tolerance = 1e-8
        self.gradient_tolerance = 1e-10
        self.parameter_tolerance = 1e-10
        self.initial_trust_region_radius = 1e4
        self.max_trust_region_radius = 1e16
        self.min_trust_region_radius = 1e-32
        self.min_relative_decrease = 1e-3
        self.min_lm_diagonal = 1e-6
        self.max_lm_diagonal = 1e32
        self.max_num_consecutive_invalid_steps = 5
        self.jacobi_scaling = 1
        self.num_threads = 1
        self.minimizer_progress_to_stdout = 0
        self.linear_solver = 0
        self.use_cuda_graph = 0
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError(k)
            setattr(self, k, v)


class Iteration(C.Structure):
    _fields_ = [
        ("iteration", C.c_int32),
        ("cost", C.c_double), ("cost_change", C.c_double), ("gradient_max_norm", C.c_double),
        ("gradient_norm", C.c_double), ("step_norm", C.c_double), ("relative_decrease", C.c_double),
        ("trust_region_radius", C.c_double),
        ("step_is_valid", C.c_int32), ("step_is_successful", C.c_int32),
        ("iteration_time", C.c_double),
    ]


class Summary(C.Structure):
    _fields_ = [
        ("termination_type", C.c_int32),
        ("initial_cost", C.c_double), ("final_cost", C.c_double), ("fixed_cost", C.c_double),
        ("num_successful_steps", C.c_int32), ("num_unsuccessful_steps", C.c_int32), ("num_iterations", C.c_int32),
        ("num_parameter_blocks", C.c_int32), ("num_parameters", C.c_int32), ("num_effective_parameters", C.c_int32),
        ("num_residual_blocks", C.c_int32), ("num_residuals", C.c_int32),
        ("num_parameter_blocks_reduced", C.c_int32), ("num_parameters_reduced", C.c_int32),
        ("num_effective_parameters_reduced", C.c_int32), ("num_residual_blocks_reduced", C.c_int32),
        ("num_residuals_reduced", C.c_int32),
        ("jacobian_time", C.c_double), ("linear_solver_time", C.c_double), ("total_time", C.c_double),
        ("message", C.c_char * 256),
    ]

    def BriefReport(self) -> str:  # ceres::Solver::Summary::BriefReport shape (calico.cpp:353)
        term = {0: "CONVERGENCE", 1: "NO_CONVERGENCE", 2: "FAILURE"}.get(self.termination_type, "?")
        return (f"Ceres Solver Report: Iterations: {self.num_iterations}, Initial cost: {self.initial_cost:e}, "
                f"Final cost: {self.final_cost:e}, Termination: {term}")

    def IsSolutionUsable(self) -> bool:
        return self.termination_type in (CONVERGENCE, NO_CONVERGENCE)


class Stats(C.Structure):
    _fields_ = [
        ("kernel_launches", C.c_int64), ("jacobian_sweeps", C.c_int64), ("jacobian_blocks", C.c_int64),
        ("jacobian_kernel_ms", C.c_double), ("jacobian_bytes", C.c_double),
        ("normal_eq_ms", C.c_double), ("schur_ms", C.c_double), ("cost_eval_ms", C.c_double),
        ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
        ("lm_loop_ms", C.c_double), ("lm_iterations", C.c_int64),
        ("camera_kernel_ms", C.c_double), ("camera_kernel_bytes", C.c_double), ("camera_kernel_gram_bytes", C.c_double),
    ]


class CalicoError(RuntimeError):
    """Raised for a non-OK status; message formatted as the reference's pybind layer does (calico.cpp:417-421)."""

    def __init__(self, code: int, msg: str):
        super().__init__("Error: " + msg)
        self.code = code


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_up = C.POINTER(C.c_uint8)


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def _u(a):
    return None if a is None else a.ctypes.data_as(_up)


def f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a if shape is None else a.reshape(shape)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class CApi:
    """Thin object wrapper over one problem handle of a library exposing the calico_b200 C ABI shape."""

    def __init__(self, lib_path: str = LIB_PATH, prefix: str = "cb2_", options_cls=Options):
        if not os.path.exists(lib_path):
            raise ImportError(
                f"{lib_path} not found: the CUDA extension has not been built. Run "
                f"`python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback).")
        self.lib = C.CDLL(lib_path)
        self.prefix = prefix
        self.options_cls = options_cls
        self._bind()
        h = C.c_void_p()
        self._check(self._f("problem_create")(C.byref(h)))
        self.h = h
        self.n_intr = {}
        self.n_obs = {}
        self.kind = {}
        self.k = 6
        self.n_cp = 0

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    def _bind(self):
        vp = C.c_void_p
        sig = {
            "problem_create": ([C.POINTER(vp)], C.c_int),
            "problem_destroy": ([vp], None),
            "last_error": ([vp], C.c_char_p),
            "set_trajectory": ([vp, C.c_int, C.c_int, _dp, C.c_int, _dp], C.c_int),
            "set_gravity": ([vp, _dp], C.c_int),
            "add_rigid_body": ([vp, C.c_int, _dp, _dp, C.c_int, _ip, _dp, C.c_int, C.c_int], C.c_int),
            "add_sensor": ([vp, C.c_int, C.c_int, C.c_char_p, C.c_int, _dp, _dp, _dp, C.c_double, C.c_double, C.c_int,
                            C.c_double, C.c_int, C.c_int, C.c_int, _ip], C.c_int),
            "add_camera_observations": ([vp, C.c_int, C.c_int, _dp, _ip, _ip, _ip, _dp, _up], C.c_int),
            "add_imu_observations": ([vp, C.c_int, C.c_int, _dp, _ip, _dp], C.c_int),
            "optimize": ([vp, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, _ip], C.c_int),
            "evaluate_sensor": ([vp, C.c_int, _dp, _dp, _up], C.c_int),
            "cost": ([vp, _dp, _ip], C.c_int),
            "get_sensor": ([vp, C.c_int, _dp, _dp, _dp, _dp], C.c_int),
            "get_trajectory": ([vp, _dp], C.c_int),
            "get_rigid_body": ([vp, C.c_int, _dp, _dp, _dp], C.c_int),
            "get_residuals": ([vp, C.c_int, _dp, _up], C.c_int),
        }
        if self.prefix == "cb2_":
            sig.update({
                "upload": ([vp], C.c_int),
                "reset_parameters": ([vp], C.c_int),
                "stats_reset": ([vp], C.c_int),
                "stats_get": ([vp, C.c_void_p], C.c_int),
                "set_sensor": ([vp, C.c_int, _dp, _dp, _dp, C.c_double], C.c_int),
            })
            self.lib.cb2_comm_unique_id.argtypes = [_up]
            self.lib.cb2_comm_unique_id.restype = C.c_int
            self.lib.cb2_comm_init.argtypes = [vp, C.c_int, C.c_int, _up]
            self.lib.cb2_comm_init.restype = C.c_int
            self.lib.cb2_comm_clone.argtypes = [vp, vp]
            self.lib.cb2_comm_clone.restype = C.c_int
            self.lib.cb2_shard_plan.argtypes = [vp, C.c_int, C.c_int, _ip, _ip, _ip, _ip, _ip]
            self.lib.cb2_shard_plan.restype = C.c_int
            self.lib.cb2_set_device.argtypes = [C.c_int]
            self.lib.cb2_set_device.restype = C.c_int
            self.lib.cb2_version.restype = C.c_char_p
        for name, (argtypes, restype) in sig.items():
            fn = self._f(name)
            fn.argtypes = argtypes
            fn.restype = restype

    def _check(self, rc: int):
        if rc != OK:
            msg = self._f("last_error")(self.h) if getattr(self, "h", None) else b""
            raise CalicoError(rc, (msg or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self._f("problem_destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- assembly ----
    def set_trajectory(self, spline_order, knots, ctrl):
        knots = f64(knots)
        ctrl = f64(ctrl, (-1, 6))
        self.k, self.n_cp = int(spline_order), ctrl.shape[0]
        self._check(self._f("set_trajectory")(self.h, spline_order, knots.size, _d(knots), ctrl.shape[0], _d(ctrl)))

    def set_gravity(self, g):
        g = f64(g)
        self._check(self._f("set_gravity")(self.h, _d(g)))

    def add_rigid_body(self, id, q_xyzw, t, feature_ids, pts, pose_const=True, model_const=True):
        q, t, fid, pts = f64(q_xyzw), f64(t), i32(feature_ids), f64(pts, (-1, 3))
        self._check(self._f("add_rigid_body")(self.h, id, _d(q), _d(t), fid.size, _i(fid), _d(pts), int(pose_const), int(model_const)))

    def add_sensor(self, kind, model, name, intr, q_xyzw, t, latency, sigma, loss_type, loss_scale, en_intr, en_extr, en_lat):
        intr, q, t = f64(intr), f64(q_xyzw), f64(t)
        sid = C.c_int(-1)
        self._check(self._f("add_sensor")(self.h, kind, model, name.encode(), intr.size, _d(intr), _d(q), _d(t), latency, sigma,
                                          loss_type, loss_scale, int(en_intr), int(en_extr), int(en_lat), C.byref(sid)))
        self.n_intr[sid.value] = intr.size
        self.n_obs[sid.value] = 0
        self.kind[sid.value] = kind
        return sid.value

    def add_camera_observations(self, sid, stamp, image_id, model_id, feature_id, pixel, outlier=None):
        stamp, pixel = f64(stamp), f64(pixel, (-1, 2))
        image_id, model_id, feature_id = i32(image_id), i32(model_id), i32(feature_id)
        out = None if outlier is None else np.ascontiguousarray(outlier, dtype=np.uint8)
        self._check(self._f("add_camera_observations")(self.h, sid, stamp.size, _d(stamp), _i(image_id), _i(model_id), _i(feature_id),
                                                       _d(pixel), _u(out)))
        self.n_obs[sid] += stamp.size

    def add_imu_observations(self, sid, stamp, seq, xyz):
        stamp, seq, xyz = f64(stamp), i32(seq), f64(xyz, (-1, 3))
        self._check(self._f("add_imu_observations")(self.h, sid, stamp.size, _d(stamp), _i(seq), _d(xyz)))
        self.n_obs[sid] += stamp.size

    # ---- hot path ----
    def optimize(self, options=None, log_cap=256):
        opt = options if options is not None else self.options_cls()
        summ = Summary()
        log = (Iteration * log_cap)()
        n = C.c_int(0)
        rc = self._f("optimize")(self.h, C.addressof(opt), C.addressof(summ), C.addressof(log), log_cap, C.byref(n))
        self.last_summary, self.last_log = summ, [log[i] for i in range(min(n.value, log_cap))]
        self._check(rc)
        return summ, self.last_log

    def evaluate_sensor(self, sid, want_jac=True, raw_flags=False):
        """raw_flags: return the uint8 flags of cb2_evaluate_sensor (bit 0 evaluated, bit 1 camera point behind the image plane) instead of a bool mask."""
        m = 2 if self.kind[sid] == CAMERA else 3
        n = self.n_obs[sid]
        W = 6 * self.k + self.n_intr[sid] + 7
        r = np.zeros((n, m))
        J = np.zeros((n, m, W)) if want_jac else None
        valid = np.zeros(n, dtype=np.uint8)
        self._check(self._f("evaluate_sensor")(self.h, sid, _d(r), _d(J), _u(valid)))
        return r, J, (valid if raw_flags else valid.astype(bool))

    def cost(self):
        c = C.c_double(0)
        ok = C.c_int(0)
        self._check(self._f("cost")(self.h, C.byref(c), C.byref(ok)))
        return c.value, bool(ok.value)

    def get_sensor(self, sid):
        intr = np.zeros(self.n_intr[sid])
        q, t = np.zeros(4), np.zeros(3)
        lat = C.c_double(0)
        self._check(self._f("get_sensor")(self.h, sid, _d(intr), _d(q), _d(t), C.byref(lat)))
        return intr, q, t, lat.value

    def get_rigid_body(self, id, n_pts):
        q, t, pts = np.zeros(4), np.zeros(3), np.zeros((n_pts, 3))
        self._check(self._f("get_rigid_body")(self.h, id, _d(q), _d(t), _d(pts)))
        return q, t, pts

    def get_trajectory(self):
        ctrl = np.zeros((self.n_cp, 6))
        self._check(self._f("get_trajectory")(self.h, _d(ctrl)))
        return ctrl

    def set_device(self, device: int):
        if self.lib.cb2_set_device(int(device)) != OK:
            raise CalicoError(INTERNAL, f"cannot select CUDA device {device}")

    def shard_plan(self, world: int, rank: int):
        """(n_chunks, chunk_lo, chunk_hi, seg_lo, seg_hi) of rank `rank` of `world` (host-side plan, no GPU needed)."""
        v = [C.c_int(0) for _ in range(5)]
        self._check(self.lib.cb2_shard_plan(self.h, world, rank, *[C.byref(x) for x in v]))
        return tuple(x.value for x in v)

    def comm_init(self, world: int, rank: int, unique_id: bytes):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._check(self.lib.cb2_comm_init(self.h, world, rank, buf))

    def comm_clone(self, other: "CApi"):
        self._check(self.lib.cb2_comm_clone(self.h, other.h))

    def comm_unique_id(self) -> bytes:
        buf = (C.c_uint8 * 128)()
        if self.lib.cb2_comm_unique_id(buf) != OK:
            raise CalicoError(INTERNAL, "ncclGetUniqueId failed (is libnccl.so.2 loadable?)")
        return bytes(buf)

    def comm_init_torch(self, world: int, rank: int):
        """NCCL communicator over the ranks of an initialised torch.distributed job: rank 0's unique id is broadcast."""
        import torch
        import torch.distributed as dist
        t = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(self.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(t, src=0)
        self.comm_init(world, rank, bytes(t.cpu().numpy().tobytes()))

    def comm_init_local(self, world: int, rank: int, group: str):
        """Emulation build only (tests): ranks are host threads of this process."""
        fn = self.lib.cb2_comm_init_local
        fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p]
        fn.restype = C.c_int
        self._check(fn(self.h, world, rank, group.encode()))

    def version(self) -> str:
        return self.lib.cb2_version().decode()

    def upload(self):
        self._check(self._f("upload")(self.h))

    def reset_parameters(self):
        self._check(self._f("reset_parameters")(self.h))

    def stats_reset(self):
        self._check(self._f("stats_reset")(self.h))

    def stats(self) -> "Stats":
        st = Stats()
        self._check(self._f("stats_get")(self.h, C.addressof(st)))
        return st

    def get_residuals(self, sid):
        m = 2 if self.kind[sid] == CAMERA else 3
        r = np.zeros((self.n_obs[sid], m))
        valid = np.zeros(self.n_obs[sid], dtype=np.uint8)
        self._check(self._f("get_residuals")(self.h, sid, _d(r), _u(valid)))
        return r, valid.astype(bool)


# ---- trajectory spline fit (stateless entry points) ----
def _fit_lib(lib_path):
    if not os.path.exists(lib_path):
        raise ImportError(f"{lib_path} not found: the CUDA extension has not been built (there is no CPU fallback).")
    lib = C.CDLL(lib_path)
    lib.cb2_fit_spline_size.argtypes = [C.c_int, _dp, C.c_int, C.c_double, _ip, _ip]
    lib.cb2_fit_spline_size.restype = C.c_int
    lib.cb2_fit_spline.argtypes = [C.c_int, _dp, _dp, C.c_int, C.c_double, C.c_int, _dp, C.c_int, _dp]
    lib.cb2_fit_spline.restype = C.c_int
    lib.cb2_fit_trajectory.argtypes = [C.c_int, _dp, _dp, _dp, C.c_int, C.c_double, C.c_int, _dp, C.c_int, _dp]
    lib.cb2_fit_trajectory.restype = C.c_int
    lib.cb2_fit_last_error.restype = C.c_char_p
    return lib


def _fit_check(lib, rc):
    if rc != OK:
        raise CalicoError(rc, (lib.cb2_fit_last_error() or b"").decode())


def fit_spline(times, data6, spline_order=6, knot_frequency=10.0, lib_path: str = LIB_PATH):
    """BSpline<6>::FitToData (bspline.hpp:20-38) on the device: returns (knots[n_knots], ctrl[n_cp, 6])."""
    lib = _fit_lib(lib_path)
    times, data6 = f64(times), f64(data6, (-1, 6))
    if data6.shape[0] != times.size:
        raise CalicoError(INVALID_ARGUMENT, "Data and time vectors are not the same size.")
    nk, ncp = C.c_int(0), C.c_int(0)
    _fit_check(lib, lib.cb2_fit_spline_size(times.size, _d(times), spline_order, knot_frequency, C.byref(nk), C.byref(ncp)))
    knots, ctrl = np.zeros(nk.value), np.zeros((ncp.value, 6))
    _fit_check(lib, lib.cb2_fit_spline(times.size, _d(times), _d(data6), spline_order, knot_frequency, nk.value, _d(knots), ncp.value, _d(ctrl)))
    return knots, ctrl


def fit_trajectory(stamps, q_xyzw, t_world_rig, knot_frequency=10.0, spline_order=6, lib_path: str = LIB_PATH):
    """Trajectory::FitSpline (trajectory.cpp:14-49) on the device: returns (knots, ctrl[n_cp, 6] = [axis-angle ; translation])."""
    lib = _fit_lib(lib_path)
    stamps, q, t = f64(stamps), f64(q_xyzw, (-1, 4)), f64(t_world_rig, (-1, 3))
    ts = np.sort(stamps)
    nk, ncp = C.c_int(0), C.c_int(0)
    _fit_check(lib, lib.cb2_fit_spline_size(ts.size, _d(ts), spline_order, knot_frequency, C.byref(nk), C.byref(ncp)))
    knots, ctrl = np.zeros(nk.value), np.zeros((ncp.value, 6))
    _fit_check(lib, lib.cb2_fit_trajectory(stamps.size, _d(stamps), _d(q), _d(t), spline_order, knot_frequency, nk.value, _d(knots), ncp.value, _d(ctrl)))
    return knots, ctrl


_num_intr_libs = {}


def num_intrinsics(kind: int, model: int, lib_path: str = LIB_PATH) -> int:
    """cb2_num_intrinsics: CameraModel::NumberOfParameters / the IMU models' parameter counts; -1 = unknown. Needs no GPU."""
    lib = _num_intr_libs.get(lib_path)
    if lib is None:
        if not os.path.exists(lib_path):
            raise ImportError(f"{lib_path} not found: the CUDA extension has not been built (there is no CPU fallback).")
        lib = C.CDLL(lib_path)
        lib.cb2_num_intrinsics.argtypes = [C.c_int, C.c_int]
        lib.cb2_num_intrinsics.restype = C.c_int
        _num_intr_libs[lib_path] = lib
    return int(lib.cb2_num_intrinsics(int(kind), int(model)))
