"""ORACLE — TEST INFRASTRUCTURE ONLY. ctypes access to oracle/liboracle.so (the CPU restatement of the reference).

Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from calico_b200 import _capi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "liboracle.so")


def _cpu_stamp() -> str:
    """The host's instruction-set flags: the library is built with -march=native, so a binary built on another CPU model must be rebuilt."""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    import hashlib
                    return hashlib.sha1(line.encode()).hexdigest()
    except OSError:
        pass
    return "unknown"


def build(force: bool = False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".hpp", ".cpp")) or f == "Makefile"]
    stamp_path = os.path.join(_HERE, ".build_cpu")
    stamp = _cpu_stamp()
    try:
        with open(stamp_path) as f:
            same_cpu = f.read().strip() == stamp
    except OSError:
        same_cpu = False
    if force or not same_cpu or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
        with open(stamp_path, "w") as f:
            f.write(stamp)
    return LIB


class OracleOptions(C.Structure):
    """orc::Options (oracle/calico_problem.hpp)."""
    _fields_ = _capi.Options._fields_[:-1]

    def __init__(self, **kw):
        super().__init__()
        d = _capi.Options()
        for name, _ in self._fields_:
            setattr(self, name, getattr(d, name))
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError(k)
            setattr(self, k, v)


def oracle_api() -> _capi.CApi:
    build()
    return _capi.CApi(LIB, "orc_", OracleOptions)


def _oracle_fit_trajectory(stamps, q_xyzw, t_world_rig, knot_frequency=10.0, spline_order=6):
    """Trajectory::FitSpline restated on the CPU (oracle/spline_fit.py) for workloads generated without a GPU (tests only)."""
    from calico_b200 import spline as sp
    from oracle import spline_fit
    knots, _, _, ctrl = spline_fit.fit_trajectory(stamps, q_xyzw, t_world_rig, knot_frequency, spline_order)
    return sp.Spline(spline_order, knots, ctrl)


oracle_api.fit_trajectory = _oracle_fit_trajectory


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        dp = C.POINTER(C.c_double)
        _lib.orc_radius_after_accept.restype = C.c_double
        _lib.orc_radius_after_accept.argtypes = [C.c_double, C.c_double]
        _lib.orc_project_point.argtypes = [C.c_int, dp, dp, dp]
        _lib.orc_imu_project.argtypes = [C.c_int, dp, dp, dp]
        _lib.orc_basis_matrix.argtypes = [C.c_int, C.c_int, dp, C.c_int, dp]
        for f in ("orc_exp_so3", "orc_exp_so3_jacobian"):
            getattr(_lib, f).argtypes = [dp, dp]
            getattr(_lib, f).restype = None
        _lib.orc_exp_so3_jacobian_dot.argtypes = [dp, dp, dp]
        _lib.orc_exp_so3_jacobian_dot.restype = None
        _lib.orc_quaternion_plus.argtypes = [dp, dp, dp]
        _lib.orc_quaternion_plus.restype = None
        _lib.orc_angle_axis_to_quaternion.argtypes = [dp, dp]
        _lib.orc_angle_axis_to_quaternion.restype = None
        _lib.orc_loss.argtypes = [C.c_int, C.c_double, C.c_double, dp]
        _lib.orc_loss.restype = None
        _lib.orc_spline_interpolate.argtypes = [C.c_void_p, C.c_int, dp, C.c_int, dp]
        _lib.orc_project_camera.argtypes = [C.c_void_p, C.c_int, C.c_int, dp, C.c_int, dp, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                            C.POINTER(C.c_int), dp, C.POINTER(C.c_int)]
        _lib.orc_project_imu.argtypes = [C.c_void_p, C.c_int, C.c_int, dp, dp, dp]
    return _lib


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def project_point(model, intr, p):
    intr, p, out = _capi.f64(intr), _capi.f64(p), np.zeros(2)
    ok = lib().orc_project_point(model, _d(intr), _d(p), _d(out))
    return bool(ok), out


def imu_project(model, intr, w):
    intr, w, out = _capi.f64(intr), _capi.f64(w), np.zeros(3)
    ok = lib().orc_imu_project(model, _d(intr), _d(w), _d(out))
    return bool(ok), out


def basis_matrix(k, knots, i):
    knots, out = _capi.f64(knots), np.zeros((k, k))
    lib().orc_basis_matrix(k, knots.size, _d(knots), i, _d(out))
    return out


def exp_so3(phi):
    phi, out = _capi.f64(phi), np.zeros((3, 3))
    lib().orc_exp_so3(_d(phi), _d(out))
    return out


def exp_so3_jacobian(phi):
    phi, out = _capi.f64(phi), np.zeros((3, 3))
    lib().orc_exp_so3_jacobian(_d(phi), _d(out))
    return out


def exp_so3_jacobian_dot(phi, phi_dot):
    phi, phi_dot, out = _capi.f64(phi), _capi.f64(phi_dot), np.zeros((3, 3))
    lib().orc_exp_so3_jacobian_dot(_d(phi), _d(phi_dot), _d(out))
    return out


def quaternion_plus(q, d):
    q, d, out = _capi.f64(q), _capi.f64(d), np.zeros(4)
    lib().orc_quaternion_plus(_d(q), _d(d), _d(out))
    return out


def angle_axis_to_quaternion(aa):
    aa, out = _capi.f64(aa), np.zeros(4)
    lib().orc_angle_axis_to_quaternion(_d(aa), _d(out))
    return out


def loss(kind, a, s):
    out = np.zeros(3)
    lib().orc_loss(kind, a, s, _d(out))
    return out


def radius_after_accept(radius, ratio):
    return lib().orc_radius_after_accept(radius, ratio)


def spline_interpolate(api, times, derivative=0):
    times = _capi.f64(times)
    out = np.zeros((times.size, 6))
    rc = lib().orc_spline_interpolate(api.h, times.size, _d(times), derivative, _d(out))
    api._check(rc)
    return out


def project_camera(api, sid, times, cap=None):
    """Camera::Project restatement (camera.cpp:155-208) at the handle's current state."""
    times = _capi.f64(times)
    cap = cap or times.size * 1024
    stamp, pix = np.zeros(cap), np.zeros((cap, 2))
    img, mid, fid = (np.zeros(cap, dtype=np.int32) for _ in range(3))
    n = C.c_int(0)
    ip = C.POINTER(C.c_int)
    rc = lib().orc_project_camera(api.h, sid, times.size, _d(times), cap, _d(stamp), img.ctypes.data_as(ip), mid.ctypes.data_as(ip),
                                  fid.ctypes.data_as(ip), _d(pix), C.byref(n))
    api._check(rc)
    n = n.value
    return stamp[:n], img[:n], mid[:n], fid[:n], pix[:n]


def project_imu(api, sid, times):
    times = _capi.f64(times)
    stamp, xyz = np.zeros(times.size), np.zeros((times.size, 3))
    api._check(lib().orc_project_imu(api.h, sid, times.size, _d(times), _d(stamp), _d(xyz)))
    return stamp, xyz
