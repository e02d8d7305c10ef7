"""Host-side spline bookkeeping: knot vector, Qin basis matrices, least-squares fit, evaluation.

This is the step immediately before the hot path: the reference's `BSpline<N>::FitToData`
(calico/bspline.hpp:20-38,247-297) and `Trajectory::FitSpline` (calico/trajectory.cpp:14-49). It produces the knot
vector and control points handed to `cb2_set_trajectory`. The fit itself runs on the device (cb2_fit_spline, csrc/cb2_fit.cuh): the
same normal equations X'X c = X'd as the reference, solved by the banded Cholesky its own TODO asks for (bspline.hpp:287-289)
instead of a dense column-pivoted QR of an N_cp x N_cp matrix. This module keeps the bookkeeping (knots, basis, evaluation).
"""
from __future__ import annotations

import numpy as np


def compute_knot_vector(t_front: float, t_back: float, knot_frequency: float, spline_order: int):
    """BSpline::ComputeKnotVector, bspline.hpp:164-180. Returns (knots, valid_knots)."""
    deg = spline_order - 1
    duration = t_back - t_front
    dt = 1.0 / knot_frequency
    num_valid = 1 + int(np.ceil(duration * knot_frequency))
    num_knots = num_valid + 2 * deg
    idx = np.arange(-deg, num_knots - deg)
    knots = t_front + dt * idx
    return knots, knots[deg:deg + num_valid].copy()


def basis_matrix(knots: np.ndarray, k: int, i: int) -> np.ndarray:
    """BSpline::M(k, i) with d_0/d_1, bspline.hpp:192-244 (Qin's general matrix representation)."""
    if k == 1:
        return np.array([[float(k)]])
    Mkm1 = basis_matrix(knots, k - 1, i)
    n = k - 1
    M1 = np.zeros((k, n))
    M2 = np.zeros((k, n))
    M1[:n, :] = Mkm1
    M2[1:, :] = Mkm1
    A = np.zeros((n, k))
    B = np.zeros((n, k))
    for index in range(k - 1):
        j = i - k + 2 + index
        den = knots[j + k - 1] - knots[j]
        d0 = 0.0 if den <= 0.0 else (knots[i] - knots[j]) / den
        d1 = 0.0 if den <= 0.0 else (knots[i + 1] - knots[i]) / den
        A[index, index] = 1.0 - d0
        A[index, index + 1] = d0
        B[index, index] = -d1
        B[index, index + 1] = d1
    return M1 @ A + M2 @ B


def spline_index(valid_knots: np.ndarray, t: np.ndarray) -> np.ndarray:
    """BSpline::GetSplineIndex, bspline.hpp:139-151 (vectorised). -1 where t is past the last valid knot."""
    t = np.asarray(t, dtype=np.float64)
    idx = np.searchsorted(valid_knots, t, side="right") - 1
    idx = np.where(t == valid_knots[-1], len(valid_knots) - 2, idx)
    idx = np.where(t > valid_knots[-1], -1, idx)
    return idx.astype(np.int64)


def uniform_basis(k: int) -> np.ndarray:
    """Basis matrix of an interior segment of a uniform knot vector (same for every segment up to rounding)."""
    knots = np.arange(4 * k, dtype=np.float64)
    return basis_matrix(knots, k, 2 * k)


class Spline:
    """Data holder mirroring what BSpline<6> keeps after FitToData: knots, valid knots, per-segment M, control points."""

    def __init__(self, spline_order: int, knots: np.ndarray, ctrl: np.ndarray):
        self.k = int(spline_order)
        self.knots = np.asarray(knots, dtype=np.float64)
        self.ctrl = np.asarray(ctrl, dtype=np.float64).reshape(-1, 6)
        deg = self.k - 1
        if self.knots.size != self.ctrl.shape[0] + self.k:
            raise ValueError("knot vector size must equal number of control points + spline order")
        self.valid_knots = self.knots[deg:self.knots.size - deg]
        self._basis = None

    @property
    def basis(self):
        if self._basis is None:
            deg = self.k - 1
            nseg = self.valid_knots.size - 1
            self._basis = np.stack([basis_matrix(self.knots, self.k, i + deg) for i in range(nseg)])
        return self._basis

    def weights(self, t, derivative=0):
        """Rows of U*M for each time (BSpline::GetSplineBasis, bspline.hpp:103-136) and the segment index."""
        t = np.asarray(t, dtype=np.float64)
        seg = spline_index(self.valid_knots, t)
        if np.any(seg < 0) or np.any(t < self.valid_knots[0]):
            raise ValueError("Cannot interpolate. Value is not within valid knots.")
        deg = self.k - 1
        k0 = self.knots[seg + deg]
        k1 = self.knots[seg + deg + 1]
        dt_inv = 1.0 / (k1 - k0)
        u = (t - k0) * dt_inv
        U = np.zeros((t.size, self.k))
        for i in range(derivative, self.k):
            coeff = 1.0
            for j in range(i - derivative, i):
                coeff *= (j + 1)
            U[:, i] = coeff * u ** (i - derivative) * dt_inv ** derivative
        W = np.einsum("ni,nij->nj", U, self.basis[seg])
        return W, seg

    def evaluate(self, t, derivative=0):
        """BSpline::Interpolate, bspline.hpp:74-101."""
        if derivative < 0 or derivative > self.k - 1:
            raise ValueError("Invalid derivative for interpolation.")
        W, seg = self.weights(t, derivative)
        idx = seg[:, None] + np.arange(self.k)[None, :]
        return np.einsum("nj,njd->nd", W, self.ctrl[idx])


def fit_spline(times, data, spline_order=6, knot_frequency=10.0, lib_path=None) -> Spline:
    """BSpline::FitToData + FitSpline (bspline.hpp:20-38, 247-297) on the device through the C ABI (cb2_fit_spline): banded
    normal equations, banded Cholesky. There is no host fallback: without the CUDA library this raises."""
    from . import _capi
    data = np.asarray(data, dtype=np.float64)
    if data.ndim != 2 or data.shape[1] != 6:
        raise ValueError("fit_spline expects [n, 6] pose vectors [axis-angle ; translation]")
    knots, ctrl = _capi.fit_spline(times, data, spline_order, knot_frequency, **({"lib_path": lib_path} if lib_path else {}))
    return Spline(spline_order, knots, ctrl)


# ---- SO(3) helpers used by Trajectory::FitSpline (trajectory.cpp:14-49,81-93) ----
def quat_xyzw_to_angle_axis(q):
    """Eigen::AngleAxisd(Quaternion): angle = 2 atan2(|vec|, |w|) with axis sign-flipped when w < 0."""
    q = np.asarray(q, dtype=np.float64).reshape(-1, 4)
    vec, w = q[:, :3].copy(), q[:, 3].copy()
    n = np.linalg.norm(vec, axis=1)
    neg = w < 0
    angle = 2.0 * np.arctan2(n, np.abs(w))
    axis = np.where(n[:, None] > 0, vec / np.where(n > 0, n, 1.0)[:, None], np.array([1.0, 0, 0])[None, :])
    axis = np.where(neg[:, None], -axis, axis)
    return axis * angle[:, None]


def unwrap_phase_log_map(phi):
    """Trajectory::UnwrapPhaseLogMap, trajectory.cpp:81-93 (sequential by construction)."""
    phi = np.array(phi, dtype=np.float64)
    for i in range(1, phi.shape[0]):
        theta = np.linalg.norm(phi[i])
        if theta == 0:
            continue
        kk = np.round((phi[i] @ phi[i - 1] - theta * theta) / (2.0 * np.pi * theta))
        phi[i] = phi[i] * (1.0 + 2.0 * np.pi * kk / theta)
    return phi


def fit_trajectory(stamps, q_xyzw, t_world_rig, knot_frequency=10.0, spline_order=6, lib_path=None) -> Spline:
    """Trajectory::FitSpline, trajectory.cpp:14-49, through the C ABI (cb2_fit_trajectory): poses -> [axis-angle ; translation]
    6-vectors (sorted, phase-unwrapped on the host side of the library) -> device spline fit."""
    from . import _capi
    knots, ctrl = _capi.fit_trajectory(stamps, q_xyzw, t_world_rig, knot_frequency, spline_order, **({"lib_path": lib_path} if lib_path else {}))
    return Spline(spline_order, knots, ctrl)


def angle_axis_to_quat_xyzw(aa):
    aa = np.asarray(aa, dtype=np.float64).reshape(-1, 3)
    th = np.linalg.norm(aa, axis=1)
    small = th == 0
    k = np.where(small, 0.5, np.sin(0.5 * th) / np.where(small, 1.0, th))
    return np.concatenate([aa * k[:, None], np.where(small, 1.0, np.cos(0.5 * th))[:, None]], axis=1)


def quat_mul_xyzw(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx,
                     aw * bw - ax * bx - ay * by - az * bz], axis=-1)
