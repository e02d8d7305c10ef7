"""TEST INFRASTRUCTURE ONLY — CPU restatement (numpy) of the reference's trajectory spline fit, the oracle for cb2_fit_spline /
cb2_fit_trajectory. Imported by tests/ only; nothing under calico_b200/ may import it.

Restated from the reference (paths relative to /root/reference):
  BSpline<N,T>::FitToData               calico/bspline.hpp:20-38
  BSpline<N,T>::ComputeKnotVector       calico/bspline.hpp:164-180
  BSpline<N,T>::M / d_0 / d_1           calico/bspline.hpp:192-244   (Qin's recursive basis matrices)
  BSpline<N,T>::FitSpline               calico/bspline.hpp:247-297   (dense X, X'X, column-pivoted Householder QR solve)
  BSpline<N,T>::Interpolate / basis     calico/bspline.hpp:74-136
  Trajectory::FitSpline                 calico/trajectory.cpp:14-49
  Trajectory::UnwrapPhaseLogMap         calico/trajectory.cpp:81-93
Pinned by tests/test_spline_fit.py against the reference's own acceptance test (calico/test/bspline_test.cpp:52-94).
"""
import numpy as np
import scipy.linalg


def knot_vector(t_front, t_back, knot_frequency, order):
    """bspline.hpp:164-180."""
    deg = order - 1
    dt = 1.0 / knot_frequency
    num_valid = 1 + int(np.ceil((t_back - t_front) * knot_frequency))
    num_knots = num_valid + 2 * deg
    knots = np.array([t_front + dt * i for i in range(-deg, num_knots - deg)])
    return knots, knots[deg:deg + num_valid].copy()


def _d0(knots, k, i, j):
    den = knots[j + k - 1] - knots[j]
    return 0.0 if den <= 0.0 else (knots[i] - knots[j]) / den


def _d1(knots, k, i, j):
    den = knots[j + k - 1] - knots[j]
    return 0.0 if den <= 0.0 else (knots[i + 1] - knots[i]) / den


def basis_matrix(knots, k, i):
    """bspline.hpp:192-226: M_k(i) = [M_{k-1}; 0] A + [0; M_{k-1}] B."""
    if k == 1:
        return np.array([[float(k)]])
    Mkm1 = basis_matrix(knots, k - 1, i)
    n = k - 1
    M1 = np.zeros((k, n)); M1[:n] = Mkm1
    M2 = np.zeros((k, n)); M2[1:] = Mkm1
    A = np.zeros((n, k)); B = np.zeros((n, k))
    for index in range(n):
        j = i - k + 2 + index
        d0, d1 = _d0(knots, k, i, j), _d1(knots, k, i, j)
        A[index, index], A[index, index + 1] = 1.0 - d0, d0
        B[index, index], B[index, index + 1] = -d1, d1
    return M1 @ A + M2 @ B


def fit_spline(times, data, order, knot_frequency):
    """bspline.hpp:20-38 + 247-297. Returns (knots, valid_knots, basis[n_seg], ctrl[n_cp, N])."""
    times = np.asarray(times, float); data = np.asarray(data, float)
    if times.size == 0: raise ValueError("Attempted to fit data on empty time vector.")
    if data.shape[0] != times.size: raise ValueError("Data and time vectors are not the same size.")
    if np.any(np.diff(times) < 0): raise ValueError("Time vector is not monotonically increasing.")
    if order < 2: raise ValueError(f"Spline order must be greater than 2. Got {order}")
    if knot_frequency <= 0: raise ValueError("Knot frequency must be greater than 0.")
    deg = order - 1
    knots, valid = knot_vector(times[0], times[-1], knot_frequency, order)
    basis = [basis_matrix(knots, order, i + deg) for i in range(valid.size - 1)]
    n_cp = knots.size - order
    X = np.zeros((times.size, n_cp))
    for j, t in enumerate(times):
        if t == valid[-1]: seg = len(basis) - 1
        elif t == valid[0]: seg = 0
        else: seg = int(np.searchsorted(valid, t, side="right")) - 1     # std::upper_bound - 1
        ti, tii = knots[seg + deg], knots[seg + deg + 1]
        u = (t - ti) / (tii - ti)
        U = np.ones(order)
        for i in range(1, order): U[i] = u * U[i - 1]
        X[j, seg:seg + order] = U @ basis[seg]
    XtX, Xtd = X.T @ X, X.T @ data
    # Eigen colPivHouseholderQr().solve: QR with column pivoting of X'X, basic solution.
    Q, R, piv = scipy.linalg.qr(XtX, pivoting=True)
    rank = int(np.sum(np.abs(np.diag(R)) > np.abs(R[0, 0]) * max(XtX.shape) * np.finfo(float).eps))
    y = Q.T @ Xtd
    sol = np.zeros_like(Xtd)
    sol[piv[:rank]] = scipy.linalg.solve_triangular(R[:rank, :rank], y[:rank])
    return knots, valid, basis, sol


def interpolate(knots, valid, basis, ctrl, t, derivative, order):
    """bspline.hpp:74-136."""
    if derivative < 0 or derivative > order - 1: raise ValueError("Invalid derivative for interpolation.")
    deg = order - 1
    out = []
    for tt in np.atleast_1d(t):
        if tt < valid[0] or tt > valid[-1]: raise ValueError("Cannot interpolate. Value is not within valid knots.")
        seg = len(basis) - 1 if tt == valid[-1] else int(np.searchsorted(valid, tt, side="right")) - 1
        ti, tii = knots[seg + deg], knots[seg + deg + 1]
        dt_inv = 1.0 / (tii - ti)
        u = (tt - ti) * dt_inv
        U = np.zeros(order)
        for i in range(derivative, order):
            c = 1.0
            for j in range(i - derivative, i): c *= (j + 1)
            U[i] = c * u ** (i - derivative) * dt_inv ** derivative
        out.append(U @ basis[seg] @ ctrl[seg:seg + order])
    return np.array(out)


def quat_to_angle_axis(q_xyzw):
    """Eigen::AngleAxisd(Quaterniond) as used at trajectory.cpp:33-34."""
    out = []
    for x, y, z, w in np.asarray(q_xyzw, float).reshape(-1, 4):
        n = np.sqrt(x * x + y * y + z * z)
        if n == 0.0:
            out.append([0.0, 0.0, 0.0]); continue
        angle = 2.0 * np.arctan2(n, abs(w))
        s = -1.0 if w < 0 else 1.0
        out.append([s * x / n * angle, s * y / n * angle, s * z / n * angle])
    return np.array(out)


def unwrap_phase_log_map(phi):
    """trajectory.cpp:81-93."""
    phi = np.array(phi, float)
    for i in range(1, len(phi)):
        theta = np.linalg.norm(phi[i])
        if theta == 0: continue
        k = np.round((phi[i] @ phi[i - 1] - theta * theta) / (2.0 * np.pi * theta))
        phi[i] *= 1.0 + 2.0 * np.pi * k / theta
    return phi


def fit_trajectory(stamps, q_xyzw, t_world_rig, knot_frequency, order=6):
    """trajectory.cpp:14-49."""
    o = np.argsort(np.asarray(stamps, float), kind="stable")
    ts = np.asarray(stamps, float)[o]
    phi = unwrap_phase_log_map(quat_to_angle_axis(np.asarray(q_xyzw, float).reshape(-1, 4)[o]))
    data = np.concatenate([phi, np.asarray(t_world_rig, float).reshape(-1, 3)[o]], axis=1)
    return fit_spline(ts, data, order, knot_frequency)
