// ORACLE — TEST INFRASTRUCTURE ONLY (see jet.hpp header). CPU restatement of the reference's
// geometry, spline, sensor models and cost functors, templated on T (double or orc::Jet<N>)
// exactly as the reference templates them on ceres::Jet. Each function cites the reference
// file:line it follows (paths relative to /root/reference/).
#pragma once
#include <cmath>
#include <vector>

#include "jet.hpp"

namespace orc {

// ---------------------------------------------------------------------------------------
// Tiny fixed-size algebra standing in for the Eigen types the reference uses.
// ---------------------------------------------------------------------------------------
template <typename T> struct V3 { T x, y, z; };
template <typename T> struct M3 { T m[3][3]; };   // row-major m[r][c]
// Eigen::Quaternion storage order is coeffs() = x,y,z,w (typedefs.h:69-81).
template <typename T> struct Qt { T x, y, z, w; };

template <typename T> inline V3<T> operator+(const V3<T>& a, const V3<T>& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <typename T> inline V3<T> operator-(const V3<T>& a, const V3<T>& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <typename T> inline V3<T> operator-(const V3<T>& a) { return {-a.x, -a.y, -a.z}; }
template <typename T> inline V3<T> operator*(const T& s, const V3<T>& a) { return {s * a.x, s * a.y, s * a.z}; }
template <typename T> inline T dot(const V3<T>& a, const V3<T>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename T> inline V3<T> cross(const V3<T>& a, const V3<T>& b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
template <typename T> inline M3<T> m3_zero() { M3<T> r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = T(0.0); return r; }
template <typename T> inline M3<T> m3_identity() { M3<T> r = m3_zero<T>(); r.m[0][0] = r.m[1][1] = r.m[2][2] = T(1.0); return r; }
template <typename T> inline M3<T> operator+(const M3<T>& a, const M3<T>& b) { M3<T> r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][j] + b.m[i][j]; return r; }
template <typename T> inline M3<T> operator*(const T& s, const M3<T>& a) { M3<T> r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = s * a.m[i][j]; return r; }
template <typename T> inline M3<T> operator-(const M3<T>& a) { M3<T> r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = -a.m[i][j]; return r; }
template <typename T> inline M3<T> operator*(const M3<T>& a, const M3<T>& b) {
  M3<T> r;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
  return r;
}
template <typename T> inline V3<T> operator*(const M3<T>& a, const V3<T>& v) {
  return {a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z,
          a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z,
          a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z};
}

// Eigen quaternion product (Hamilton), Eigen/src/Geometry/Quaternion.h semantics.
template <typename T> inline Qt<T> operator*(const Qt<T>& a, const Qt<T>& b) {
  Qt<T> r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}
// Eigen QuaternionBase::inverse(): conjugate / squaredNorm (zero quaternion if norm is 0).
template <typename T> inline Qt<T> q_inverse(const Qt<T>& q) {
  const T n2 = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
  if (n2 > T(0.0)) return {-q.x / n2, -q.y / n2, -q.z / n2, q.w / n2};
  return {T(0.0), T(0.0), T(0.0), T(0.0)};
}
// Eigen QuaternionBase::_transformVector(): v + w*uv + vec x uv with uv = 2 (vec x v).
template <typename T> inline V3<T> q_rotate(const Qt<T>& q, const V3<T>& v) {
  const V3<T> qv{q.x, q.y, q.z};
  V3<T> uv = cross(qv, v);
  uv = uv + uv;
  return v + q.w * uv + cross(qv, uv);
}

// ceres::AngleAxisToQuaternion (ceres/rotation.h; Ceres is external — restated from its
// published source). Output order w,x,y,z; call sites camera_cost_functor.h:122,
// accelerometer_cost_functor.h:115, trajectory.h:98 feed it to Eigen::Quaternion(w,x,y,z).
template <typename T> inline Qt<T> AngleAxisToQuaternion(const V3<T>& aa) {
  const T theta_squared = aa.x * aa.x + aa.y * aa.y + aa.z * aa.z;
  Qt<T> q;
  if (scalar_part(theta_squared) > 0.0) {
    const T theta = sqrt(theta_squared);
    const T half_theta = theta * T(0.5);
    const T k = sin(half_theta) / theta;
    q.w = cos(half_theta); q.x = aa.x * k; q.y = aa.y * k; q.z = aa.z * k;
  } else {
    const T k(0.5);
    q.w = T(1.0); q.x = aa.x * k; q.y = aa.y * k; q.z = aa.z * k;
  }
  return q;
}

// ---------------------------------------------------------------------------------------
// geometry.h
// ---------------------------------------------------------------------------------------
// geometry.h:12-23
template <typename T> inline M3<T> Skew(const V3<T>& v) {
  M3<T> V = m3_zero<T>();
  V.m[0][1] = -v.z; V.m[1][0] = v.z; V.m[0][2] = v.y; V.m[2][0] = -v.y; V.m[1][2] = -v.x; V.m[2][1] = v.x;
  return V;
}
// geometry.h:36-42
template <typename T> inline T SmallAngleSin(const T theta) {
  const T theta_sq = theta * theta;
  return theta * (T(1.0) - theta_sq * (T(1.0 / 6.0) + theta_sq * (T(1.0 / 120.0) - theta_sq * T(1.0 / 5040.0))));
}
// geometry.h:45-51
template <typename T> inline T SmallAngleCos(const T theta) {
  const T theta_sq = theta * theta;
  return T(1.0) - theta_sq * (T(0.5) - theta_sq * (T(1.0 / 24.0) + theta_sq * (T(1.0 / 720.0) - theta_sq * T(1.0 / 40320.0))));
}
// geometry.h:54-75
template <typename T> inline M3<T> ExpSO3(const V3<T>& phi) {
  const T theta = sqrt(dot(phi, phi));
  if (theta == T(0.0)) return m3_identity<T>();
  T sin_theta, one_m_cos_theta;
  if (theta < T(1e-7)) { sin_theta = SmallAngleSin(theta); one_m_cos_theta = T(1.0) - SmallAngleCos(theta); }
  else { sin_theta = sin(theta); one_m_cos_theta = T(1.0) - cos(theta); }
  const V3<T> phi_hat = (T(1.0) / theta) * phi;
  const M3<T> Phi = Skew(phi_hat);
  return m3_identity<T>() + sin_theta * Phi + one_m_cos_theta * (Phi * Phi);
}
// geometry.h:138-161
template <typename T> inline M3<T> ExpSO3Jacobian(const V3<T>& phi) {
  const T theta_sq = dot(phi, phi);
  M3<T> J = m3_identity<T>();
  if (theta_sq == T(0.0)) return J;
  const T theta = sqrt(theta_sq);
  T one_m_cos_theta, sin_theta;
  if (theta < T(1e-7)) { sin_theta = SmallAngleSin(theta); one_m_cos_theta = T(1.0) - SmallAngleCos(theta); }
  else { sin_theta = sin(theta); one_m_cos_theta = T(1.0) - cos(theta); }
  const T inv_theta = T(1.0) / theta;
  const V3<T> phi_hat = inv_theta * phi;
  const M3<T> phi_hat_x = Skew(phi_hat);
  J = J + inv_theta * (one_m_cos_theta * phi_hat_x + (theta - sin_theta) * (phi_hat_x * phi_hat_x));
  return J;
}
// geometry.h:173-210 — reproduced verbatim including the non-standard c0/c2 (SURVEY §8 trap 2).
template <typename T> inline void ExpSO3Hessian(const V3<T>& phi, M3<T> H[3]) {
  const M3<T> G[3] = {Skew(V3<T>{T(1.0), T(0.0), T(0.0)}), Skew(V3<T>{T(0.0), T(1.0), T(0.0)}), Skew(V3<T>{T(0.0), T(0.0), T(1.0)})};
  for (int i = 0; i < 3; ++i) H[i] = m3_zero<T>();
  const T theta_sq = dot(phi, phi);
  if (theta_sq == T(0.0)) return;
  const T theta = sqrt(theta_sq);
  T ct, st;
  if (theta < T(1e-7)) { ct = SmallAngleCos(theta); st = SmallAngleSin(theta); }
  else { ct = cos(theta); st = sin(theta); }
  const T inv_theta = T(1.0) / theta;
  const T inv_theta_sq = inv_theta * inv_theta;
  const V3<T> phi_hat = inv_theta * phi;
  const M3<T> phi_hat_x = Skew(phi_hat);
  const T c0 = ct - st * inv_theta;
  const T c1 = (T(1.0) - ct) * inv_theta_sq;
  const T c2 = T(3.0) * inv_theta_sq * st - inv_theta * (ct - T(2.0));
  const T c3 = inv_theta_sq * (theta - st);
  const T ph[3] = {phi_hat.x, phi_hat.y, phi_hat.z};
  for (int i = 0; i < 3; ++i) {
    H[i] = (c0 * ph[i]) * phi_hat_x + c1 * G[i] + (c2 * ph[i]) * (phi_hat_x * phi_hat_x) +
           c3 * (G[i] * phi_hat_x + phi_hat_x * G[i]);
  }
}
// geometry.h:214-222 — column i of Jdot is H[i] * phi_dot.
template <typename T> inline M3<T> ExpSO3JacobianDot(const V3<T>& phi, const V3<T>& phi_dot) {
  M3<T> H[3];
  ExpSO3Hessian(phi, H);
  M3<T> Jdot;
  for (int i = 0; i < 3; ++i) {
    const V3<T> c = H[i] * phi_dot;
    Jdot.m[0][i] = c.x; Jdot.m[1][i] = c.y; Jdot.m[2][i] = c.z;
  }
  return Jdot;
}

// ---------------------------------------------------------------------------------------
// bspline.hpp
// ---------------------------------------------------------------------------------------
// BSpline<6,T>::Evaluate, bspline.hpp:40-72. control points are k x 6 (row i = control point i),
// basis is k x k row-major; returns (U * M * C)^T as 6 values.
template <typename T>
inline void SplineEvaluate(const T* ctrl /*k*6*/, int k, const T& knot0, const T& knot1,
                           const double* basis /*k*k*/, const T& stamp, int derivative, T out[6]) {
  const T dt = knot1 - knot0;
  const T dt_inv = T(1.0) / dt;
  const T u = (stamp - knot0) * dt_inv;
  T dnu_dtn = T(1.0);
  for (int j = 0; j < derivative; ++j) dnu_dtn *= dt_inv;
  std::vector<T> dc(k, T(1.0)), U(k, T(1.0));
  for (int i = 0; i < derivative && i < k; ++i) dc[i] = T(0.0);
  for (int i = derivative; i < k; ++i) {
    T coeff = T(1.0);
    for (int j = i - derivative; j < i; ++j) coeff *= T(double(j + 1));
    dc[i] = coeff;
    U[i] = (i > derivative) ? (u * U[i - 1]) : U[i];
  }
  for (int i = 0; i < k; ++i) U[i] = U[i] * dc[i] * dnu_dtn;
  std::vector<T> UM(k, T(0.0));
  for (int c = 0; c < k; ++c) { T s = T(0.0); for (int r = 0; r < k; ++r) s += U[r] * T(basis[r * k + c]); UM[c] = s; }
  for (int d = 0; d < 6; ++d) { T s = T(0.0); for (int c = 0; c < k; ++c) s += UM[c] * ctrl[c * 6 + d]; out[d] = s; }
}

// BSpline::M(k,i) with d_0/d_1, bspline.hpp:192-244 (Qin's general matrix recursion), on the
// spline's own knot vector. Returns the k x k row-major basis for knot interval i.
inline std::vector<double> BasisMatrix(const std::vector<double>& knots, int k, int i) {
  if (k == 1) return {double(k)};
  const std::vector<double> Mkm1 = BasisMatrix(knots, k - 1, i);
  const int n = k - 1;  // Mkm1 is n x n
  std::vector<double> M1(k * n, 0.0), M2(k * n, 0.0);   // (n+1) x n
  for (int r = 0; r < n; ++r) for (int c = 0; c < n; ++c) { M1[r * n + c] = Mkm1[r * n + c]; M2[(r + 1) * n + c] = Mkm1[r * n + c]; }
  std::vector<double> A(n * k, 0.0), B(n * k, 0.0);      // (k-1) x k
  for (int index = 0; index < k - 1; ++index) {
    const int j = i - k + 2 + index;
    const double den = knots[j + k - 1] - knots[j];
    const double d0 = den <= 0.0 ? 0.0 : (knots[i] - knots[j]) / den;
    const double d1 = den <= 0.0 ? 0.0 : (knots[i + 1] - knots[i]) / den;
    A[index * k + index] = 1.0 - d0; A[index * k + index + 1] = d0;
    B[index * k + index] = -d1;      B[index * k + index + 1] = d1;
  }
  std::vector<double> Mk(k * k, 0.0);
  for (int r = 0; r < k; ++r) for (int c = 0; c < k; ++c) {
    double s = 0.0;
    for (int t = 0; t < n; ++t) s += M1[r * n + t] * A[t * k + c] + M2[r * n + t] * B[t * k + c];
    Mk[r * k + c] = s;
  }
  return Mk;
}

// ---------------------------------------------------------------------------------------
// sensors/camera_models.h — ProjectPoint bodies. Return false where the reference returns a
// non-OK status. Enum values are ABI (camera_models.h:16-33).
// ---------------------------------------------------------------------------------------
enum CameraModelType { kCamNone = 0, kOpenCv5 = 1, kOpenCv8 = 2, kKannalaBrandt = 3, kDoubleSphere = 4,
                       kFieldOfView = 5, kUnifiedCamera = 6, kExtendedUnifiedCamera = 7 };
inline int CameraModelNumParams(int model) {
  switch (model) { case kOpenCv5: return 8; case kOpenCv8: return 11; case kKannalaBrandt: return 7; case kDoubleSphere: return 5;
                   case kFieldOfView: return 4; case kUnifiedCamera: return 4; case kExtendedUnifiedCamera: return 5; default: return -1; }
}

template <typename T> inline bool ProjectPoint(int model, const T* in, const V3<T>& p, T out[2]) {
  switch (model) {
    case kOpenCv5: {  // camera_models.h:105-141
      if (p.z <= T(0.0)) return false;
      const T &f = in[0], &cx = in[1], &cy = in[2], &k1 = in[3], &k2 = in[4], &p1 = in[5], &p2 = in[6], &k3 = in[7];
      T px = p.x / p.z, py = p.y / p.z;
      const T x = px, y = py;
      const T r2 = px * px + py * py;
      const T s = T(1.0) + r2 * (k1 + r2 * (k2 + r2 * k3));
      px *= s; py *= s;
      px += T(2.0) * p1 * x * y + p2 * (r2 + T(2.0) * x * x);
      py += T(2.0) * p2 * x * y + p1 * (r2 + T(2.0) * y * y);
      out[0] = px * f + cx; out[1] = py * f + cy;
      return true;
    }
    case kOpenCv8: {  // camera_models.h:257-298
      if (p.z <= T(0.0)) return false;
      const T &f = in[0], &cx = in[1], &cy = in[2], &k1 = in[3], &k2 = in[4], &p1 = in[5], &p2 = in[6], &k3 = in[7],
              &k4 = in[8], &k5 = in[9], &k6 = in[10];
      T px = p.x / p.z, py = p.y / p.z;
      const T x = px, y = py;
      const T r2 = px * px + py * py;
      const T s_num = T(1.0) + r2 * (k1 + r2 * (k2 + r2 * k3));
      const T s_den = T(1.0) + r2 * (k4 + r2 * (k5 + r2 * k6));
      const T s = s_num / s_den;
      px *= s; py *= s;
      px += T(2.0) * p1 * x * y + p2 * (r2 + T(2.0) * x * x);
      py += T(2.0) * p2 * x * y + p1 * (r2 + T(2.0) * y * y);
      out[0] = px * f + cx; out[1] = py * f + cy;
      return true;
    }
    case kKannalaBrandt: {  // camera_models.h:420-462 (Taylor branch :444-446)
      if (p.z <= T(0.0)) return false;
      const T &f = in[0], &cx = in[1], &cy = in[2], &k1 = in[3], &k2 = in[4], &k3 = in[5], &k4 = in[6];
      T px = p.x / p.z, py = p.y / p.z;
      const T r = sqrt(px * px + py * py);
      T s;
      if (r < T(1e-9)) {
        const T r2 = r * r;
        s = T(1.0) + r2 * (k1 - T(1.0 / 3.0) + r2 * (-k1 + k2 + 0.2));
      } else {
        const T theta = atan(r);
        const T theta2 = theta * theta;
        const T theta_d = theta * (T(1.0) + theta2 * (k1 + theta2 * (k2 + theta2 * (k3 + theta2 * k4))));
        s = theta_d / r;
      }
      out[0] = px * s * f + cx; out[1] = py * s * f + cy;
      return true;
    }
    case kDoubleSphere: {  // camera_models.h:623-657
      const T &xi = in[3], &alpha = in[4];
      const T w1 = alpha > T(0.5) ? (T(1.0) - alpha) / alpha : alpha / (T(1.0) - alpha);
      const T num = w1 + xi;
      const T w2_sq = num * num / (T(2.0) * w1 * xi + xi * xi + T(1.0));
      const T r2 = dot(p, p);
      if (p.z * p.z <= -w2_sq * r2) return false;
      const T &f = in[0], &cx = in[1], &cy = in[2];
      const T r = sqrt(r2);
      const T d = sqrt(r2 * (T(1.0) + xi * xi) + T(2.0) * xi * r * p.z);
      const T s = T(1.0) / (alpha * d + (T(1.0) - alpha) * (xi * r + p.z));
      out[0] = p.x * s * f + cx; out[1] = p.y * s * f + cy;
      return true;
    }
    case kFieldOfView: {  // camera_models.h:740-781 (branches :762-772)
      const T &f = in[0], &cx = in[1], &cy = in[2], &w = in[3];
      if (p.z <= T(0.0)) return false;
      T px = p.x / p.z, py = p.y / p.z;
      const T r = sqrt(px * px + py * py);
      T s;
      if (w * w < 1e-5) {
        s = T(1.0);
      } else {
        const T tan_term = T(2.0) * tan(w * T(0.5));
        if (r * r < 1e-5) s = tan_term / w;
        else s = atan(r * tan_term) / (r * w);
      }
      out[0] = px * s * f + cx; out[1] = py * s * f + cy;
      return true;
    }
    case kUnifiedCamera: {  // camera_models.h:872-901
      const T& alpha = in[3];
      const T w = alpha > T(0.5) ? (T(1.0) - alpha) / alpha : alpha / (T(1.0) - alpha);
      const T d = sqrt(dot(p, p));
      if (p.z <= -w * d) return false;
      const T &f = in[0], &cx = in[1], &cy = in[2];
      const T s = T(1.0) / (alpha * d + (T(1.0) - alpha) * p.z);
      out[0] = p.x * s * f + cx; out[1] = p.y * s * f + cy;
      return true;
    }
    case kExtendedUnifiedCamera: {  // camera_models.h:985-1015; note beta * norm (NOT squared), :995
      const T &alpha = in[3], &beta = in[4];
      const T d = sqrt(beta * sqrt(p.x * p.x + p.y * p.y) + p.z * p.z);
      const T w = alpha > T(0.5) ? (T(1.0) - alpha) / alpha : alpha / (T(1.0) - alpha);
      if (p.z <= -w * d) return false;
      const T &f = in[0], &cx = in[1], &cy = in[2];
      const T s = T(1.0) / (alpha * d + (T(1.0) - alpha) * p.z);
      out[0] = p.x * s * f + cx; out[1] = p.y * s * f + cy;
      return true;
    }
    default: return false;  // camera_models.h:1101-1103 "not supported"
  }
}

// sensors/accelerometer_models.h:80-85,129-141,208-235 and gyroscope_models.h:82,130,208
// (textual mirrors of each other). Enum values: kNone=0, ScaleOnly=1, ScaleAndBias=2, VectorNav=3.
enum ImuModelType { kImuNone = 0, kScaleOnly = 1, kScaleAndBias = 2, kVectorNav = 3 };
inline int ImuModelNumParams(int model) {
  switch (model) { case kScaleOnly: return 1; case kScaleAndBias: return 4; case kVectorNav: return 12; default: return -1; }
}
template <typename T> inline bool ImuProject(int model, const T* in, const V3<T>& w, V3<T>* out) {
  switch (model) {
    case kScaleOnly: *out = in[0] * w; return true;
    case kScaleAndBias: *out = {in[0] * w.x + in[1], in[0] * w.y + in[2], in[0] * w.z + in[3]}; return true;
    case kVectorNav: {
      const T &sx = in[0], &sy = in[1], &sz = in[2], &a1 = in[3], &a2 = in[4], &a3 = in[5], &a4 = in[6], &a5 = in[7], &a6 = in[8],
              &bx = in[9], &by = in[10], &bz = in[11];
      out->x = bx + sx * (w.x + a1 * w.y + a2 * w.z);
      out->y = by + sy * (w.y + a3 * w.x + a4 * w.z);
      out->z = bz + sz * (w.z + a5 * w.x + a6 * w.y);
      return true;
    }
    default: return false;
  }
}

// ---------------------------------------------------------------------------------------
// Cost functors. Parameter block order follows the reference enums.
// ---------------------------------------------------------------------------------------
// TrajectoryEvaluationParams (trajectory.h:17-24), frozen at construction from the stamp
// WITHOUT latency (camera_cost_functor.cpp:13-14; SURVEY §8 trap 1).
struct SegmentParams {
  int spline_index = -1;
  double knot0 = 0, knot1 = 0, stamp = 0;
  int k = 6;
  const double* basis = nullptr;  // k*k row-major
};

// CameraCostFunctor::operator(), camera_cost_functor.h:72-147. Parameter blocks
// (CameraParameterIndices, camera_cost_functor.h:12-32): 0 intrinsics, 1 q_rc (x,y,z,w),
// 2 t_rc, 3 latency, 4 model point, 5 q_wm, 6 t_wm, 7.. k control points of 6.
struct CameraFunctor {
  int model; double pixel[2]; double information; SegmentParams seg;
  static constexpr int kNumResiduals = 2;
  template <typename T> bool operator()(T const* const* P, T* residual) const {
    const T* intr = P[0];
    const Qt<T> q_rc{P[1][0], P[1][1], P[1][2], P[1][3]};
    const V3<T> t_rc{P[2][0], P[2][1], P[2][2]};
    const T latency = P[3][0];
    const V3<T> t_model_point{P[4][0], P[4][1], P[4][2]};
    const Qt<T> q_wm{P[5][0], P[5][1], P[5][2], P[5][3]};
    const V3<T> t_wm{P[6][0], P[6][1], P[6][2]};
    std::vector<T> ctrl(seg.k * 6);
    for (int i = 0; i < seg.k; ++i) for (int d = 0; d < 6; ++d) ctrl[i * 6 + d] = P[7 + i][d];
    const T knot0(seg.knot0), knot1(seg.knot1);
    const T stamp = T(seg.stamp) - latency;
    T pose[6];
    SplineEvaluate<T>(ctrl.data(), seg.k, knot0, knot1, seg.basis, stamp, 0, pose);
    const V3<T> phi_rw{-pose[0], -pose[1], -pose[2]};
    const Qt<T> q_rw = AngleAxisToQuaternion(phi_rw);
    const V3<T> t_wr{pose[3], pose[4], pose[5]};
    const Qt<T> q_cm = q_inverse(q_rc) * q_rw * q_wm;
    const V3<T> t_wc = t_wr + q_rotate(q_inverse(q_rw), t_rc);
    const V3<T> t_mc = q_rotate(q_inverse(q_wm), t_wc - t_wm);
    const V3<T> t_cp = q_rotate(q_cm, t_model_point - t_mc);
    T proj[2];
    if (!ProjectPoint<T>(model, intr, t_cp, proj)) return false;
    residual[0] = (T(pixel[0]) - proj[0]) * T(information);
    residual[1] = (T(pixel[1]) - proj[1]) * T(information);
    return true;
  }
};

// GyroscopeCostFunctor::operator(), gyroscope_cost_functor.h:59-118. Blocks
// (GyroscopeParameterIndices :13-24): 0 intrinsics, 1 q_rg, 2 t_rg (unused), 3 latency, 4.. control points.
struct GyroFunctor {
  int model; double meas[3]; double information; SegmentParams seg;
  static constexpr int kNumResiduals = 3;
  template <typename T> bool operator()(T const* const* P, T* residual) const {
    const T* intr = P[0];
    const Qt<T> q_rg{P[1][0], P[1][1], P[1][2], P[1][3]};
    const T latency = P[3][0];
    std::vector<T> ctrl(seg.k * 6);
    for (int i = 0; i < seg.k; ++i) for (int d = 0; d < 6; ++d) ctrl[i * 6 + d] = P[4 + i][d];
    const T knot0(seg.knot0), knot1(seg.knot1);
    const T stamp = T(seg.stamp) - latency;
    T pose[6], pose_dot[6];
    SplineEvaluate<T>(ctrl.data(), seg.k, knot0, knot1, seg.basis, stamp, 0, pose);
    SplineEvaluate<T>(ctrl.data(), seg.k, knot0, knot1, seg.basis, stamp, 1, pose_dot);
    const V3<T> phi{-pose[0], -pose[1], -pose[2]};
    const V3<T> phi_dot{-pose_dot[0], -pose_dot[1], -pose_dot[2]};
    const M3<T> J = ExpSO3Jacobian(phi);
    const V3<T> omega_rw = J * phi_dot;
    const V3<T> omega_g = -q_rotate(q_inverse(q_rg), omega_rw);
    V3<T> proj;
    if (!ImuProject<T>(model, intr, omega_g, &proj)) return false;
    residual[0] = (T(meas[0]) - proj.x) * T(information);
    residual[1] = (T(meas[1]) - proj.y) * T(information);
    residual[2] = (T(meas[2]) - proj.z) * T(information);
    return true;
  }
};

// AccelerometerCostFunctor::operator(), accelerometer_cost_functor.h:63-147. Blocks
// (AccelerometerParameterIndices :13-26): 0 intrinsics, 1 q_ra, 2 t_ra, 3 latency, 4 gravity, 5.. control points.
struct AccelFunctor {
  int model; double meas[3]; double information; SegmentParams seg;
  static constexpr int kNumResiduals = 3;
  template <typename T> bool operator()(T const* const* P, T* residual) const {
    const T* intr = P[0];
    const Qt<T> q_ra{P[1][0], P[1][1], P[1][2], P[1][3]};
    const V3<T> t_ra{P[2][0], P[2][1], P[2][2]};
    const T latency = P[3][0];
    const V3<T> gravity{P[4][0], P[4][1], P[4][2]};
    std::vector<T> ctrl(seg.k * 6);
    for (int i = 0; i < seg.k; ++i) for (int d = 0; d < 6; ++d) ctrl[i * 6 + d] = P[5 + i][d];
    const T knot0(seg.knot0), knot1(seg.knot1);
    const T stamp = T(seg.stamp) - latency;
    T pose[6], pose_dot[6], pose_ddot[6];
    SplineEvaluate<T>(ctrl.data(), seg.k, knot0, knot1, seg.basis, stamp, 0, pose);
    SplineEvaluate<T>(ctrl.data(), seg.k, knot0, knot1, seg.basis, stamp, 1, pose_dot);
    SplineEvaluate<T>(ctrl.data(), seg.k, knot0, knot1, seg.basis, stamp, 2, pose_ddot);
    const V3<T> phi{-pose[0], -pose[1], -pose[2]};
    const V3<T> phi_dot{-pose_dot[0], -pose_dot[1], -pose_dot[2]};
    const V3<T> phi_ddot{-pose_ddot[0], -pose_ddot[1], -pose_ddot[2]};
    const V3<T> ddt_wr{pose_ddot[3], pose_ddot[4], pose_ddot[5]};
    const Qt<T> q_rw = AngleAxisToQuaternion(phi);
    const M3<T> J = ExpSO3Jacobian(phi);
    const M3<T> Jdot = ExpSO3JacobianDot(phi, phi_dot);
    const V3<T> omega = J * phi_dot;
    const V3<T> alpha = Jdot * phi_dot + J * phi_ddot;
    const M3<T> Alpha = -Skew(alpha);
    const M3<T> Omega = -Skew(omega);
    const V3<T> a = q_rotate(q_inverse(q_ra), q_rotate(q_rw, ddt_wr - gravity) + (Omega * Omega + Alpha) * t_ra);
    V3<T> proj;
    if (!ImuProject<T>(model, intr, a, &proj)) return false;
    residual[0] = (T(meas[0]) - proj.x) * T(information);
    residual[1] = (T(meas[1]) - proj.y) * T(information);
    residual[2] = (T(meas[2]) - proj.z) * T(information);
    return true;
  }
};

}  // namespace orc
