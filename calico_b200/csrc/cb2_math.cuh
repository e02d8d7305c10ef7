// calico_b200 — small fixed-size FP64 algebra and the SO(3) pieces of the residual functors, with their analytic
// derivatives. Host+device so that the derivative formulas can be checked on the CPU against dual numbers.
//
// Reference functions restated here (paths relative to the reference tree):
//   Skew                       calico/geometry.h:12-23
//   ExpSO3Jacobian  (J_l)      calico/geometry.h:138-161
//   ExpSO3Hessian/JacobianDot  calico/geometry.h:173-222  (non-standard c0,c2 kept verbatim — SURVEY §8 trap 2)
//   ceres::AngleAxisToQuaternion (Ceres external; call sites camera_cost_functor.h:122, accelerometer_cost_functor.h:115)
//   Eigen quaternion rotate/inverse semantics as used through calico/typedefs.h:39-153
#pragma once
#include <math.h>

#include "cb2_platform.h"

namespace cb2 {

struct V3 { double x, y, z; };
struct M3 { double m[9]; };  // row-major

CB2_HD V3 v3(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
CB2_HD V3 operator+(const V3& a, const V3& b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
CB2_HD V3 operator-(const V3& a, const V3& b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
CB2_HD V3 operator-(const V3& a) { return v3(-a.x, -a.y, -a.z); }
CB2_HD V3 operator*(double s, const V3& a) { return v3(s * a.x, s * a.y, s * a.z); }
CB2_HD double dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
CB2_HD V3 cross(const V3& a, const V3& b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
CB2_HD double get(const V3& a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

CB2_HD M3 m3_zero() { M3 r; for (int i = 0; i < 9; ++i) r.m[i] = 0.0; return r; }
CB2_HD M3 m3_identity() { M3 r = m3_zero(); r.m[0] = r.m[4] = r.m[8] = 1.0; return r; }
CB2_HD M3 operator+(const M3& a, const M3& b) { M3 r; for (int i = 0; i < 9; ++i) r.m[i] = a.m[i] + b.m[i]; return r; }
CB2_HD M3 operator-(const M3& a, const M3& b) { M3 r; for (int i = 0; i < 9; ++i) r.m[i] = a.m[i] - b.m[i]; return r; }
CB2_HD M3 operator-(const M3& a) { M3 r; for (int i = 0; i < 9; ++i) r.m[i] = -a.m[i]; return r; }
CB2_HD M3 operator*(double s, const M3& a) { M3 r; for (int i = 0; i < 9; ++i) r.m[i] = s * a.m[i]; return r; }
CB2_HD M3 operator*(const M3& a, const M3& b) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[3 * i + j] = a.m[3 * i] * b.m[j] + a.m[3 * i + 1] * b.m[3 + j] + a.m[3 * i + 2] * b.m[6 + j];
  return r;
}
CB2_HD V3 operator*(const M3& a, const V3& v) {
  return v3(a.m[0] * v.x + a.m[1] * v.y + a.m[2] * v.z, a.m[3] * v.x + a.m[4] * v.y + a.m[5] * v.z, a.m[6] * v.x + a.m[7] * v.y + a.m[8] * v.z);
}
CB2_HD M3 transpose(const M3& a) { M3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[3 * i + j] = a.m[3 * j + i]; return r; }
// a * b^T (outer product)
CB2_HD M3 outer(const V3& a, const V3& b) {
  M3 r;
  r.m[0] = a.x * b.x; r.m[1] = a.x * b.y; r.m[2] = a.x * b.z;
  r.m[3] = a.y * b.x; r.m[4] = a.y * b.y; r.m[5] = a.y * b.z;
  r.m[6] = a.z * b.x; r.m[7] = a.z * b.y; r.m[8] = a.z * b.z;
  return r;
}
// geometry.h:12-23
CB2_HD M3 skew(const V3& v) {
  M3 r = m3_zero();
  r.m[1] = -v.z; r.m[3] = v.z; r.m[2] = v.y; r.m[6] = -v.y; r.m[5] = -v.x; r.m[7] = v.x;
  return r;
}

// Quaternion storage x,y,z,w (Eigen coeffs(), typedefs.h:69-81).
struct Q4 { double x, y, z, w; };
// Rotation matrix of Eigen's q * v (QuaternionBase::_transformVector: v + 2w (qv x v) + 2 qv x (qv x v)).
CB2_HD M3 quat_matrix(const Q4& q) {
  const M3 S = skew(v3(q.x, q.y, q.z));
  return m3_identity() + (2.0 * q.w) * S + 2.0 * (S * S);
}
// Eigen QuaternionBase::inverse(): conjugate / squaredNorm.
CB2_HD Q4 quat_inverse(const Q4& q) {
  const double n2 = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
  Q4 r;
  if (n2 > 0.0) { const double i = 1.0 / n2; r.x = -q.x * i; r.y = -q.y * i; r.z = -q.z * i; r.w = q.w * i; }
  else { r.x = r.y = r.z = r.w = 0.0; }
  return r;
}
CB2_HD Q4 quat_mul(const Q4& a, const Q4& b) {
  Q4 r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}
// ceres::AngleAxisToQuaternion (Ceres external): theta^2 > 0 -> [cos(t/2), sin(t/2)/t * aa]; else [1, aa/2].
CB2_HD Q4 angle_axis_to_quat(const V3& aa) {
  const double t2 = dot(aa, aa);
  Q4 q;
  if (t2 > 0.0) {
    const double t = sqrt(t2), h = 0.5 * t;
    double s, c;
    sincos(h, &s, &c);
    const double k = s / t;
    q.w = c; q.x = aa.x * k; q.y = aa.y * k; q.z = aa.z * k;
  } else {
    q.w = 1.0; q.x = 0.5 * aa.x; q.y = 0.5 * aa.y; q.z = 0.5 * aa.z;
  }
  return q;
}
// ceres::EigenQuaternionManifold::Plus (Ceres external): [sin|d| d/|d|, cos|d|] (x) q — left-multiplicative, d a half-angle.
CB2_HD Q4 quat_plus(const Q4& q, const V3& d) {
  const double nd = sqrt(dot(d, d));
  if (nd == 0.0) return q;
  double s, c;
  sincos(nd, &s, &c);
  const double k = s / nd;
  Q4 e; e.x = k * d.x; e.y = k * d.y; e.z = k * d.z; e.w = c;
  return quat_mul(e, q);
}

// geometry.h:36-51 small-angle Taylor series (used below 1e-7 only).
CB2_HD double small_angle_sin(double t) { const double q = t * t; return t * (1.0 - q * (1.0 / 6.0 + q * (1.0 / 120.0 - q * (1.0 / 5040.0)))); }
CB2_HD double small_angle_cos(double t) { const double q = t * t; return 1.0 - q * (0.5 - q * (1.0 / 24.0 + q * (1.0 / 720.0 - q * (1.0 / 40320.0)))); }

// Scalar coefficients of J_l(phi) = I + a [phi]x + b [phi]x^2 and what their derivatives need.
//   a = (1 - cos t)/t^2,  b = (t - sin t)/t^3,  da = (da/dt)/t,  db = (db/dt)/t.
struct SO3Coef { double theta, a, b, da, db; bool zero; };
CB2_HD SO3Coef so3_coef(const V3& phi) {
  SO3Coef c;
  const double t2 = dot(phi, phi);
  c.zero = (t2 == 0.0);
  if (c.zero) { c.theta = 0.0; c.a = c.b = c.da = c.db = 0.0; return c; }  // geometry.h:141-144: J = I, a constant
  const double t = sqrt(t2);
  c.theta = t;
  if (t < 0.05) {
    // Series: the closed forms cancel catastrophically for small angles.
    c.a = 0.5 - t2 * (1.0 / 24.0 - t2 * (1.0 / 720.0 - t2 * (1.0 / 40320.0)));
    c.b = 1.0 / 6.0 - t2 * (1.0 / 120.0 - t2 * (1.0 / 5040.0 - t2 * (1.0 / 362880.0)));
    c.da = -1.0 / 12.0 + t2 * (1.0 / 180.0 - t2 * (1.0 / 6720.0 - t2 * (1.0 / 453600.0)));
    c.db = -1.0 / 60.0 + t2 * (1.0 / 1260.0 - t2 * (1.0 / 60480.0 - t2 * (1.0 / 4989600.0)));
  } else {
    double s, co;
    sincos(t, &s, &co);
    const double it2 = 1.0 / t2;
    c.a = (1.0 - co) * it2;
    c.b = (t - s) * it2 / t;
    c.da = (t * s - 2.0 * (1.0 - co)) * it2 * it2;                    // (da/dt)/t
    c.db = ((1.0 - co) * t - 3.0 * (t - s)) * it2 * it2 / t;          // (db/dt)/t
  }
  return c;
}
// ExpSO3Jacobian, geometry.h:138-161.
CB2_HD M3 so3_jacobian(const V3& phi, const SO3Coef& c) {
  if (c.zero) return m3_identity();
  const M3 P = skew(phi);
  return m3_identity() + c.a * P + c.b * (P * P);
}
// d( J_l(phi) v ) / d phi — what forward-mode differentiation of ExpSO3Jacobian(phi) * v yields.
CB2_HD M3 so3_jacobian_times_vec_dphi(const V3& phi, const V3& v, const SO3Coef& c) {
  if (c.zero) return m3_zero();
  const V3 pv = cross(phi, v);
  const V3 ppv = cross(phi, pv);
  const double pdv = dot(phi, v);
  M3 D = (-c.a) * skew(v) + c.da * outer(pv, phi) + c.db * outer(ppv, phi);
  // d(phi x (phi x v))/dphi = (phi.v) I + phi v^T - 2 v phi^T
  M3 E = pdv * m3_identity() + outer(phi, v) - 2.0 * outer(v, phi);
  return D + c.b * E;
}

// n(phi, phid) = ExpSO3JacobianDot(phi, phid) * phid with the reference's ExpSO3Hessian (geometry.h:173-222):
//   column i of Jdot = H_i phid, H_i = c0 ph_i [ph]x + c1 G_i + c2 ph_i [ph]x^2 + c3 (G_i [ph]x + [ph]x G_i), ph = phi/theta,
// which contracts to  n = (ph.phid) (c0 ph x phid + c2 ph x (ph x phid)) + c3 phid x (ph x phid).
// Returns n and its derivatives with respect to phi and phid.
CB2_HD V3 so3_jdot_phid(const V3& phi, const V3& phid, M3* dn_dphi, M3* dn_dphid) {
  const double t2 = dot(phi, phi);
  if (t2 == 0.0) { *dn_dphi = m3_zero(); *dn_dphid = m3_zero(); return v3(0, 0, 0); }  // geometry.h:180-183
  const double t = sqrt(t2);
  double st, ct;
  if (t < 1e-7) { ct = small_angle_cos(t); st = small_angle_sin(t); } else { sincos(t, &st, &ct); }
  const double it = 1.0 / t, it2 = it * it;
  const double c0 = ct - st * it;
  const double c2 = 3.0 * it2 * st - it * (ct - 2.0);
  const double c3 = it2 * (t - st);
  // d/dtheta of the coefficients exactly as the closed forms above define them.
  const double dc0 = -st - ct * it + st * it2;
  const double dc2 = 3.0 * ct * it2 - 6.0 * st * it2 * it + st * it + (ct - 2.0) * it2;
  const double dc3 = (1.0 - ct) * it2 - 2.0 * (t - st) * it2 * it;
  const V3 ph = it * phi;
  const double s = dot(ph, phid);
  const V3 A = cross(ph, phid);
  const V3 B = cross(ph, A);
  const V3 C = cross(phid, A);
  const V3 n = s * (c0 * A + c2 * B) + c3 * C;
  const M3 Sph = skew(ph), Spd = skew(phid), SA = skew(A);
  const V3 u = c0 * A + c2 * B;
  // d/dphid with ph fixed.
  *dn_dphid = outer(u, ph) + s * (c0 * Sph + c2 * (Sph * Sph)) + c3 * (Spd * Sph - SA);
  // d/dph with the coefficients fixed, then chain through ph = phi/theta and theta.
  const M3 dn_dph = outer(u, phid) + s * ((-c0) * Spd + c2 * (-SA - Sph * Spd)) - c3 * (Spd * Spd);
  const M3 dph_dphi = it * (m3_identity() - outer(ph, ph));
  const V3 dn_dt = s * (dc0 * A + dc2 * B) + dc3 * C;
  *dn_dphi = dn_dph * dph_dphi + outer(dn_dt, ph);
  return n;
}

}  // namespace cb2
