// calico_b200 — the three residual functors with ANALYTIC Jacobians (host+device).
//
// Each function evaluates one residual block exactly as the reference's templated functor does for T=double and,
// instead of differentiating it by 4-wide dual-number passes (ceres::DynamicAutoDiffCostFunction,
// camera_cost_functor.cpp:25), writes a COMPACT derivative record:
//     G_d = d r / d P_d   (P_d = d-th time derivative of the spline pose 6-vector), the basis weights w_d[0..5],
//     d r / d intrinsics, d r / d (extrinsic rotation tangent), d r / d (extrinsic translation), d r / d latency.
// The full Jacobian row over the 6 control points is  d r / d cp_i = sum_d G_d * w_d[i]  (jac_entry below).
//
// Reference functors restated (paths relative to the reference tree):
//   CameraCostFunctor::operator()        calico/sensors/camera_cost_functor.h:72-147
//   GyroscopeCostFunctor::operator()     calico/sensors/gyroscope_cost_functor.h:59-118
//   AccelerometerCostFunctor::operator() calico/sensors/accelerometer_cost_functor.h:63-147
//   BSpline<6,T>::Evaluate               calico/bspline.hpp:40-72
// The segment (knot0, knot1, basis matrix, control points) is the one frozen from the stamp WITHOUT latency
// (camera_cost_functor.cpp:13-14) and is evaluated at stamp - latency (camera_cost_functor.h:114-115).
// Rotation tangents are those of ceres::EigenQuaternionManifold: q' = [sin|d| d/|d|, cos|d|] (x) q.
#pragma once
#include "cb2_models.cuh"

namespace cb2 {

constexpr int kK = 6;  // spline order (calico/trajectory.h:28)

enum { kCamera = 0, kGyroscope = 1, kAccelerometer = 2 };
enum { kLossNone = 0, kLossHuber = 1, kLossCauchy = 2 };

// Field offsets of the compact record, per sensor kind. Ji (m x ni, row-major with stride ni) is last.
// `one` / `zero` hold the constants 1.0 / 0.0 so that every Jacobian entry is a fixed-length sum of products of two record fields
// (jac_terms below) and the expansion loop of the sweep kernel is branch-free.
struct CamRec { enum { r = 0, G0 = 2, Jq = 14, Jt = 20, Jl = 26, w0 = 28, one = 34, zero = 35, Ji = 36 }; };
struct GyrRec { enum { r = 0, G0 = 3, G1 = 12, Jq = 21, Jl = 30, w0 = 33, w1 = 39, one = 45, zero = 46, Ji = 47 }; };
struct AccRec { enum { r = 0, G0 = 3, G1 = 12, G2 = 21, Jq = 39, Jt = 48, Jl = 57, w0 = 60, w1 = 66, w2 = 72, one = 78, zero = 79, Ji = 80 }; };
CB2_HD int rec_size(int kind, int ni) { return kind == kCamera ? CamRec::Ji + 2 * ni : (kind == kGyroscope ? GyrRec::Ji + 3 * ni : AccRec::Ji + 3 * ni); }
CB2_HD int residual_dim(int kind) { return kind == kCamera ? 2 : 3; }

// A record sink: field f of this block lives at base[f * stride] (stride = tile width in shared memory, 1 on the host).
struct Rec {
  double* base; int stride;
  CB2_HD void put(int f, double v) const { base[f * stride] = v; }
  CB2_HD double get(int f) const { return base[f * stride]; }
};

// ceres::HuberLoss / ceres::CauchyLoss (Ceres external; chosen by optimization_utils.h:31-47). rho(s) and rho'(s).
CB2_HD void loss_eval(int type, double a, double s, double* rho0, double* rho1) {
  if (type == kLossHuber) {
    const double b = a * a;
    if (s > b) { const double r = sqrt(s); *rho0 = 2.0 * a * r - b; *rho1 = fmax(2.2250738585072014e-308, a / r); }
    else { *rho0 = s; *rho1 = 1.0; }
  } else if (type == kLossCauchy) {
    const double b = a * a, c = 1.0 / b;
    const double sum = 1.0 + s * c;
    *rho0 = b * log(sum); *rho1 = fmax(2.2250738585072014e-308, 1.0 / sum);
  } else { *rho0 = s; *rho1 = 1.0; }
}

// Basis weights w[d][c] = (U_d * M)[c] for derivative orders d < ND (bspline.hpp:40-72):
// U_d[i] = i!/(i-d)! u^(i-d) dt^-d for i >= d, u = (t - knot0)/(knot1 - knot0). M is the 6x6 row-major basis matrix.
// Falling factorial i (i-1) ... (i-d+1) = i! / (i-d)!, a compile-time constant once the loops below are unrolled.
CB2_HD constexpr double falling_factorial(int i, int d) { double c = 1.0; for (int j = i - d; j < i; ++j) c *= double(j + 1); return c; }
template <int ND>
CB2_HD void spline_weights(const double* __restrict__ M, double knot0, double knot1, double t, double w[ND][kK]) {
  const double dt_inv = 1.0 / (knot1 - knot0);
  const double u = (t - knot0) * dt_inv;
  double pw[kK];
  pw[0] = 1.0;
#pragma unroll
  for (int i = 1; i < kK; ++i) pw[i] = pw[i - 1] * u;
  double Mr[kK * kK];
#pragma unroll
  for (int i = 0; i < kK * kK; ++i) Mr[i] = M[i];
  double scale = 1.0;
#pragma unroll
  for (int d = 0; d < ND; ++d) {
#pragma unroll
    for (int c = 0; c < kK; ++c) {
      double s = 0.0;
#pragma unroll
      for (int i = d; i < kK; ++i) s += (falling_factorial(i, d) * pw[i - d]) * Mr[i * kK + c];
      w[d][c] = s * scale;
    }
    scale *= dt_inv;
  }
}
// P[j] = sum_c w[c] * cp[c][j] for j in [j0, j1)
CB2_HD void spline_combine(const double w[kK], const double* __restrict__ cp, int j0, int j1, double* P) {
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    if (j < j0 || j >= j1) continue;
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < kK; ++c) s += w[c] * cp[c * 6 + j];
    P[j] = s;
  }
}
// Rodrigues rotation of the angle-axis phi (== matrix of ceres::AngleAxisToQuaternion(phi) up to rounding).
CB2_HD M3 so3_exp(const V3& phi, const SO3Coef& c) {
  if (c.zero) return m3_identity();
  double sa;
  const double t = c.theta;
  if (t < 0.05) { const double t2 = t * t; sa = 1.0 - t2 * (1.0 / 6.0 - t2 * (1.0 / 120.0 - t2 * (1.0 / 5040.0))); }
  else sa = sin(t) / t;
  const M3 P = skew(phi);
  return m3_identity() + sa * P + c.a * (P * P);
}

struct SensorState {  // one sensor's parameters, as Ceres sees them
  int kind, model, ni;
  double intr[kMaxIntrinsics];
  Q4 q;        // extrinsic rotation q_sensorrig_sensor, x,y,z,w
  V3 t;        // extrinsic translation
  double latency;
  double inv_sigma;   // camera_cost_functor.cpp:15
  int loss_type; double loss_scale;
};

// ---- camera: r = (pixel - Project(intr, p_c)) / sigma,  p_c = R_rc^T (R_rw (p_w - t_wr) - t_rc) ----
// Everything that depends only on (sensor, stamp) — basis weights, spline pose and its time derivative, R_rw, J_l(phi), R_rc^T —
// is shared by all corners of one image (25..144 residual blocks). It is computed once per image into a FRAME RECORD
// (camera_frame) and each residual block only does the point-specific part (camera_block_from_frame).
struct FrameRec { enum { w0 = 0, P1 = 6, R_rw = 12, Jl = 21, t_wr = 30, R_cr = 33, kSize = 44 }; };   // doubles, padded to 16 bytes

CB2_HD void camera_frame(const SensorState& S, const double* __restrict__ M, double knot0, double knot1, const double* __restrict__ cp,
                         double stamp, double* __restrict__ fr) {
  double w[2][kK];
  spline_weights<2>(M, knot0, knot1, stamp - S.latency, w);
  double P0[6], P1[6];
  spline_combine(w[0], cp, 0, 6, P0);
  spline_combine(w[1], cp, 0, 6, P1);
  const V3 phi = v3(-P0[0], -P0[1], -P0[2]);
  const SO3Coef c = so3_coef(phi);
  const M3 R_rw = so3_exp(phi, c);
  const M3 Jl = so3_jacobian(phi, c);
  const M3 R_cr = quat_matrix(quat_inverse(S.q));   // R_rc^T
  for (int i = 0; i < 6; ++i) { fr[FrameRec::w0 + i] = w[0][i]; fr[FrameRec::P1 + i] = P1[i]; }
  for (int i = 0; i < 9; ++i) { fr[FrameRec::R_rw + i] = R_rw.m[i]; fr[FrameRec::Jl + i] = Jl.m[i]; fr[FrameRec::R_cr + i] = R_cr.m[i]; }
  fr[FrameRec::t_wr] = P0[3]; fr[FrameRec::t_wr + 1] = P0[4]; fr[FrameRec::t_wr + 2] = P0[5];
}

template <bool kJac>
CB2_HD bool camera_block_from_frame(const SensorState& S, const double* __restrict__ fr, double px, double py, const V3& p_w, const Rec& out,
                                    double* pc_z = nullptr) {
  M3 R_rw, R_cr;
  for (int i = 0; i < 9; ++i) { R_rw.m[i] = fr[FrameRec::R_rw + i]; R_cr.m[i] = fr[FrameRec::R_cr + i]; }
  const V3 t_wr = v3(fr[FrameRec::t_wr], fr[FrameRec::t_wr + 1], fr[FrameRec::t_wr + 2]);
  const V3 a = R_rw * (p_w - t_wr);
  const V3 b = a - S.t;
  const V3 pc = R_cr * b;
  if (pc_z) *pc_z = pc.z;
  double uv[2], dp[2][3], di[2][kMaxIntrinsics];
  if (!camera_project<kJac>(S.model, S.intr, pc, uv, dp, di)) return false;
  const double r0 = (px - uv[0]) * S.inv_sigma, r1 = (py - uv[1]) * S.inv_sigma;
  if (!isfinite(r0) || !isfinite(r1)) return false;
  out.put(CamRec::r, r0); out.put(CamRec::r + 1, r1);
  if (kJac) {
    M3 Jl;
    for (int i = 0; i < 9; ++i) Jl.m[i] = fr[FrameRec::Jl + i];
    const M3 AJ = skew(a) * Jl;
    for (int row = 0; row < 2; ++row) {
      const V3 D = v3(-S.inv_sigma * dp[row][0], -S.inv_sigma * dp[row][1], -S.inv_sigma * dp[row][2]);   // dr/dp_c
      // DR = D^T R_cr (row vector)
      const V3 DR = v3(D.x * R_cr.m[0] + D.y * R_cr.m[3] + D.z * R_cr.m[6], D.x * R_cr.m[1] + D.y * R_cr.m[4] + D.z * R_cr.m[7],
                       D.x * R_cr.m[2] + D.y * R_cr.m[5] + D.z * R_cr.m[8]);
      double g[6];
      for (int j = 0; j < 3; ++j) {
        g[j] = DR.x * AJ.m[j] + DR.y * AJ.m[3 + j] + DR.z * AJ.m[6 + j];
        g[3 + j] = -(DR.x * R_rw.m[j] + DR.y * R_rw.m[3 + j] + DR.z * R_rw.m[6 + j]);
      }
      double jl = 0.0;
      for (int j = 0; j < 6; ++j) { out.put(CamRec::G0 + row * 6 + j, g[j]); jl -= g[j] * fr[FrameRec::P1 + j]; }
      out.put(CamRec::Jl + row, jl);
      out.put(CamRec::Jt + row * 3 + 0, -DR.x); out.put(CamRec::Jt + row * 3 + 1, -DR.y); out.put(CamRec::Jt + row * 3 + 2, -DR.z);
      const V3 jq = cross(DR, b);   // DR^T [b]x = (DR x b)^T
      out.put(CamRec::Jq + row * 3 + 0, 2.0 * jq.x); out.put(CamRec::Jq + row * 3 + 1, 2.0 * jq.y); out.put(CamRec::Jq + row * 3 + 2, 2.0 * jq.z);
#pragma unroll
      for (int j = 0; j < kMaxIntrinsics; ++j) if (j < S.ni) out.put(CamRec::Ji + row * S.ni + j, -S.inv_sigma * di[row][j]);   // constant indices: di stays in registers
    }
    for (int i = 0; i < kK; ++i) out.put(CamRec::w0 + i, fr[FrameRec::w0 + i]);
    out.put(CamRec::one, 1.0); out.put(CamRec::zero, 0.0);
  }
  return true;
}

// One residual block evaluated on its own (frame record on the stack); the kernels share the record between the corners of an image.
template <bool kJac>
CB2_HD bool camera_block(const SensorState& S, const double* __restrict__ M, double knot0, double knot1, const double* __restrict__ cp,
                         double stamp, double px, double py, const V3& p_w, const Rec& out) {
  double fr[FrameRec::kSize];
  camera_frame(S, M, knot0, knot1, cp, stamp, fr);
  return camera_block_from_frame<kJac>(S, fr, px, py, p_w, out);
}

// ---- gyroscope: r = (meas - Project(intr, -R_rg^T J_l(phi) phid)) / sigma ----
template <bool kJac>
CB2_HD bool gyro_block(const SensorState& S, const double* __restrict__ M, double knot0, double knot1, const double* __restrict__ cp,
                       double stamp, const V3& meas, const Rec& out) {
  double w[3][kK];
  spline_weights<3>(M, knot0, knot1, stamp - S.latency, w);
  double P0[6], P1[6];
  spline_combine(w[0], cp, 0, 3, P0);
  spline_combine(w[1], cp, 0, 3, P1);
  const V3 phi = v3(-P0[0], -P0[1], -P0[2]);
  const V3 phid = v3(-P1[0], -P1[1], -P1[2]);
  const SO3Coef c = so3_coef(phi);
  const M3 J = so3_jacobian(phi, c);
  const V3 omega = J * phid;
  const M3 R_gr = quat_matrix(quat_inverse(S.q));
  const V3 omega_g = -(R_gr * omega);
  V3 proj; M3 dw; double di[3][kMaxIntrinsics];
  if (!imu_project<kJac>(S.model, S.intr, omega_g, &proj, &dw, di)) return false;
  const V3 r = S.inv_sigma * (meas - proj);
  if (!isfinite(r.x) || !isfinite(r.y) || !isfinite(r.z)) return false;
  out.put(GyrRec::r, r.x); out.put(GyrRec::r + 1, r.y); out.put(GyrRec::r + 2, r.z);
  if (kJac) {
    double P2[6];
    spline_combine(w[2], cp, 0, 3, P2);
    const M3 K = (-S.inv_sigma) * dw;            // dr / d omega_g
    const M3 Kw = -(K * R_gr);                   // dr / d omega
    const M3 G0 = -(Kw * so3_jacobian_times_vec_dphi(phi, phid, c));
    const M3 G1 = -(Kw * J);
    const M3 Jq = (-2.0) * (K * (R_gr * skew(omega)));
    for (int i = 0; i < 9; ++i) { out.put(GyrRec::G0 + i, G0.m[i]); out.put(GyrRec::G1 + i, G1.m[i]); out.put(GyrRec::Jq + i, Jq.m[i]); }
    for (int row = 0; row < 3; ++row) {
      double jl = 0.0;
      for (int j = 0; j < 3; ++j) jl -= G0.m[row * 3 + j] * P1[j] + G1.m[row * 3 + j] * P2[j];
      out.put(GyrRec::Jl + row, jl);
#pragma unroll
      for (int j = 0; j < kMaxIntrinsics; ++j) if (j < S.ni) out.put(GyrRec::Ji + row * S.ni + j, -S.inv_sigma * di[row][j]);   // constant indices: di stays in registers
    }
    for (int i = 0; i < kK; ++i) { out.put(GyrRec::w0 + i, w[0][i]); out.put(GyrRec::w1 + i, w[1][i]); }
    out.put(GyrRec::one, 1.0); out.put(GyrRec::zero, 0.0);
  }
  return true;
}

// ---- accelerometer: r = (meas - Project(intr, R_ra^T (R_rw (tdd - g) + (Omega^2 + Alpha) t_ra))) / sigma ----
template <bool kJac>
CB2_HD bool accel_block(const SensorState& S, const V3& gravity, const double* __restrict__ M, double knot0, double knot1,
                        const double* __restrict__ cp, double stamp, const V3& meas, const Rec& out) {
  double w[4][kK];
  spline_weights<(kJac ? 4 : 3)>(M, knot0, knot1, stamp - S.latency, w);
  double P0[6], P1[6], P2[6];
  spline_combine(w[0], cp, 0, 3, P0);
  spline_combine(w[1], cp, 0, 3, P1);
  spline_combine(w[2], cp, 0, 6, P2);
  const V3 phi = v3(-P0[0], -P0[1], -P0[2]);
  const V3 phid = v3(-P1[0], -P1[1], -P1[2]);
  const V3 phidd = v3(-P2[0], -P2[1], -P2[2]);
  const V3 tdd = v3(P2[3], P2[4], P2[5]);
  const SO3Coef c = so3_coef(phi);
  const M3 R_rw = so3_exp(phi, c);
  const M3 J = so3_jacobian(phi, c);
  const V3 omega = J * phid;
  M3 dn_dphi, dn_dphid;
  const V3 n = so3_jdot_phid(phi, phid, &dn_dphi, &dn_dphid);
  const V3 alpha = n + J * phidd;
  const V3 a1 = R_rw * (tdd - gravity);
  const V3 a_r = a1 + cross(omega, cross(omega, S.t)) - cross(alpha, S.t);
  const M3 R_ar = quat_matrix(quat_inverse(S.q));
  const V3 acc = R_ar * a_r;
  V3 proj; M3 dw; double di[3][kMaxIntrinsics];
  if (!imu_project<kJac>(S.model, S.intr, acc, &proj, &dw, di)) return false;
  const V3 r = S.inv_sigma * (meas - proj);
  if (!isfinite(r.x) || !isfinite(r.y) || !isfinite(r.z)) return false;
  out.put(AccRec::r, r.x); out.put(AccRec::r + 1, r.y); out.put(AccRec::r + 2, r.z);
  if (kJac) {
    double P3[6];
    spline_combine(w[3], cp, 0, 6, P3);
    const M3 K = (-S.inv_sigma) * dw;   // dr / d acc
    const M3 KR = K * R_ar;             // dr / d a_r
    const M3 Lw = dot(omega, S.t) * m3_identity() + outer(omega, S.t) - 2.0 * outer(S.t, omega);
    const M3 La = skew(S.t);
    const M3 D1 = so3_jacobian_times_vec_dphi(phi, phid, c);
    const M3 D2 = so3_jacobian_times_vec_dphi(phi, phidd, c);
    const M3 Fphi = -(skew(a1) * J) + Lw * D1 + La * (dn_dphi + D2);
    const M3 Fphid = Lw * J + La * dn_dphid;
    const M3 Fphidd = La * J;
    const M3 G0 = -(KR * Fphi), G1 = -(KR * Fphid), G2r = -(KR * Fphidd), G2t = KR * R_rw;
    const M3 So = skew(omega);
    const M3 Jt = KR * (So * So - skew(alpha));
    const M3 Jq = 2.0 * (KR * skew(a_r));
    for (int i = 0; i < 9; ++i) { out.put(AccRec::G0 + i, G0.m[i]); out.put(AccRec::G1 + i, G1.m[i]); out.put(AccRec::Jq + i, Jq.m[i]); out.put(AccRec::Jt + i, Jt.m[i]); }
    for (int row = 0; row < 3; ++row) {
      double jl = 0.0;
      for (int j = 0; j < 3; ++j) {
        out.put(AccRec::G2 + row * 6 + j, G2r.m[row * 3 + j]);
        out.put(AccRec::G2 + row * 6 + 3 + j, G2t.m[row * 3 + j]);
        jl -= G0.m[row * 3 + j] * P1[j] + G1.m[row * 3 + j] * P2[j] + G2r.m[row * 3 + j] * P3[j] + G2t.m[row * 3 + j] * P3[3 + j];
      }
      out.put(AccRec::Jl + row, jl);
#pragma unroll
      for (int j = 0; j < kMaxIntrinsics; ++j) if (j < S.ni) out.put(AccRec::Ji + row * S.ni + j, -S.inv_sigma * di[row][j]);   // constant indices: di stays in registers
    }
    for (int i = 0; i < kK; ++i) { out.put(AccRec::w0 + i, w[0][i]); out.put(AccRec::w1 + i, w[1][i]); out.put(AccRec::w2 + i, w[2][i]); }
    out.put(AccRec::one, 1.0); out.put(AccRec::zero, 0.0);
  }
  return true;
}

// Canonical Jacobian column order of one residual block (include/calico_b200.h, cb2_evaluate_sensor):
//   [control points 6k | intrinsics ni | extrinsic rotation 3 | extrinsic translation 3 | latency 1],  W = 6k + ni + 7.
// Entry (row, col) of the un-robustified Jacobian = sum_{q < kind_terms} rec[fa[q]] * rec[fb[q]]; unused terms point at `zero`.
CB2_HD constexpr int kind_terms(int kind) { return kind == kCamera ? 1 : (kind == kGyroscope ? 2 : 3); }
CB2_HD void jac_terms(int kind, int ni, int row, int col, int fa[3], int fb[3]) {
  const int one = kind == kCamera ? int(CamRec::one) : (kind == kGyroscope ? int(GyrRec::one) : int(AccRec::one));
  const int zero = one + 1;
  for (int q = 0; q < 3; ++q) { fa[q] = zero; fb[q] = zero; }
  if (col < 6 * kK) {
    const int i = col / 6, d = col % 6;
    if (kind == kCamera) { fa[0] = CamRec::G0 + row * 6 + d; fb[0] = CamRec::w0 + i; }
    else if (kind == kGyroscope) {
      if (d < 3) { fa[0] = GyrRec::G0 + row * 3 + d; fb[0] = GyrRec::w0 + i; fa[1] = GyrRec::G1 + row * 3 + d; fb[1] = GyrRec::w1 + i; }
    } else {
      fa[0] = AccRec::G2 + row * 6 + d; fb[0] = AccRec::w2 + i;
      if (d < 3) { fa[1] = AccRec::G0 + row * 3 + d; fb[1] = AccRec::w0 + i; fa[2] = AccRec::G1 + row * 3 + d; fb[2] = AccRec::w1 + i; }
    }
    return;
  }
  col -= 6 * kK;
  fb[0] = one;
  if (col < ni) { fa[0] = (kind == kCamera ? int(CamRec::Ji) : (kind == kGyroscope ? int(GyrRec::Ji) : int(AccRec::Ji))) + row * ni + col; return; }
  col -= ni;
  if (col < 3) { fa[0] = (kind == kCamera ? int(CamRec::Jq) : (kind == kGyroscope ? int(GyrRec::Jq) : int(AccRec::Jq))) + row * 3 + col; return; }
  col -= 3;
  if (col < 3) { if (kind != kGyroscope) fa[0] = (kind == kCamera ? int(CamRec::Jt) : int(AccRec::Jt)) + row * 3 + col; return; }
  fa[0] = (kind == kCamera ? int(CamRec::Jl) : (kind == kGyroscope ? int(GyrRec::Jl) : int(AccRec::Jl))) + row;
}
// Camera record field behind compact Gram column `col` of residual row q: [g_0..g_5 | r | 0 | calibration unknown col - 8] (SensorDesc::gslots).
// canon_of_unknown[u] = canonical calibration column (0 .. ni + 6) of calibration-local unknown u, -1 = not stored.
CB2_HD int gram_field(int ni, int q, int col, const int* canon_of_unknown, int n_unknowns) {
  if (col < 6) return CamRec::G0 + q * 6 + col;
  if (col == 6) return CamRec::r + q;
  if (col < 8 || col - 8 >= n_unknowns || canon_of_unknown[col - 8] < 0) return CamRec::zero;
  int fa[3], fb[3];
  jac_terms(kCamera, ni, q, 6 * kK + canon_of_unknown[col - 8], fa, fb);
  return fa[0];
}
CB2_HD double jac_entry(int kind, int ni, const Rec& rec, int row, int col) {
  int fa[3], fb[3];
  jac_terms(kind, ni, row, col, fa, fb);
  double v = 0.0;
  for (int q = 0; q < kind_terms(kind); ++q) v += rec.get(fa[q]) * rec.get(fb[q]);
  return v;
}

}  // namespace cb2
