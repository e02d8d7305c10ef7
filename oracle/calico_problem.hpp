// ORACLE — TEST INFRASTRUCTURE ONLY (see jet.hpp header).
//
// CPU restatement of calico::BatchOptimizer::Optimize (batch_optimizer.cpp:53-81): problem
// assembly as the reference's Add{Parameters,Residuals}ToProblem do it, and the Ceres
// trust-region Levenberg-Marquardt loop that ceres::Solve runs for DefaultSolverOptions()
// (batch_optimizer.cpp:10-17).
//
// PARITY STATUS. The residual definitions are restated from in-tree reference source and are
// pinned by the reference's own test constants (tests/test_oracle_*.py). Ceres itself is an
// un-vendored, un-pinned dependency (CMakeLists.txt:15; API use implies >= 2.1) whose source is
// absent from /root/reference: the LM loop below restates Ceres's published algorithm
// (trust_region_minimizer.cc, levenberg_marquardt_strategy.cc, corrector.cc, manifold.h of
// Ceres 2.1/2.2). It is pinned only by (i) the reference's integration test acceptance
// (batch_optimizer_test.cpp:185-210) and (ii) the Ceres iteration log stored in
// demos/imu_camera_calibration.ipynb:350-442 (trust-region radius / ratio schedule).
// Per-iteration cost and the LM step sequence are otherwise PARITY UNPINNED.
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "calico_math.hpp"

namespace orc {

enum SensorType { kCamera = 0, kGyroscope = 1, kAccelerometer = 2 };
enum LossType { kLossNone = 0, kLossHuber = 1, kLossCauchy = 2 };  // optimization_utils.h:15-22
// absl::StatusCode values used by the reference (SURVEY §8b "Errors").
enum Status { kOk = 0, kInvalidArgument = 3, kFailedPrecondition = 9, kInternal = 13 };

struct RigidBody {  // world_model.h:41-69
  int id = 0;
  double q[4] = {0, 0, 0, 1};  // x,y,z,w
  double t[3] = {0, 0, 0};
  bool pose_const = true, model_const = true;
  std::vector<int> feature_ids;
  std::vector<double> pts;  // 3 per feature
};

struct Sensor {
  int type = kCamera, model = 0;
  std::string name;
  std::vector<double> intr;
  double q[4] = {0, 0, 0, 1};  // extrinsics rotation, x,y,z,w
  double t[3] = {0, 0, 0};
  double latency = 0.0;        // camera.h:176
  double sigma = 1.0;          // camera.h:177
  int loss_type = kLossNone;
  double loss_scale = 1.0;     // camera.h:179
  bool en_intr = false, en_extr = false, en_lat = false;
  // Observations.
  std::vector<double> stamp;
  std::vector<int> body_slot, feat_slot;  // camera: index into bodies / that body's points
  std::vector<double> meas;               // 2 (camera) or 3 (imu) per observation
  std::vector<uint8_t> outlier;
  // Filled by Optimize/UpdateResiduals (camera.cpp:70-80): un-robustified residuals.
  std::vector<double> residuals;
  std::vector<uint8_t> residual_valid;
  int m() const { return type == kCamera ? 2 : 3; }
  int n_obs() const { return int(stamp.size()); }
};

struct Options {  // the subset of ceres::Solver::Options the reference touches (calico.cpp:378-394)
  int max_num_iterations = 50;
  double function_tolerance = 1e-8;      // batch_optimizer.cpp:14
  double gradient_tolerance = 1e-10;     // Ceres default
  double parameter_tolerance = 1e-10;    // batch_optimizer.cpp:15
  double initial_trust_region_radius = 1e4;
  double max_trust_region_radius = 1e16;
  double min_trust_region_radius = 1e-32;
  double min_relative_decrease = 1e-3;
  double min_lm_diagonal = 1e-6;
  double max_lm_diagonal = 1e32;
  int max_num_consecutive_invalid_steps = 5;
  int jacobi_scaling = 1;
  int num_threads = 1;
  int minimizer_progress_to_stdout = 0;
  int linear_solver = 0;  // 0 dense normal Cholesky, 1 banded Schur (control points first), 2 Ceres-ordered dense Schur
};

enum Termination { kConvergence = 0, kNoConvergence = 1, kFailure = 2 };  // ceres::TerminationType order

struct IterationLog {
  int iteration; double cost, cost_change, gradient_max_norm, gradient_norm, step_norm, relative_decrease,
      trust_region_radius; int step_is_valid, step_is_successful; double iteration_time;
};

struct Summary {
  int termination_type = kFailure;
  double initial_cost = 0, final_cost = 0, fixed_cost = 0;
  int num_successful_steps = 0, num_unsuccessful_steps = 0, num_iterations = 0;
  int num_parameter_blocks = 0, num_parameters = 0, num_effective_parameters = 0;
  int num_residual_blocks = 0, num_residuals = 0;
  int num_parameter_blocks_reduced = 0, num_parameters_reduced = 0, num_effective_parameters_reduced = 0;
  int num_residual_blocks_reduced = 0, num_residuals_reduced = 0;
  double jacobian_time = 0, linear_solver_time = 0, total_time = 0;
  char message[256] = {0};
};

struct ParamBlock {
  double* ptr; int size; int tsize; bool constant; bool quaternion;
  int off = -1;       // offset into the reduced tangent vector, -1 if not in the reduced program
  int aoff = -1;      // offset into the reduced ambient vector
  bool referenced = false;
  bool is_control_point = false;
};

struct ResidualBlock {
  int sensor, obs; int m;
  std::vector<int> blocks;  // parameter block ids in the functor's order
  SegmentParams seg;
  std::vector<double> basis;  // own copy, as the functor copies the 6x6 matrix (camera_cost_functor.cpp:13)
};

// ceres::EigenQuaternionManifold (manifold.h, Ceres external): Plus(x, d) = [sin|d| d/|d|, cos|d|] (x) x,
// left-multiplicative, d a half-angle vector; storage x,y,z,w. SURVEY §8 trap 5.
inline void QuaternionPlus(const double* x, const double* d, double* out) {
  const double nd = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  if (nd == 0.0) { for (int i = 0; i < 4; ++i) out[i] = x[i]; return; }
  const double s = std::sin(nd) / nd;
  const Qt<double> qd{s * d[0], s * d[1], s * d[2], std::cos(nd)};
  const Qt<double> qx{x[0], x[1], x[2], x[3]};
  const Qt<double> r = qd * qx;
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}
// PlusJacobian at delta = 0: 4 x 3 row-major, rows in storage order x,y,z,w.
inline void QuaternionPlusJacobian(const double* q, double J[12]) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double R[4][3] = {{w, z, -y}, {-z, w, x}, {y, -x, w}, {-x, -y, -z}};
  for (int r = 0; r < 4; ++r) for (int c = 0; c < 3; ++c) J[r * 3 + c] = R[r][c];
}

// ceres::HuberLoss / ceres::CauchyLoss (loss_function.cc, Ceres external). rho[0..2].
inline void EvaluateLoss(int type, double a, double s, double rho[3]) {
  if (type == kLossHuber) {
    const double b = a * a;
    if (s > b) {
      const double r = std::sqrt(s);
      rho[0] = 2.0 * a * r - b;
      rho[1] = std::max(std::numeric_limits<double>::min(), a / r);
      rho[2] = -rho[1] / (2.0 * s);
    } else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
  } else if (type == kLossCauchy) {
    const double b = a * a, c = 1.0 / b;
    const double sum = 1.0 + s * c, inv = 1.0 / sum;
    rho[0] = b * std::log(sum);
    rho[1] = std::max(std::numeric_limits<double>::min(), inv);
    rho[2] = -c * (inv * inv);
  } else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
}

struct Problem {
  // --- user-facing state (what the reference keeps in Trajectory / WorldModel / Sensor objects) ---
  int k = 6;
  std::vector<double> knots;   // full knot vector (bspline.hpp:164-180)
  std::vector<double> ctrl;    // n_cp x 6
  double gravity[3] = {0, 0, -9.80665};  // world_model.h:78
  std::vector<RigidBody> bodies;
  std::vector<Sensor> sensors;
  std::string error;

  // --- derived ---
  std::vector<double> valid_knots;
  std::vector<std::vector<double>> basis;  // per valid segment
  std::vector<ParamBlock> blocks;
  std::vector<ResidualBlock> rblocks;
  std::vector<int> cp_block0;
  int n_cp() const { return int(ctrl.size() / 6); }

  // BSpline::ComputeKnotVector layout: valid knots are knots[k-1 .. n_knots-k] (bspline.hpp:164-180).
  void PrepareSpline() {
    const int deg = k - 1;
    valid_knots.assign(knots.begin() + deg, knots.end() - deg);
    const int nseg = int(valid_knots.size()) - 1;
    basis.resize(nseg);
    for (int i = 0; i < nseg; ++i) basis[i] = BasisMatrix(knots, k, i + deg);  // bspline.hpp:183-189
  }
  // BSpline::GetSplineIndex, bspline.hpp:139-151.
  int GetSplineIndex(double t) const {
    int idx = -1;
    if (t == valid_knots.back()) idx = int(valid_knots.size()) - 2;
    else if (t < valid_knots.back()) {
      auto it = std::upper_bound(valid_knots.begin(), valid_knots.end(), t);
      idx = int(it - valid_knots.begin()) - 1;
    }
    return idx;
  }
  // Trajectory::GetEvaluationParams, trajectory.cpp:63-79. Returns false where .at() would throw.
  bool GetEvaluationParams(double stamp, SegmentParams* sp) const {
    const int idx = GetSplineIndex(stamp);
    if (idx < 0 || idx >= int(basis.size())) return false;
    const int knot_idx = idx + (k - 1);
    sp->spline_index = idx; sp->knot0 = knots[knot_idx]; sp->knot1 = knots[knot_idx + 1];
    sp->stamp = stamp; sp->k = k; sp->basis = basis[idx].data();
    return true;
  }

  int AddBlock(double* ptr, int size, bool constant, bool quaternion = false, bool is_cp = false) {
    ParamBlock b{ptr, size, quaternion ? 3 : size, constant, quaternion};
    b.is_control_point = is_cp;
    blocks.push_back(b);
    return int(blocks.size()) - 1;
  }

  // Problem assembly: batch_optimizer.cpp:57-70 → world_model.cpp:40-77, bspline.hpp:10-17,
  // camera.cpp:92-153, gyroscope.cpp:10-54, accelerometer.cpp:10-56.
  int Build() {
    error.clear();
    blocks.clear(); rblocks.clear();
    if (knots.empty() || ctrl.empty()) { error = "Trajectory has not been set."; return kFailedPrecondition; }
    PrepareSpline();
    // World model (world_model.cpp:40-77): model points, pose (translation, then quaternion —
    // optimization_utils.h:51-61), gravity (always constant: world_model.cpp:79-81).
    std::vector<int> body_pt0(bodies.size()), body_t(bodies.size()), body_q(bodies.size());
    for (size_t b = 0; b < bodies.size(); ++b) {
      RigidBody& rb = bodies[b];
      body_pt0[b] = int(blocks.size());
      for (size_t i = 0; i < rb.feature_ids.size(); ++i) AddBlock(&rb.pts[3 * i], 3, rb.model_const);
      body_t[b] = AddBlock(rb.t, 3, rb.pose_const);
      body_q[b] = AddBlock(rb.q, 4, rb.pose_const, true);
    }
    const int gravity_block = AddBlock(gravity, 3, true);
    // Trajectory control points, all free (bspline.hpp:10-17).
    const int cp0 = int(blocks.size());
    for (int i = 0; i < n_cp(); ++i) AddBlock(&ctrl[6 * i], 6, false, false, true);
    // Sensors.
    for (size_t si = 0; si < sensors.size(); ++si) {
      Sensor& s = sensors[si];
      const int want = s.type == kCamera ? CameraModelNumParams(s.model) : ImuModelNumParams(s.model);
      if (want < 0) { error = "Cannot add sensor parameters. Model is not yet defined."; return kFailedPrecondition; }
      if (int(s.intr.size()) != want) { error = "Invalid number of intrinsics parameters."; return kInvalidArgument; }
      const int b_intr = AddBlock(s.intr.data(), int(s.intr.size()), !s.en_intr);
      const int b_t = AddBlock(s.t, 3, !s.en_extr);
      const int b_q = AddBlock(s.q, 4, !s.en_extr, true);
      const int b_lat = AddBlock(&s.latency, 1, !s.en_lat);
      for (int o = 0; o < s.n_obs(); ++o) {
        if (s.type == kCamera && !s.outlier.empty() && s.outlier[o]) continue;  // camera.cpp:121-124
        ResidualBlock rb;
        rb.sensor = int(si); rb.obs = o; rb.m = s.m();
        if (!GetEvaluationParams(s.stamp[o], &rb.seg)) {
          error = "Observation stamp is outside the valid knots of the trajectory.";  // .at() would throw
          return kInvalidArgument;
        }
        rb.basis.assign(rb.seg.basis, rb.seg.basis + k * k);
        rb.blocks = {b_intr, b_q, b_t, b_lat};
        if (s.type == kCamera) {
          const int bs = s.body_slot[o];
          if (bs < 0) {  // camera.cpp:125-131
            error = "Attempted to create cost function from an observation for a rigidbody that does not exist in the world model.";
            return kFailedPrecondition;
          }
          rb.blocks.push_back(body_pt0[bs] + s.feat_slot[o]);
          rb.blocks.push_back(body_q[bs]);
          rb.blocks.push_back(body_t[bs]);
        } else if (s.type == kAccelerometer) {
          rb.blocks.push_back(gravity_block);
        }
        for (int i = 0; i < k; ++i) rb.blocks.push_back(cp0 + rb.seg.spline_index + i);
        rblocks.push_back(std::move(rb));
      }
    }
    for (auto& rb : rblocks) rb.seg.basis = rb.basis.data();
    return kOk;
  }

  // Functor evaluation at the current state. If jac != nullptr it receives, for each parameter
  // block of the residual block, an m x tangent_size row-major Jacobian (nullptr entry = skip),
  // computed by 4-wide dual-number passes like ceres::DynamicAutoDiffCostFunction and projected
  // to the tangent space by the manifold's PlusJacobian (Ceres external).
  bool EvaluateBlock(const ResidualBlock& rb, double* r, double** jac) const {
    const Sensor& s = sensors[rb.sensor];
    const int nb = int(rb.blocks.size());
    const double info = s.sigma > 0.0 ? 1.0 / s.sigma : 1.0;  // camera_cost_functor.cpp:15
    CameraFunctor cf; GyroFunctor gf; AccelFunctor af;
    if (s.type == kCamera) { cf.model = s.model; cf.pixel[0] = s.meas[2 * rb.obs]; cf.pixel[1] = s.meas[2 * rb.obs + 1]; cf.information = info; cf.seg = rb.seg; }
    else if (s.type == kGyroscope) { gf.model = s.model; for (int i = 0; i < 3; ++i) gf.meas[i] = s.meas[3 * rb.obs + i]; gf.information = info; gf.seg = rb.seg; }
    else { af.model = s.model; for (int i = 0; i < 3; ++i) af.meas[i] = s.meas[3 * rb.obs + i]; af.information = info; af.seg = rb.seg; }
    auto call = [&](auto const* const* P, auto* res) -> bool {
      if (s.type == kCamera) return cf(P, res);
      if (s.type == kGyroscope) return gf(P, res);
      return af(P, res);
    };
    if (!jac) {
      const double* P[32];
      for (int i = 0; i < nb; ++i) P[i] = blocks[rb.blocks[i]].ptr;
      if (!call(P, r)) return false;
      for (int i = 0; i < rb.m; ++i) if (!std::isfinite(r[i])) return false;
      return true;
    }
    // Active ambient parameters (those with a requested Jacobian).
    typedef Jet<4> J4;
    std::vector<J4> storage;
    std::vector<int> start(nb);
    int total = 0;
    for (int i = 0; i < nb; ++i) { start[i] = total; total += blocks[rb.blocks[i]].size; }
    storage.resize(total);
    std::vector<int> active_index(total, -1);
    int n_active = 0;
    for (int i = 0; i < nb; ++i) if (jac[i]) for (int j = 0; j < blocks[rb.blocks[i]].size; ++j) active_index[start[i] + j] = n_active++;
    std::vector<double> ambient(size_t(rb.m) * std::max(n_active, 1), 0.0);  // m x n_active
    const J4* P[32];
    for (int i = 0; i < nb; ++i) P[i] = storage.data() + start[i];
    const int passes = std::max(1, (n_active + 3) / 4);
    J4 res[3];
    for (int pass = 0; pass < passes; ++pass) {
      for (int i = 0; i < nb; ++i) for (int j = 0; j < blocks[rb.blocks[i]].size; ++j) {
        J4 v(blocks[rb.blocks[i]].ptr[j]);
        const int ai = active_index[start[i] + j];
        if (ai >= pass * 4 && ai < pass * 4 + 4) v.v[ai - pass * 4] = 1.0;
        storage[start[i] + j] = v;
      }
      if (!call(P, res)) return false;
      for (int q = 0; q < rb.m; ++q) {
        r[q] = res[q].a;
        for (int d = 0; d < 4; ++d) { const int ai = pass * 4 + d; if (ai < n_active) ambient[size_t(q) * n_active + ai] = res[q].v[d]; }
      }
    }
    for (int i = 0; i < rb.m; ++i) if (!std::isfinite(r[i])) return false;
    for (int i = 0; i < nb; ++i) {
      if (!jac[i]) continue;
      const ParamBlock& pb = blocks[rb.blocks[i]];
      const int a0 = active_index[start[i]];
      if (pb.quaternion) {
        double PJ[12]; QuaternionPlusJacobian(pb.ptr, PJ);
        for (int q = 0; q < rb.m; ++q) for (int c = 0; c < 3; ++c) {
          double sacc = 0.0;
          for (int a = 0; a < 4; ++a) sacc += ambient[size_t(q) * n_active + a0 + a] * PJ[a * 3 + c];
          jac[i][q * 3 + c] = sacc;
        }
      } else {
        for (int q = 0; q < rb.m; ++q) for (int c = 0; c < pb.size; ++c) jac[i][q * pb.size + c] = ambient[size_t(q) * n_active + a0 + c];
      }
      for (int e = 0; e < rb.m * pb.tsize; ++e) if (!std::isfinite(jac[i][e])) return false;
    }
    return true;
  }

  // Sensor::UpdateResiduals (camera.cpp:70-80, gyroscope.cpp:171-182, accelerometer.cpp:58-69):
  // un-robustified residual per (non-outlier) observation. Returns kInternal on failure.
  int UpdateResiduals() {
    for (auto& s : sensors) { s.residuals.assign(size_t(s.n_obs()) * s.m(), 0.0); s.residual_valid.assign(s.n_obs(), 0); }
    for (const auto& rb : rblocks) {
      double r[3];
      Sensor& s = sensors[rb.sensor];
      if (!EvaluateBlock(rb, r, nullptr)) {
        const char* kind = s.type == kCamera ? "camera " : (s.type == kGyroscope ? "gyroscope " : "accelerometer ");
        error = std::string("Failed to update residual for ") + kind + s.name;
        return kInternal;
      }
      for (int i = 0; i < rb.m; ++i) s.residuals[size_t(rb.obs) * rb.m + i] = r[i];
      s.residual_valid[rb.obs] = 1;
    }
    return kOk;
  }
};

}  // namespace orc
