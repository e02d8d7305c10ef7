// ORACLE — TEST INFRASTRUCTURE ONLY (see jet.hpp header). C entry points (ctypes) over the oracle.
// The shape mirrors include/calico_b200.h so that one problem description drives both the oracle
// and the product, but nothing here is used by the product.
#include <cstdint>
#include <cstring>
#include <unordered_map>

#include "calico_schur.hpp"

using namespace orc;

namespace {
struct Handle {
  Problem p;
  std::unordered_map<int, int> body_slot;                            // rigid body id → slot
  std::vector<std::unordered_map<int, int>> feat_slot;               // per body: feature id → slot
  std::string error;
  bool built = false;
};
int fail(Handle* h, int code, const std::string& msg) { h->error = msg; return code; }
}  // namespace

extern "C" {

int orc_problem_create(void** out) { *out = new Handle(); return kOk; }
void orc_problem_destroy(void* hp) { delete static_cast<Handle*>(hp); }
const char* orc_last_error(void* hp) { return static_cast<Handle*>(hp)->error.c_str(); }

int orc_set_trajectory(void* hp, int spline_order, int n_knots, const double* knots, int n_cp, const double* ctrl) {
  Handle* h = static_cast<Handle*>(hp);
  if (spline_order < 2) return fail(h, kInvalidArgument, "Spline order must be greater than 2.");
  if (n_knots != n_cp + spline_order) return fail(h, kInvalidArgument, "Knot vector size must equal control points + spline order.");
  h->p.k = spline_order;
  h->p.knots.assign(knots, knots + n_knots);
  h->p.ctrl.assign(ctrl, ctrl + size_t(n_cp) * 6);
  h->built = false;
  return kOk;
}
int orc_set_gravity(void* hp, const double* g) { Handle* h = static_cast<Handle*>(hp); std::memcpy(h->p.gravity, g, 24); return kOk; }

int orc_add_rigid_body(void* hp, int id, const double* q_xyzw, const double* t, int n_pts, const int* feature_ids, const double* pts,
                       int pose_const, int model_const) {
  Handle* h = static_cast<Handle*>(hp);
  if (h->body_slot.count(id)) return fail(h, kInvalidArgument, "Rigid body with id " + std::to_string(id) + " already exists in world model.");  // world_model.cpp:32-35
  RigidBody rb; rb.id = id; std::memcpy(rb.q, q_xyzw, 32); std::memcpy(rb.t, t, 24);
  rb.pose_const = pose_const != 0; rb.model_const = model_const != 0;
  rb.feature_ids.assign(feature_ids, feature_ids + n_pts); rb.pts.assign(pts, pts + size_t(n_pts) * 3);
  std::unordered_map<int, int> fs;
  for (int i = 0; i < n_pts; ++i) fs[feature_ids[i]] = i;
  h->body_slot[id] = int(h->p.bodies.size());
  h->p.bodies.push_back(std::move(rb));
  h->feat_slot.push_back(std::move(fs));
  h->built = false;
  return kOk;
}

int orc_add_sensor(void* hp, int type, int model, const char* name, int n_intr, const double* intr, const double* q_xyzw, const double* t,
                   double latency, double sigma, int loss_type, double loss_scale, int en_intr, int en_extr, int en_lat, int* sensor_id) {
  Handle* h = static_cast<Handle*>(hp);
  Sensor s; s.type = type; s.model = model; s.name = name ? name : "";
  s.intr.assign(intr, intr + n_intr); std::memcpy(s.q, q_xyzw, 32); std::memcpy(s.t, t, 24);
  s.latency = latency;
  if (sigma <= 0.0) return fail(h, kInvalidArgument, "Sigma must be greater than 0.");  // camera.cpp:62-65
  s.sigma = sigma; s.loss_type = loss_type; s.loss_scale = loss_scale;
  s.en_intr = en_intr; s.en_extr = en_extr; s.en_lat = en_lat;
  *sensor_id = int(h->p.sensors.size());
  h->p.sensors.push_back(std::move(s));
  h->built = false;
  return kOk;
}

int orc_add_camera_observations(void* hp, int sensor, int n, const double* stamp, const int* image_id, const int* model_id,
                                const int* feature_id, const double* pixel_xy, const uint8_t* outlier) {
  Handle* h = static_cast<Handle*>(hp);
  (void)image_id;
  if (sensor < 0 || sensor >= int(h->p.sensors.size()) || h->p.sensors[sensor].type != kCamera) return fail(h, kInvalidArgument, "Not a camera sensor id.");
  Sensor& s = h->p.sensors[sensor];
  for (int i = 0; i < n; ++i) {
    s.stamp.push_back(stamp[i]);
    auto it = h->body_slot.find(model_id[i]);
    int bs = -1, fs = -1;
    if (it != h->body_slot.end()) {
      bs = it->second;
      auto f = h->feat_slot[bs].find(feature_id[i]);
      if (f == h->feat_slot[bs].end()) return fail(h, kInvalidArgument, "Feature id not in rigid body model definition.");  // .at() would throw
      fs = f->second;
    }
    s.body_slot.push_back(bs); s.feat_slot.push_back(fs);
    s.meas.push_back(pixel_xy[2 * i]); s.meas.push_back(pixel_xy[2 * i + 1]);
    s.outlier.push_back(outlier ? outlier[i] : 0);
  }
  h->built = false;
  return kOk;
}

int orc_add_imu_observations(void* hp, int sensor, int n, const double* stamp, const int* seq, const double* xyz) {
  Handle* h = static_cast<Handle*>(hp);
  (void)seq;
  if (sensor < 0 || sensor >= int(h->p.sensors.size()) || h->p.sensors[sensor].type == kCamera) return fail(h, kInvalidArgument, "Not an IMU sensor id.");
  Sensor& s = h->p.sensors[sensor];
  for (int i = 0; i < n; ++i) { s.stamp.push_back(stamp[i]); for (int d = 0; d < 3; ++d) s.meas.push_back(xyz[3 * i + d]); s.outlier.push_back(0); }
  h->built = false;
  return kOk;
}

static int ensure_built(Handle* h) {
  const int rc = h->p.Build();
  if (rc != kOk) h->error = h->p.error;
  h->built = rc == kOk;
  return rc;
}

// calico::BatchOptimizer::Optimize (batch_optimizer.cpp:53-81): build, solve, UpdateResiduals.
int orc_optimize(void* hp, const Options* opt, Summary* summary, IterationLog* log, int cap, int* n_log) {
  Handle* h = static_cast<Handle*>(hp);
  int rc = ensure_built(h);
  if (rc != kOk) return rc;
  std::vector<IterationLog> L;
  *summary = Summary();
  Minimizer M(h->p, *opt, summary, &L);
  M.Minimize();
  if (n_log) *n_log = int(L.size());
  for (int i = 0; i < int(L.size()) && i < cap; ++i) log[i] = L[i];
  rc = h->p.UpdateResiduals();
  if (rc != kOk) h->error = h->p.error;
  return rc;
}

// Analogue of ceres::Problem::Evaluate, per sensor, in a canonical column order
// [control points 6k | intrinsics | extrinsic rotation (tangent 3) | extrinsic translation 3 | latency 1],
// every sensor block treated as non-constant, world model constant, loss NOT applied.
int orc_evaluate_sensor(void* hp, int sensor, double* residuals, double* jac, uint8_t* valid) {
  Handle* h = static_cast<Handle*>(hp);
  int rc = ensure_built(h);
  if (rc != kOk) return rc;
  const Sensor& s = h->p.sensors[sensor];
  const int k = h->p.k, m = s.m(), ni = int(s.intr.size());
  const int W = 6 * k + ni + 7;
  for (const auto& rb : h->p.rblocks) {
    if (rb.sensor != sensor) continue;
    const int nb = int(rb.blocks.size());
    double r[3];
    std::vector<std::vector<double>> store(nb);
    double* jp[32];
    const int cp_first = nb - k;
    for (int i = 0; i < nb; ++i) {
      const bool want = (i <= 3) || (i >= cp_first);
      if (want && jac) { store[i].assign(size_t(m) * h->p.blocks[rb.blocks[i]].tsize, 0.0); jp[i] = store[i].data(); } else jp[i] = nullptr;
    }
    const bool ok = h->p.EvaluateBlock(rb, r, jac ? jp : nullptr);
    if (valid) valid[rb.obs] = ok;
    if (!ok) continue;
    if (residuals) for (int q = 0; q < m; ++q) residuals[size_t(rb.obs) * m + q] = r[q];
    if (jac) {
      double* J = jac + size_t(rb.obs) * m * W;
      for (int q = 0; q < m; ++q) {
        for (int c = 0; c < k; ++c) for (int d = 0; d < 6; ++d) J[q * W + c * 6 + d] = store[cp_first + c][q * 6 + d];
        for (int c = 0; c < ni; ++c) J[q * W + 6 * k + c] = store[0][q * ni + c];
        for (int c = 0; c < 3; ++c) J[q * W + 6 * k + ni + c] = store[1][q * 3 + c];       // rotation (block 1 = quaternion)
        for (int c = 0; c < 3; ++c) J[q * W + 6 * k + ni + 3 + c] = store[2][q * 3 + c];   // translation
        J[q * W + 6 * k + ni + 6] = store[3][q];                                            // latency
      }
    }
  }
  return kOk;
}

// Total cost 1/2 sum rho(|r|^2) at the current state over all non-outlier blocks; *ok = 0 if any block fails.
int orc_cost(void* hp, double* cost, int* ok) {
  Handle* h = static_cast<Handle*>(hp);
  int rc = ensure_built(h);
  if (rc != kOk) return rc;
  double total = 0; *ok = 1;
  for (const auto& rb : h->p.rblocks) {
    double r[3];
    if (!h->p.EvaluateBlock(rb, r, nullptr)) { *ok = 0; continue; }
    double sq = 0; for (int q = 0; q < rb.m; ++q) sq += r[q] * r[q];
    double rho[3]; const Sensor& s = h->p.sensors[rb.sensor];
    EvaluateLoss(s.loss_type, s.loss_scale, sq, rho);
    total += 0.5 * rho[0];
  }
  *cost = total;
  return kOk;
}

int orc_get_sensor(void* hp, int sensor, double* intr, double* q_xyzw, double* t, double* latency) {
  Handle* h = static_cast<Handle*>(hp);
  const Sensor& s = h->p.sensors[sensor];
  std::memcpy(intr, s.intr.data(), s.intr.size() * 8); std::memcpy(q_xyzw, s.q, 32); std::memcpy(t, s.t, 24); *latency = s.latency;
  return kOk;
}
int orc_get_rigid_body(void* hp, int id, double* q, double* t, double* pts) {
  Handle* h = static_cast<Handle*>(hp);
  auto it = h->body_slot.find(id);
  if (it == h->body_slot.end()) return fail(h, kInvalidArgument, "Unknown rigid body id.");
  const RigidBody& rb = h->p.bodies[it->second];
  if (q) std::memcpy(q, rb.q, 32);
  if (t) std::memcpy(t, rb.t, 24);
  if (pts) std::memcpy(pts, rb.pts.data(), rb.pts.size() * 8);
  return kOk;
}
int orc_get_trajectory(void* hp, double* ctrl) { Handle* h = static_cast<Handle*>(hp); std::memcpy(ctrl, h->p.ctrl.data(), h->p.ctrl.size() * 8); return kOk; }
int orc_get_residuals(void* hp, int sensor, double* out, uint8_t* valid) {
  Handle* h = static_cast<Handle*>(hp);
  const Sensor& s = h->p.sensors[sensor];
  if (s.residuals.empty() && s.n_obs() > 0) return fail(h, kFailedPrecondition, "Residuals have not been computed.");
  std::memcpy(out, s.residuals.data(), s.residuals.size() * 8);
  if (valid) std::memcpy(valid, s.residual_valid.data(), s.residual_valid.size());
  return kOk;
}

// ---- Spline / kinematics probes used to pin the restatement against the reference's unit tests. ----
// BSpline::Interpolate (bspline.hpp:74-101): returns kInvalidArgument outside the valid knots or for a bad derivative.
int orc_spline_interpolate(void* hp, int n, const double* times, int derivative, double* out6) {
  Handle* h = static_cast<Handle*>(hp);
  Problem& p = h->p;
  if (p.knots.empty()) return fail(h, kFailedPrecondition, "Trajectory has not been set.");
  p.PrepareSpline();
  if (derivative < 0 || derivative > p.k - 1) return fail(h, kInvalidArgument, "Invalid derivative for interpolation.");
  for (int i = 0; i < n; ++i) if (times[i] < p.valid_knots.front() || times[i] > p.valid_knots.back()) return fail(h, kInvalidArgument, "Cannot interpolate. Value is not within valid knots.");
  for (int i = 0; i < n; ++i) {
    SegmentParams sp; p.GetEvaluationParams(times[i], &sp);
    SplineEvaluate<double>(&p.ctrl[size_t(sp.spline_index) * 6], p.k, sp.knot0, sp.knot1, sp.basis, times[i], derivative, out6 + 6 * i);
  }
  return kOk;
}
int orc_basis_matrix(int k, int n_knots, const double* knots, int i, double* out) {
  std::vector<double> kn(knots, knots + n_knots);
  const std::vector<double> M = BasisMatrix(kn, k, i);
  std::memcpy(out, M.data(), M.size() * 8);
  return kOk;
}
void orc_exp_so3(const double* phi, double* R9) { const M3<double> R = ExpSO3(V3<double>{phi[0], phi[1], phi[2]}); std::memcpy(R9, R.m, 72); }
void orc_exp_so3_jacobian(const double* phi, double* J9) { const M3<double> J = ExpSO3Jacobian(V3<double>{phi[0], phi[1], phi[2]}); std::memcpy(J9, J.m, 72); }
void orc_exp_so3_jacobian_dot(const double* phi, const double* phi_dot, double* J9) {
  const M3<double> J = ExpSO3JacobianDot(V3<double>{phi[0], phi[1], phi[2]}, V3<double>{phi_dot[0], phi_dot[1], phi_dot[2]}); std::memcpy(J9, J.m, 72);
}
int orc_project_point(int model, const double* intr, const double* p3, double* out2) { return ProjectPoint<double>(model, intr, V3<double>{p3[0], p3[1], p3[2]}, out2) ? 1 : 0; }
int orc_imu_project(int model, const double* intr, const double* w3, double* out3) {
  V3<double> o; const bool ok = ImuProject<double>(model, intr, V3<double>{w3[0], w3[1], w3[2]}, &o); out3[0] = o.x; out3[1] = o.y; out3[2] = o.z; return ok;
}
void orc_quaternion_plus(const double* q, const double* d, double* out) { QuaternionPlus(q, d, out); }
void orc_angle_axis_to_quaternion(const double* aa, double* q_xyzw) { const Qt<double> q = AngleAxisToQuaternion(V3<double>{aa[0], aa[1], aa[2]}); q_xyzw[0] = q.x; q_xyzw[1] = q.y; q_xyzw[2] = q.z; q_xyzw[3] = q.w; }
void orc_loss(int type, double a, double s, double* rho3) { EvaluateLoss(type, a, s, rho3); }
// LevenbergMarquardtStrategy::StepAccepted radius rule, exposed to pin against the stored Ceres log.
double orc_radius_after_accept(double radius, double ratio) { return std::min(1e16, radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * ratio - 1.0, 3))); }
int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// ---- Synthetic measurement generators (Camera::Project camera.cpp:155-208, Gyroscope::Project
//      gyroscope.cpp:56-82, Accelerometer::Project accelerometer.cpp:76-123), at the sensor's CURRENT state. ----
// Camera: for each time and each rigid-body point with z > 0, emits (stamp+latency, image_id, model_id, feature_id, pixel).
int orc_project_camera(void* hp, int sensor, int n, const double* times, int cap, double* stamp, int* image_id, int* model_id,
                       int* feature_id, double* pixel_xy, int* n_out) {
  Handle* h = static_cast<Handle*>(hp);
  Problem& p = h->p; p.PrepareSpline();
  const Sensor& s = p.sensors[sensor];
  int cnt = 0;
  for (int i = 0; i < n; ++i) {
    if (times[i] < p.valid_knots.front() || times[i] > p.valid_knots.back()) return fail(h, kInvalidArgument, "Cannot interpolate. Value is not within valid knots.");
    SegmentParams sp; p.GetEvaluationParams(times[i], &sp);
    double pose[6];
    SplineEvaluate<double>(&p.ctrl[size_t(sp.spline_index) * 6], p.k, sp.knot0, sp.knot1, sp.basis, times[i], 0, pose);
    // Trajectory::VectorToPose3 (trajectory.h:93-101) then T_camera_world = (T_world_rig * T_rig_cam)^-1 (typedefs.h:97-127).
    const Qt<double> q_wr = AngleAxisToQuaternion(V3<double>{pose[0], pose[1], pose[2]});
    const V3<double> t_wr{pose[3], pose[4], pose[5]};
    const Qt<double> q_rc{s.q[0], s.q[1], s.q[2], s.q[3]};
    const V3<double> t_rc{s.t[0], s.t[1], s.t[2]};
    const Qt<double> q_wc = q_wr * q_rc;
    const V3<double> t_wc = q_rotate(q_wr, t_rc) + t_wr;
    const Qt<double> q_cw{-q_wc.x, -q_wc.y, -q_wc.z, q_wc.w};
    const V3<double> t_cw = -q_rotate(q_cw, t_wc);
    for (const auto& rb : p.bodies) {
      const Qt<double> q_wm{rb.q[0], rb.q[1], rb.q[2], rb.q[3]};
      const V3<double> t_wm{rb.t[0], rb.t[1], rb.t[2]};
      const Qt<double> q_cm = q_cw * q_wm;
      const V3<double> t_cm = q_rotate(q_cw, t_wm) + t_cw;
      for (size_t f = 0; f < rb.feature_ids.size(); ++f) {
        const V3<double> pt{rb.pts[3 * f], rb.pts[3 * f + 1], rb.pts[3 * f + 2]};
        const V3<double> pc = q_rotate(q_cm, pt) + t_cm;
        if (pc.z <= 0) continue;
        double px[2] = {0, 0};
        ProjectPoint<double>(s.model, s.intr.data(), pc, px);  // status unchecked in the reference (camera.cpp:196)
        if (cnt < cap) { stamp[cnt] = times[i] + s.latency; image_id[cnt] = i; model_id[cnt] = rb.id; feature_id[cnt] = rb.feature_ids[f]; pixel_xy[2 * cnt] = px[0]; pixel_xy[2 * cnt + 1] = px[1]; }
        ++cnt;
      }
    }
  }
  *n_out = cnt;
  return kOk;
}

int orc_project_imu(void* hp, int sensor, int n, const double* times, double* stamp, double* xyz) {
  Handle* h = static_cast<Handle*>(hp);
  Problem& p = h->p; p.PrepareSpline();
  const Sensor& s = p.sensors[sensor];
  for (int i = 0; i < n; ++i) {
    if (times[i] < p.valid_knots.front() || times[i] > p.valid_knots.back()) return fail(h, kInvalidArgument, "Cannot interpolate. Value is not within valid knots.");
    SegmentParams sp; p.GetEvaluationParams(times[i], &sp);
    double d0[6], d1[6], d2[6];
    const double* C = &p.ctrl[size_t(sp.spline_index) * 6];
    SplineEvaluate<double>(C, p.k, sp.knot0, sp.knot1, sp.basis, times[i], 0, d0);
    SplineEvaluate<double>(C, p.k, sp.knot0, sp.knot1, sp.basis, times[i], 1, d1);
    SplineEvaluate<double>(C, p.k, sp.knot0, sp.knot1, sp.basis, times[i], 2, d2);
    const V3<double> phi{-d0[0], -d0[1], -d0[2]}, phid{-d1[0], -d1[1], -d1[2]}, phidd{-d2[0], -d2[1], -d2[2]};
    const Qt<double> q_rs{s.q[0], s.q[1], s.q[2], s.q[3]};
    const M3<double> J = ExpSO3Jacobian(phi);
    const V3<double> omega = J * phid;
    V3<double> out;
    if (s.type == kGyroscope) {
      const V3<double> og = -q_rotate(q_inverse(q_rs), omega);
      if (!ImuProject<double>(s.model, s.intr.data(), og, &out)) return fail(h, kInvalidArgument, "Project failed.");
    } else {
      const Qt<double> q_rw = AngleAxisToQuaternion(phi);
      const M3<double> Jdot = ExpSO3JacobianDot(phi, phid);
      const V3<double> alpha = Jdot * phid + J * phidd;
      const M3<double> Alpha = -Skew(alpha), Omega = -Skew(omega);
      const V3<double> ddt{d2[3], d2[4], d2[5]};
      const V3<double> g{p.gravity[0], p.gravity[1], p.gravity[2]};
      const V3<double> t_rs{s.t[0], s.t[1], s.t[2]};
      const V3<double> a = q_rotate(q_inverse(q_rs), q_rotate(q_rw, ddt - g) + (Omega * Omega + Alpha) * t_rs);
      if (!ImuProject<double>(s.model, s.intr.data(), a, &out)) return fail(h, kInvalidArgument, "Project failed.");
    }
    stamp[i] = times[i] + s.latency;
    xyz[3 * i] = out.x; xyz[3 * i + 1] = out.y; xyz[3 * i + 2] = out.z;
  }
  return kOk;
}

}  // extern "C"
