// calico_b200.hpp — C++ host side above the C ABI: the reference's own object model for the hot path, same class and
// method names, argument meaning and error behaviour, with every numerical step behind include/calico_b200.h (CUDA).
//
//   calico::BatchOptimizer / DefaultSolverOptions     calico/batch_optimizer.h:16-72, batch_optimizer.cpp:10-81
//   calico::sensors::Sensor (plugin surface)          calico/sensors/sensor_base.h:22-102
//   calico::sensors::Camera / Gyroscope / Accelerometer   calico/sensors/camera.{h,cpp}, gyroscope.{h,cpp}, accelerometer.{h,cpp}
//   calico::WorldModel / RigidBody / Landmark         calico/world_model.{h,cpp}
//   calico::Trajectory                                calico/trajectory.{h,cpp}
//   calico::Pose3d                                    calico/typedefs.h:39-153 (quaternion stored x,y,z,w)
// The reference's signatures use Eigen, Abseil and Ceres types; none of them exists in this image, so this header carries
// layout-independent stand-ins with the same member names (Status / StatusOr, std::array / std::vector for Eigen vectors,
// SolverOptions / Summary with the ceres::Solver field names users touch, calico.cpp:352-394). Header-only; link with
// -lcalico_b200. There is no CPU fallback: without a CUDA device every Optimize / Project / FitSpline returns kInternal.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <map>
#include <memory>
#include <optional>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>

#include "calico_b200.h"

namespace calico {

// ---- absl::Status / absl::StatusOr stand-ins (code values are absl's) ----
enum class StatusCode : int { kOk = 0, kInvalidArgument = 3, kFailedPrecondition = 9, kUnimplemented = 12, kInternal = 13 };
class Status {
 public:
  Status() = default;
  Status(StatusCode code, std::string msg) : code_(code), msg_(std::move(msg)) {}
  bool ok() const { return code_ == StatusCode::kOk; }
  StatusCode code() const { return code_; }
  const std::string& message() const { return msg_; }
 private:
  StatusCode code_ = StatusCode::kOk;
  std::string msg_;
};
inline Status OkStatus() { return Status(); }
inline Status InvalidArgumentError(std::string m) { return Status(StatusCode::kInvalidArgument, std::move(m)); }
inline Status FailedPreconditionError(std::string m) { return Status(StatusCode::kFailedPrecondition, std::move(m)); }
inline Status InternalError(std::string m) { return Status(StatusCode::kInternal, std::move(m)); }
template <class T>
class StatusOr {
 public:
  StatusOr(Status s) : status_(std::move(s)) {}          // NOLINT: implicit like absl
  StatusOr(T v) : value_(std::move(v)) {}                // NOLINT
  bool ok() const { return status_.ok(); }
  const Status& status() const { return status_; }
  T& value() { return *value_; }
  const T& value() const { return *value_; }
  T& operator*() { return *value_; }
  const T& operator*() const { return *value_; }
  T* operator->() { return &*value_; }
 private:
  Status status_;
  std::optional<T> value_;
};

using Vector2d = std::array<double, 2>;
using Vector3d = std::array<double, 3>;
using VectorXd = std::vector<double>;

// typedefs.h:39-153. rotation() holds Eigen's coeffs() order x, y, z, w.
struct Pose3d {
  std::array<double, 4> q{0.0, 0.0, 0.0, 1.0};
  Vector3d t{0.0, 0.0, 0.0};
  std::array<double, 4>& rotation() { return q; }
  const std::array<double, 4>& rotation() const { return q; }
  Vector3d& translation() { return t; }
  const Vector3d& translation() const { return t; }
};

namespace utils {
enum class LossFunctionType : int { kNone = 0, kHuber = 1, kCauchy = 2 };   // optimization_utils.h:15-22
}

namespace detail {
inline Status FromCode(int rc, const char* msg) { return rc == CB2_OK ? OkStatus() : Status(static_cast<StatusCode>(rc), msg ? msg : ""); }
}

// ---- world model (world_model.h:21-115) ----
struct Landmark { Vector3d point{0, 0, 0}; int id = 0; bool point_is_constant = true; };
struct RigidBody {
  std::unordered_map<int, Vector3d> model_definition;
  Pose3d T_world_rigidbody;
  int id = 0;
  bool world_pose_is_constant = true;
  bool model_definition_is_constant = true;
};
class WorldModel {
 public:
  ~WorldModel() {
    for (auto& [id, p] : landmarks_) if (!own_landmark_[id]) p.release();
    for (auto& [id, p] : rigidbodies_) if (!own_rigidbody_[id]) p.release();
  }
  Status AddLandmark(Landmark* landmark, bool take_ownership = true) {                      // world_model.cpp:18-27
    if (landmarks_.count(landmark->id)) return InvalidArgumentError("Landmark " + std::to_string(landmark->id) + " already exists in world model.");
    landmarks_[landmark->id].reset(landmark); own_landmark_[landmark->id] = take_ownership;
    return OkStatus();
  }
  Status AddRigidBody(RigidBody* rigidbody, bool take_ownership = true) {                   // world_model.cpp:29-38
    if (rigidbodies_.count(rigidbody->id)) return InvalidArgumentError("Rigid body " + std::to_string(rigidbody->id) + " already exists in world model.");
    rigidbodies_[rigidbody->id].reset(rigidbody); own_rigidbody_[rigidbody->id] = take_ownership;
    return OkStatus();
  }
  void SetGravity(const Vector3d& g) { gravity_ = g; }
  const Vector3d& GetGravity() const { return gravity_; }
  const Vector3d& gravity() const { return gravity_; }
  void EnableGravityEstimation(bool) {}                                                      // a no-op in the reference too (world_model.cpp:79-81)
  const std::map<int, std::unique_ptr<Landmark>>& landmarks() const { return landmarks_; }
  const std::map<int, std::unique_ptr<RigidBody>>& rigidbodies() const { return rigidbodies_; }
 private:
  std::map<int, std::unique_ptr<Landmark>> landmarks_;
  std::map<int, std::unique_ptr<RigidBody>> rigidbodies_;
  std::map<int, bool> own_landmark_, own_rigidbody_;
  Vector3d gravity_{0.0, 0.0, -9.80665};                                                     // world_model.h:78
};

// ---- trajectory (trajectory.h:27-120) ----
class Trajectory {
 public:
  static constexpr int kSplineOrder = 6;                                                     // trajectory.h:28
  static constexpr double kKnotFrequency = 10.0;
  // trajectory.cpp:14-49, on the device (cb2_fit_trajectory).
  Status FitSpline(const std::map<double, Pose3d>& poses_world_body, double knot_frequency = kKnotFrequency, int spline_order = kSplineOrder) {
    poses_ = poses_world_body;
    std::vector<double> stamps, q, t;
    for (const auto& [stamp, pose] : poses_world_body) {
      stamps.push_back(stamp);
      q.insert(q.end(), pose.q.begin(), pose.q.end());
      t.insert(t.end(), pose.t.begin(), pose.t.end());
    }
    int nk = 0, ncp = 0;
    int rc = cb2_fit_spline_size(int(stamps.size()), stamps.data(), spline_order, knot_frequency, &nk, &ncp);
    if (rc != CB2_OK) return detail::FromCode(rc, cb2_fit_last_error());
    knots_.assign(nk, 0.0); ctrl_.assign(size_t(ncp) * 6, 0.0);
    rc = cb2_fit_trajectory(int(stamps.size()), stamps.data(), q.data(), t.data(), spline_order, knot_frequency, nk, knots_.data(), ncp, ctrl_.data());
    spline_order_ = spline_order;
    return detail::FromCode(rc, cb2_fit_last_error());
  }
  const std::map<double, Pose3d>& trajectory() const { return poses_; }
  int spline_order() const { return spline_order_; }
  const std::vector<double>& knots() const { return knots_; }
  std::vector<double>& control_points() { return ctrl_; }                                   // [n_cp][6] = [axis-angle ; translation], mutated by Optimize
  const std::vector<double>& control_points() const { return ctrl_; }
 private:
  std::map<double, Pose3d> poses_;
  std::vector<double> knots_, ctrl_;
  int spline_order_ = kSplineOrder;
};

namespace sensors {

enum class CameraIntrinsicsModel : int { kNone, kOpenCv5, kOpenCv8, kKannalaBrandt, kDoubleSphere, kFieldOfView, kUnifiedCamera, kExtendedUnifiedCamera };   // camera_models.h:16-33
enum class GyroscopeIntrinsicsModel : int { kNone, kGyroscopeScaleOnly, kGyroscopeScaleAndBias, kGyroscopeVectorNav };                                       // gyroscope_models.h:16-25
enum class AccelerometerIntrinsicsModel : int { kNone, kAccelerometerScaleOnly, kAccelerometerScaleAndBias, kAccelerometerVectorNav };                       // accelerometer_models.h:16-25
// camera_models.h kNumberOfParameters (:79,231,395,596,716,848,961) / {gyroscope,accelerometer}_models.h: defined once, behind the C ABI.
inline int NumberOfParameters(CameraIntrinsicsModel m) { return cb2_num_intrinsics(CB2_CAMERA, int(m)); }
inline int NumberOfImuParameters(int m) { return cb2_num_intrinsics(CB2_GYROSCOPE, m); }

struct CameraObservationId {                                                                 // camera.h:24-50
  double stamp = 0; int image_id = 0, model_id = 0, feature_id = 0;
  bool operator==(const CameraObservationId& o) const { return stamp == o.stamp && image_id == o.image_id && model_id == o.model_id && feature_id == o.feature_id; }
};
struct CameraObservationIdHash {
  size_t operator()(const CameraObservationId& i) const {
    size_t h = std::hash<double>()(i.stamp);
    for (int v : {i.image_id, i.model_id, i.feature_id}) h ^= std::hash<int>()(v) + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
    return h;
  }
};
struct CameraMeasurement { Vector2d pixel{0, 0}; CameraObservationId id; };
struct ImuObservationId {
  double stamp = 0; int sequence = 0;
  bool operator==(const ImuObservationId& o) const { return stamp == o.stamp && sequence == o.sequence; }
};
struct ImuObservationIdHash { size_t operator()(const ImuObservationId& i) const { return std::hash<double>()(i.stamp) ^ (std::hash<int>()(i.sequence) << 1); } };
using GyroscopeObservationId = ImuObservationId;
using AccelerometerObservationId = ImuObservationId;
struct GyroscopeMeasurement { Vector3d measurement{0, 0, 0}; GyroscopeObservationId id; };
struct AccelerometerMeasurement { Vector3d measurement{0, 0, 0}; AccelerometerObservationId id; };

// The plugin surface, sensor_base.h:22-102. The four Ceres-facing virtuals (AddParametersToProblem, AddResidualsToProblem,
// UpdateResiduals on a ceres::Problem) become two that talk to the C ABI handle instead.
class Sensor {
 public:
  virtual ~Sensor() = default;
  virtual void SetName(const std::string& name) = 0;
  virtual const std::string& GetName() const = 0;
  virtual void SetExtrinsics(const Pose3d& T_sensorrig_sensor) = 0;
  virtual const Pose3d& GetExtrinsics() const = 0;
  virtual Status SetIntrinsics(const VectorXd& intrinsics) = 0;
  virtual const VectorXd& GetIntrinsics() const = 0;
  virtual Status SetLatency(double latency) = 0;
  virtual double GetLatency() const = 0;
  virtual void EnableExtrinsicsEstimation(bool enable) = 0;
  virtual void EnableIntrinsicsEstimation(bool enable) = 0;
  virtual void EnableLatencyEstimation(bool enable) = 0;
  virtual void ClearResidualInfo() = 0;
  virtual void SetLossFunction(utils::LossFunctionType loss, double scale = 1.0) = 0;
  virtual Status SetMeasurementNoise(double sigma) = 0;
  // AddParametersToProblem + AddResidualsToProblem (camera.cpp:92-153): hand the sensor and its measurements to the handle.
  virtual StatusOr<int> AddToProblem(cb2_problem* problem) = 0;
  // The in-place parameter mutation Ceres does through raw pointers (camera.cpp:98-101) + UpdateResiduals (camera.cpp:70-80).
  virtual Status ReadBack(cb2_problem* problem, int sensor_id) = 0;
};

namespace detail {
// State and behaviour shared by the three sensors (the reference repeats it per class).
template <class Derived, class Measurement, class Id, class IdHash, int kKind, int kDim>
class SensorImpl : public Sensor {
 public:
  void SetName(const std::string& name) final { name_ = name; }
  const std::string& GetName() const final { return name_; }
  void SetExtrinsics(const Pose3d& T) final { T_sensorrig_sensor_ = T; }
  const Pose3d& GetExtrinsics() const final { return T_sensorrig_sensor_; }
  Status SetIntrinsics(const VectorXd& intrinsics) final {                                  // camera.cpp:22-36
    if (model_ <= 0) return InvalidArgumentError(std::string(Derived::kTypeName) + " model has not been set!");
    if (int(intrinsics.size()) != num_params_)
      return InvalidArgumentError("Tried to set intrinsics of size " + std::to_string(intrinsics.size()) + " for " + Derived::kLowerName + " " + name_ +
                                  ". Expected intrinsics size of " + std::to_string(num_params_));
    intrinsics_ = intrinsics;
    return OkStatus();
  }
  const VectorXd& GetIntrinsics() const final { return intrinsics_; }
  Status SetLatency(double latency) final { latency_ = latency; return OkStatus(); }
  double GetLatency() const final { return latency_; }
  void EnableExtrinsicsEstimation(bool e) final { extrinsics_enabled_ = e; }
  void EnableIntrinsicsEstimation(bool e) final { intrinsics_enabled_ = e; }
  void EnableLatencyEstimation(bool e) final { latency_enabled_ = e; }
  void SetLossFunction(utils::LossFunctionType loss, double scale = 1.0) final { loss_ = loss; loss_scale_ = scale; }
  Status SetMeasurementNoise(double sigma) final {                                          // camera.cpp:62-68
    if (sigma <= 0.0) return InvalidArgumentError("Sigma must be greater than 0.");
    sigma_ = sigma;
    return OkStatus();
  }
  void ClearResidualInfo() final { id_to_residual_.clear(); }
  Status AddMeasurement(const Measurement& m) {                                             // camera.cpp:224-236
    if (index_.count(m.id)) return InvalidArgumentError(Derived::RedundantMessage(m.id));
    index_[m.id] = int(measurements_.size());
    measurements_.push_back(m);
    return OkStatus();
  }
  Status AddMeasurements(const std::vector<Measurement>& ms) {                              // camera.cpp:238-252
    std::string message;
    for (const auto& m : ms) { const Status s = AddMeasurement(m); if (!s.ok()) message += s.message() + "\n"; }
    return message.empty() ? OkStatus() : InvalidArgumentError(message);
  }
  void ClearMeasurements() { measurements_.clear(); index_.clear(); id_to_residual_.clear(); outlier_ids_.clear(); }
  int NumberOfMeasurements() const { return int(measurements_.size()); }
  const std::vector<Measurement>& measurements() const { return measurements_; }
  // camera.cpp:258-279.
  StatusOr<std::vector<std::pair<Measurement, std::array<double, kDim>>>> GetMeasurementResidualPairs() const {
    if (id_to_residual_.size() > measurements_.size()) return InternalError("There are more residuals than measurements.");
    if (measurements_.empty()) return FailedPreconditionError("Measurements are empty. Nothing to return.");
    std::vector<std::pair<Measurement, std::array<double, kDim>>> pairs;
    for (const auto& [id, r] : id_to_residual_) {
      auto it = index_.find(id);
      if (it == index_.end()) return InternalError("Found a residual that doesn't correspond to any measurement.");
      pairs.push_back({measurements_[it->second], r});
    }
    return pairs;
  }

  Status ReadBack(cb2_problem* problem, int sid) final {
    Pose3d T;
    double latency = 0;
    VectorXd intr(intrinsics_.size());
    int rc = cb2_get_sensor(problem, sid, intr.data(), T.q.data(), T.t.data(), &latency);
    if (rc != CB2_OK) return calico::detail::FromCode(rc, cb2_last_error(problem));
    intrinsics_ = intr; T_sensorrig_sensor_ = T; latency_ = latency;
    std::vector<double> r(measurements_.size() * kDim);
    std::vector<uint8_t> valid(measurements_.size());
    rc = cb2_get_residuals(problem, sid, r.data(), valid.data());
    if (rc != CB2_OK) return calico::detail::FromCode(rc, cb2_last_error(problem));
    id_to_residual_.clear();
    for (size_t i = 0; i < measurements_.size(); ++i) {
      if (!valid[i]) continue;                                                               // outliers and failed blocks have no residual
      std::array<double, kDim> v;
      for (int q = 0; q < kDim; ++q) v[q] = r[i * kDim + q];
      id_to_residual_[measurements_[i].id] = v;
    }
    return OkStatus();
  }

 protected:
  StatusOr<int> AddSensorToProblem(cb2_problem* problem) const {
    if (model_ <= 0) return FailedPreconditionError("Cannot add sensor parameters. Model is not yet defined.");   // camera.cpp:95
    int sid = -1;
    const int rc = cb2_add_sensor(problem, kKind, model_, name_.c_str(), int(intrinsics_.size()), intrinsics_.data(), T_sensorrig_sensor_.q.data(),
                                  T_sensorrig_sensor_.t.data(), latency_, sigma_, int(loss_), loss_scale_, intrinsics_enabled_, extrinsics_enabled_,
                                  latency_enabled_, &sid);
    if (rc != CB2_OK) return calico::detail::FromCode(rc, cb2_last_error(problem));
    return sid;
  }
  Status SetModelImpl(int model, int num_params, const char* what) {
    if (num_params <= 0) return InvalidArgumentError(std::string("Could not create ") + what + " model for type " + std::to_string(model) + ". It is likely not yet implemented.");
    model_ = model; num_params_ = num_params;
    intrinsics_.assign(num_params, 0.0);                                                     // camera.cpp:212
    return OkStatus();
  }
  std::string name_;
  Pose3d T_sensorrig_sensor_;
  VectorXd intrinsics_;
  double latency_ = 0.0, sigma_ = 1.0, loss_scale_ = 1.0;
  utils::LossFunctionType loss_ = utils::LossFunctionType::kNone;
  bool extrinsics_enabled_ = false, intrinsics_enabled_ = false, latency_enabled_ = false;
  int model_ = 0, num_params_ = 0;
  std::vector<Measurement> measurements_;
  std::unordered_map<Id, int, IdHash> index_;
  std::unordered_map<Id, std::array<double, kDim>, IdHash> id_to_residual_;
  std::unordered_set<Id, IdHash> outlier_ids_;
};

// *::Project at the sensor's current state (camera.cpp:155-208, gyroscope.cpp:56-82, accelerometer.cpp:76-123) through the device:
// the residual of a zero measurement with unit sigma and zero latency is minus the projection.
constexpr int kLandmarkFrameId = -1;   // camera.h: model_id of landmark observations
// landmarks_as_body (Camera::Project only, camera.cpp:169-184): the landmarks ride along as a constant pseudo rigid body with identity pose and
// id kLandmarkFrameId. Optimize never pushes them: the reference rejects landmark observations (camera.cpp:125-131) and so does cb2_upload.
inline Status PushWorldAndTrajectory(cb2_problem* p, const Trajectory& trajectory, const WorldModel& world_model, bool landmarks_as_body = false) {
  int rc = cb2_set_trajectory(p, trajectory.spline_order(), int(trajectory.knots().size()), trajectory.knots().data(),
                              int(trajectory.control_points().size() / 6), trajectory.control_points().data());
  if (rc != CB2_OK) return calico::detail::FromCode(rc, cb2_last_error(p));
  cb2_set_gravity(p, world_model.gravity().data());
  for (const auto& [id, body] : world_model.rigidbodies()) {
    std::vector<int> ids; std::vector<double> pts;
    for (const auto& [pid, pt] : body->model_definition) { ids.push_back(pid); pts.insert(pts.end(), pt.begin(), pt.end()); }
    rc = cb2_add_rigid_body(p, id, body->T_world_rigidbody.q.data(), body->T_world_rigidbody.t.data(), int(ids.size()), ids.data(), pts.data(),
                            body->world_pose_is_constant, body->model_definition_is_constant);
    if (rc != CB2_OK) return calico::detail::FromCode(rc, cb2_last_error(p));
  }
  if (landmarks_as_body && !world_model.landmarks().empty()) {
    if (world_model.rigidbodies().count(kLandmarkFrameId)) return InvalidArgumentError("Rigid body id -1 is reserved for landmarks (kLandmarkFrameId).");
    std::vector<int> ids; std::vector<double> pts;
    for (const auto& [lid, lm] : world_model.landmarks()) { ids.push_back(lid); pts.insert(pts.end(), lm->point.begin(), lm->point.end()); }
    const double q[4] = {0, 0, 0, 1}, t[3] = {0, 0, 0};
    rc = cb2_add_rigid_body(p, kLandmarkFrameId, q, t, int(ids.size()), ids.data(), pts.data(), 1, 1);
    if (rc != CB2_OK) return calico::detail::FromCode(rc, cb2_last_error(p));
  }
  return OkStatus();
}
struct Handle {
  cb2_problem* p = nullptr;
  Handle() { cb2_problem_create(&p); }
  ~Handle() { cb2_problem_destroy(p); }
};
}  // namespace detail

class Camera final : public detail::SensorImpl<Camera, CameraMeasurement, CameraObservationId, CameraObservationIdHash, 0, 2> {
 public:
  static constexpr const char* kTypeName = "Camera";
  static constexpr const char* kLowerName = "camera";
  static std::string RedundantMessage(const CameraObservationId& id) {
    return "Tried to add redundant measurement - Image id: " + std::to_string(id.image_id) + ", model id: " + std::to_string(id.model_id) +
           ", feature id: " + std::to_string(id.feature_id);
  }
  Status SetModel(CameraIntrinsicsModel m) { return SetModelImpl(int(m), m == CameraIntrinsicsModel::kNone ? -1 : NumberOfParameters(m), "camera"); }
  CameraIntrinsicsModel GetModel() const { return CameraIntrinsicsModel(model_); }
  Status MarkOutlierById(const CameraObservationId& id) {                                   // camera.cpp:281-291
    if (!index_.count(id)) return InvalidArgumentError("Attempted to add id that is not within the measurement set.");
    outlier_ids_.insert(id);
    return OkStatus();
  }
  Status MarkOutliersById(const std::vector<CameraObservationId>& ids) { for (const auto& id : ids) { Status s = MarkOutlierById(id); if (!s.ok()) return s; } return OkStatus(); }
  void ClearOutliersList() { outlier_ids_.clear(); }
  StatusOr<int> AddToProblem(cb2_problem* problem) override {
    StatusOr<int> sid = AddSensorToProblem(problem);
    if (!sid.ok()) return sid;
    const size_t n = measurements_.size();
    if (n == 0) return sid;
    std::vector<double> stamp(n), pixel(2 * n);
    std::vector<int> image_id(n), model_id(n), feature_id(n);
    std::vector<uint8_t> outlier(n, 0);
    for (size_t i = 0; i < n; ++i) {
      const CameraMeasurement& m = measurements_[i];
      stamp[i] = m.id.stamp; image_id[i] = m.id.image_id; model_id[i] = m.id.model_id; feature_id[i] = m.id.feature_id;
      pixel[2 * i] = m.pixel[0]; pixel[2 * i + 1] = m.pixel[1];
      outlier[i] = outlier_ids_.count(m.id) ? 1 : 0;                                         // camera.cpp:121-124
    }
    const int rc = cb2_add_camera_observations(problem, *sid, int(n), stamp.data(), image_id.data(), model_id.data(), feature_id.data(), pixel.data(), outlier.data());
    if (rc != CB2_OK) return calico::detail::FromCode(rc, cb2_last_error(problem));
    return sid;
  }
  // camera.cpp:155-208: per time all landmarks (model_id = kLandmarkFrameId), then every rigid body's points; points with z <= 0 are skipped.
  StatusOr<std::vector<CameraMeasurement>> Project(const std::vector<double>& interp_times, const Trajectory& trajectory, const WorldModel& world_model) const {
    detail::Handle h;
    Status s = detail::PushWorldAndTrajectory(h.p, trajectory, world_model, /*landmarks_as_body=*/true);
    if (!s.ok()) return s;
    if (model_ <= 0) return FailedPreconditionError("Camera model has not been set!");
    int sid = -1;
    int rc = cb2_add_sensor(h.p, 0, model_, name_.c_str(), int(intrinsics_.size()), intrinsics_.data(), T_sensorrig_sensor_.q.data(), T_sensorrig_sensor_.t.data(),
                            0.0, 1.0, 0, 1.0, 0, 0, 0, &sid);
    if (rc != CB2_OK) return calico::detail::FromCode(rc, cb2_last_error(h.p));
    std::vector<double> stamp, pixel;
    std::vector<int> image_id, model_id, feature_id;
    for (size_t i = 0; i < interp_times.size(); ++i) {
      for (const auto& [lid, lm] : world_model.landmarks()) {
        (void)lm;
        stamp.push_back(interp_times[i]); image_id.push_back(int(i)); model_id.push_back(detail::kLandmarkFrameId); feature_id.push_back(lid);
      }
      for (const auto& [rid, body] : world_model.rigidbodies())
        for (const auto& [pid, pt] : body->model_definition) {
          (void)pt;
          stamp.push_back(interp_times[i]); image_id.push_back(int(i)); model_id.push_back(rid); feature_id.push_back(pid);
        }
    }
    pixel.assign(2 * stamp.size(), 0.0);
    const int n = int(stamp.size());
    std::vector<CameraMeasurement> out;
    if (n == 0) return out;
    rc = cb2_add_camera_observations(h.p, sid, n, stamp.data(), image_id.data(), model_id.data(), feature_id.data(), pixel.data(), nullptr);
    if (rc != CB2_OK) return calico::detail::FromCode(rc, cb2_last_error(h.p));
    std::vector<double> r(2 * size_t(n));
    std::vector<uint8_t> valid(n);
    rc = cb2_evaluate_sensor(h.p, sid, r.data(), nullptr, valid.data());
    if (rc != CB2_OK) return calico::detail::FromCode(rc, cb2_last_error(h.p));
    for (int i = 0; i < n; ++i) {
      if (valid[i] != 1) continue;   // bit 1 = point_camera.z() <= 0: skipped for every model (camera.cpp:172-174,186-188); 0 = projection failed
      out.push_back(CameraMeasurement{{-r[2 * i], -r[2 * i + 1]}, {stamp[i] + latency_, image_id[i], model_id[i], feature_id[i]}});
    }
    return out;
  }
};

namespace detail {
template <class Derived, class Measurement, class ModelEnum, int kKind>
class ImuImpl : public SensorImpl<Derived, Measurement, ImuObservationId, ImuObservationIdHash, kKind, 3> {
  using Base = SensorImpl<Derived, Measurement, ImuObservationId, ImuObservationIdHash, kKind, 3>;
 public:
  static std::string RedundantMessage(const ImuObservationId& id) { return "Tried to add redundant measurement - stamp: " + std::to_string(id.stamp) + ", sequence: " + std::to_string(id.sequence); }
  Status SetModel(ModelEnum m) { return this->SetModelImpl(int(m), NumberOfImuParameters(int(m)), Derived::kLowerName); }
  ModelEnum GetModel() const { return ModelEnum(this->model_); }
  StatusOr<int> AddToProblem(cb2_problem* problem) override {
    StatusOr<int> sid = this->AddSensorToProblem(problem);
    if (!sid.ok()) return sid;
    const size_t n = this->measurements_.size();
    if (n == 0) return sid;
    std::vector<double> stamp(n), xyz(3 * n);
    std::vector<int> seq(n);
    for (size_t i = 0; i < n; ++i) {
      const Measurement& m = this->measurements_[i];
      stamp[i] = m.id.stamp; seq[i] = m.id.sequence;
      for (int q = 0; q < 3; ++q) xyz[3 * i + q] = m.measurement[q];
    }
    const int rc = cb2_add_imu_observations(problem, *sid, int(n), stamp.data(), seq.data(), xyz.data());
    if (rc != CB2_OK) return calico::detail::FromCode(rc, cb2_last_error(problem));
    return sid;
  }
  // gyroscope.cpp:56-82 / accelerometer.cpp:76-123.
  StatusOr<std::vector<Measurement>> Project(const std::vector<double>& interp_times, const Trajectory& trajectory, const WorldModel& world_model) const {
    Handle h;
    Status s = PushWorldAndTrajectory(h.p, trajectory, world_model);
    if (!s.ok()) return s;
    if (this->model_ <= 0) return FailedPreconditionError(std::string(Derived::kTypeName) + " model has not been set!");
    int sid = -1;
    int rc = cb2_add_sensor(h.p, kKind, this->model_, this->name_.c_str(), int(this->intrinsics_.size()), this->intrinsics_.data(), this->T_sensorrig_sensor_.q.data(),
                            this->T_sensorrig_sensor_.t.data(), 0.0, 1.0, 0, 1.0, 0, 0, 0, &sid);
    if (rc != CB2_OK) return calico::detail::FromCode(rc, cb2_last_error(h.p));
    const int n = int(interp_times.size());
    std::vector<Measurement> out;
    if (n == 0) return out;
    std::vector<int> seq(n);
    for (int i = 0; i < n; ++i) seq[i] = i;
    std::vector<double> zeros(3 * size_t(n), 0.0), r(3 * size_t(n));
    std::vector<uint8_t> valid(n);
    rc = cb2_add_imu_observations(h.p, sid, n, interp_times.data(), seq.data(), zeros.data());
    if (rc != CB2_OK) return calico::detail::FromCode(rc, cb2_last_error(h.p));
    rc = cb2_evaluate_sensor(h.p, sid, r.data(), nullptr, valid.data());
    if (rc != CB2_OK) return calico::detail::FromCode(rc, cb2_last_error(h.p));
    for (int i = 0; i < n; ++i) {
      if (!valid[i]) return InternalError(std::string("Failed to project ") + Derived::kLowerName + " measurement.");
      out.push_back(Measurement{{-r[3 * i], -r[3 * i + 1], -r[3 * i + 2]}, {interp_times[i] + this->latency_, i}});
    }
    return out;
  }
};
}  // namespace detail

class Gyroscope final : public detail::ImuImpl<Gyroscope, GyroscopeMeasurement, GyroscopeIntrinsicsModel, 1> {
 public:
  static constexpr const char* kTypeName = "Gyroscope";
  static constexpr const char* kLowerName = "gyroscope";
};
class Accelerometer final : public detail::ImuImpl<Accelerometer, AccelerometerMeasurement, AccelerometerIntrinsicsModel, 2> {
 public:
  static constexpr const char* kTypeName = "Accelerometer";
  static constexpr const char* kLowerName = "accelerometer";
};

}  // namespace sensors

// ---- ceres::Solver::Options / Summary stand-ins: the fields Calico's users touch (calico.cpp:352-394, batch_optimizer.cpp:10-17) ----
enum TerminationType { CONVERGENCE = 0, NO_CONVERGENCE = 1, FAILURE = 2 };
struct SolverOptions {
  int max_num_iterations = 50;
  int num_threads = 1;
  double function_tolerance = 1e-8, gradient_tolerance = 1e-10, parameter_tolerance = 1e-10;
  bool minimizer_progress_to_stdout = true;
};
inline SolverOptions DefaultSolverOptions() { return SolverOptions{}; }                      // batch_optimizer.cpp:10-17
struct Summary {
  TerminationType termination_type = FAILURE;
  std::string message;
  double initial_cost = 0, final_cost = 0, total_time_in_seconds = 0;
  int num_successful_steps = 0, num_unsuccessful_steps = 0, num_iterations = 0;
  int num_residual_blocks = 0, num_residuals = 0, num_parameter_blocks = 0, num_parameters = 0;
  int num_parameter_blocks_reduced = 0, num_parameters_reduced = 0, num_effective_parameters_reduced = 0, num_residual_blocks_reduced = 0,
      num_residuals_reduced = 0;
  bool IsSolutionUsable() const { return termination_type == CONVERGENCE || termination_type == NO_CONVERGENCE; }
  std::string BriefReport() const {
    char buf[512];
    std::snprintf(buf, sizeof buf, "calico_b200 Report: Iterations: %d, Initial cost: %e, Final cost: %e, Termination: %s", num_iterations, initial_cost,
                  final_cost, termination_type == CONVERGENCE ? "CONVERGENCE" : (termination_type == NO_CONVERGENCE ? "NO_CONVERGENCE" : "FAILURE"));
    return buf;
  }
  std::string FullReport() const {
    char buf[1024];
    std::snprintf(buf, sizeof buf, "%s\nResidual blocks %d, residuals %d, parameter blocks %d (reduced %d), parameters %d (reduced %d)\nSuccessful steps %d, "
                  "unsuccessful steps %d, total time %.6f s\n%s", BriefReport().c_str(), num_residual_blocks, num_residuals, num_parameter_blocks,
                  num_parameter_blocks_reduced, num_parameters, num_parameters_reduced, num_successful_steps, num_unsuccessful_steps, total_time_in_seconds,
                  message.c_str());
    return buf;
  }
};

// ---- batch_optimizer.h:16-72 ----
class BatchOptimizer {
 public:
  ~BatchOptimizer() {                                                                        // batch_optimizer.cpp:19-33
    for (size_t i = 0; i < sensors_.size(); ++i) if (!own_sensors_[i]) sensors_[i].release();
    if (!own_world_model_) world_model_.release();
    if (!own_trajectory_world_body_) trajectory_world_body_.release();
  }
  void AddSensor(sensors::Sensor* sensor, bool take_ownership = true) { sensors_.emplace_back(sensor); own_sensors_.push_back(take_ownership); }
  void AddWorldModel(WorldModel* world_model, bool take_ownership = true) { world_model_.reset(world_model); own_world_model_ = take_ownership; }
  void AddTrajectory(Trajectory* trajectory_world_sensorrig, bool take_ownership = true) { trajectory_world_body_.reset(trajectory_world_sensorrig); own_trajectory_world_body_ = take_ownership; }

  // batch_optimizer.cpp:53-81: a new problem per call, warm-started from the objects' current values; parameters are written back
  // into the objects; residuals are refreshed. Non-convergence is not an error (the summary says so).
  StatusOr<Summary> Optimize(const SolverOptions& options = DefaultSolverOptions()) {
    if (!trajectory_world_body_ || !world_model_) return FailedPreconditionError("Trajectory and world model must be added before optimizing.");
    sensors::detail::Handle h;
    Status s = sensors::detail::PushWorldAndTrajectory(h.p, *trajectory_world_body_, *world_model_);
    if (!s.ok()) return s;
    std::vector<int> ids;
    for (auto& sensor : sensors_) {
      sensor->ClearResidualInfo();                                                           // batch_optimizer.cpp:63
      StatusOr<int> sid = sensor->AddToProblem(h.p);
      if (!sid.ok()) return sid.status();
      ids.push_back(*sid);
    }
    cb2_options o;
    cb2_default_options(&o);
    o.max_num_iterations = options.max_num_iterations; o.num_threads = options.num_threads;
    o.function_tolerance = options.function_tolerance; o.gradient_tolerance = options.gradient_tolerance;
    o.parameter_tolerance = options.parameter_tolerance; o.minimizer_progress_to_stdout = options.minimizer_progress_to_stdout ? 1 : 0;
    cb2_summary cs;
    const int rc = cb2_optimize(h.p, &o, &cs, nullptr, 0, nullptr);
    // Parameters are written back even when the residual refresh failed, as Ceres has already mutated them in the reference.
    cb2_get_trajectory(h.p, trajectory_world_body_->control_points().data());
    for (const auto& [rid, body] : world_model_->rigidbodies()) {   // freed world-model blocks are mutated in place too (world_model.cpp:52-70)
      if (body->world_pose_is_constant && body->model_definition_is_constant) continue;
      std::vector<double> pts(3 * body->model_definition.size());
      if (cb2_get_rigid_body(h.p, rid, body->T_world_rigidbody.q.data(), body->T_world_rigidbody.t.data(), pts.data()) != CB2_OK) continue;
      size_t k = 0;   // same iteration order as PushWorldAndTrajectory (the map is not modified in between)
      for (auto& [pid, pt] : body->model_definition) { (void)pid; pt = {pts[3 * k], pts[3 * k + 1], pts[3 * k + 2]}; ++k; }
    }
    for (size_t i = 0; i < sensors_.size(); ++i) {
      const Status rs = sensors_[i]->ReadBack(h.p, ids[i]);
      if (rc == CB2_OK && !rs.ok()) return rs;
    }
    if (rc != CB2_OK) return detail::FromCode(rc, cb2_last_error(h.p));
    Summary S;
    S.termination_type = TerminationType(cs.termination_type); S.message = cs.message;
    S.initial_cost = cs.initial_cost; S.final_cost = cs.final_cost; S.total_time_in_seconds = cs.total_time;
    S.num_successful_steps = cs.num_successful_steps; S.num_unsuccessful_steps = cs.num_unsuccessful_steps; S.num_iterations = cs.num_iterations;
    S.num_residual_blocks = cs.num_residual_blocks; S.num_residuals = cs.num_residuals; S.num_parameter_blocks = cs.num_parameter_blocks;
    S.num_parameters = cs.num_parameters; S.num_parameter_blocks_reduced = cs.num_parameter_blocks_reduced; S.num_parameters_reduced = cs.num_parameters_reduced;
    S.num_effective_parameters_reduced = cs.num_effective_parameters_reduced; S.num_residual_blocks_reduced = cs.num_residual_blocks_reduced;
    S.num_residuals_reduced = cs.num_residuals_reduced;
    return S;
  }

 private:
  bool own_trajectory_world_body_ = true, own_world_model_ = true;
  std::vector<bool> own_sensors_;
  std::vector<std::unique_ptr<sensors::Sensor>> sensors_;
  std::unique_ptr<WorldModel> world_model_;
  std::unique_ptr<Trajectory> trajectory_world_body_;
};

}  // namespace calico
