// The reference's integration test, calico/test/batch_optimizer_test.cpp:32-213 (ToyStereoCameraAndImuCalibration) with the
// fixture of calico/test_utils.h:11-116 (DefaultSyntheticTest), transcribed against include/calico_b200.hpp: same objects, same
// call sequence, same acceptance (CONVERGENCE, final_cost < 1e-7, every calibrated parameter within 1e-7 of the truth).
// Plus the error behaviour of the sensor setters (camera.cpp:22-36, 62-68, 224-236). Built by tests/test_cpp_mirror.py against
// libcalico_b200.so (GPU) or the SIMT-emulation build of the same sources (CPU, tiny variant).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>

#include "calico_b200.hpp"

using namespace calico;

static int g_failures = 0;
#define EXPECT_TRUE(c) do { if (!(c)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); ++g_failures; } } while (0)
#define EXPECT_OK(e) do { const Status s_ = (e); if (!s_.ok()) { std::printf("FAILED %s:%d: %s -> %s\n", __FILE__, __LINE__, #e, s_.message().c_str()); ++g_failures; } } while (0)
#define ASSERT_OK(e) do { const Status s_ = (e); if (!s_.ok()) { std::printf("FAILED %s:%d: %s -> %s\n", __FILE__, __LINE__, #e, s_.message().c_str()); return 1; } } while (0)
#define EXPECT_NEAR(a, b, tol) do { if (!(std::fabs((a) - (b)) <= (tol))) { std::printf("FAILED %s:%d: |%s - %s| = %.3e > %.1e\n", __FILE__, __LINE__, #a, #b, std::fabs((a) - (b)), double(tol)); ++g_failures; } } while (0)

using Quat = std::array<double, 4>;   // x, y, z, w
static Quat AngleAxis(double angle, const Vector3d& axis) {
  const double s = std::sin(0.5 * angle);
  return {axis[0] * s, axis[1] * s, axis[2] * s, std::cos(0.5 * angle)};
}
static Quat Mul(const Quat& a, const Quat& b) {
  return {a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1], a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2],
          a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0], a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2]};
}
static std::mt19937_64 g_rng(7);
static Vector3d RandomVector() { std::uniform_real_distribution<double> u(-1.0, 1.0); return {u(g_rng), u(g_rng), u(g_rng)}; }   // Eigen::Vector3d::Random()
static Vector3d Normalized(Vector3d v) { const double n = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); return {v[0] / n, v[1] / n, v[2] / n}; }

// test_utils.h:11-116.
struct DefaultSyntheticTest {
  std::map<double, Pose3d> trajectory_world_sensorrig;
  std::vector<double> stamps;
  std::vector<Vector3d> t_world_points;
  explicit DefaultSyntheticTest(int samples_per_segment) {
    const Quat q0 = Mul(AngleAxis(M_PI, {0, 0, 1}), AngleAxis(M_PI, {1, 0, 0}));
    const Vector3d t0{0.0, 0.0, 1.0};
    const double kAngle = 30.0 * M_PI / 180.0, kPos = 0.5, kSegmentDuration = 0.75;
    const std::vector<double> ang{0.0, kAngle, 0.0, -kAngle, 0.0}, pos{0.0, kPos, 0.0, -kPos, 0.0};
    std::vector<double> interp(samples_per_segment);
    const double dti = 1.0 / samples_per_segment, dt = dti * kSegmentDuration;
    for (int i = 0; i < samples_per_segment; ++i) interp[i] = (std::sin(dti * i * M_PI - M_PI_2) + 1.0) / 2.0;
    double now = 0.0;
    for (const Vector3d& axis : std::vector<Vector3d>{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}) {
      for (size_t i = 1; i < ang.size(); ++i)
        for (double u : interp) {
          Pose3d T; T.q = Mul(q0, AngleAxis((ang[i] - ang[i - 1]) * u + ang[i - 1], axis)); T.t = t0;
          trajectory_world_sensorrig[now] = T; now += dt;
        }
      for (size_t i = 1; i < pos.size(); ++i)
        for (double u : interp) {
          const double p = (pos[i] - pos[i - 1]) * u + pos[i - 1];
          Pose3d T; T.q = q0; T.t = {axis[0] * p + t0[0], axis[1] * p + t0[1], axis[2] * p + t0[2]};
          trajectory_world_sensorrig[now] = T; now += dt;
        }
    }
    for (const auto& [stamp, pose] : trajectory_world_sensorrig) { (void)pose; stamps.push_back(stamp); }
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) t_world_points.push_back({i * 0.3 - 0.75, j * 0.3 - 0.75, 0.0});
  }
};

static double MaxAbsDiff(const VectorXd& a, const VectorXd& b) { double m = 0; for (size_t i = 0; i < a.size(); ++i) m = std::max(m, std::fabs(a[i] - b[i])); return a.size() == b.size() ? m : 1e300; }
static double PoseDiff(const Pose3d& a, const Pose3d& b) {
  double m = 0, sgn = (a.q[0] * b.q[0] + a.q[1] * b.q[1] + a.q[2] * b.q[2] + a.q[3] * b.q[3]) < 0 ? -1.0 : 1.0;
  for (int i = 0; i < 4; ++i) m = std::max(m, std::fabs(a.q[i] - sgn * b.q[i]));
  for (int i = 0; i < 3; ++i) m = std::max(m, std::fabs(a.t[i] - b.t[i]));
  return m;
}

static int SetterErrors() {
  sensors::Camera camera;
  camera.SetName("cam");
  EXPECT_TRUE(camera.SetIntrinsics(VectorXd(8, 1.0)).code() == StatusCode::kInvalidArgument);            // model not set (camera.cpp:23-25)
  EXPECT_OK(camera.SetModel(sensors::CameraIntrinsicsModel::kOpenCv5));
  EXPECT_TRUE(camera.GetIntrinsics().size() == 8);
  EXPECT_TRUE(camera.SetIntrinsics(VectorXd(7, 1.0)).code() == StatusCode::kInvalidArgument);            // wrong size (camera.cpp:26-33)
  EXPECT_TRUE(camera.SetMeasurementNoise(0.0).code() == StatusCode::kInvalidArgument);                   // camera.cpp:62-68
  sensors::CameraMeasurement m{{1.0, 2.0}, {0.5, 0, 0, 3}};
  EXPECT_OK(camera.AddMeasurement(m));
  EXPECT_TRUE(camera.AddMeasurement(m).code() == StatusCode::kInvalidArgument);                          // redundant (camera.cpp:224-232)
  EXPECT_TRUE(camera.AddMeasurements({m, m}).code() == StatusCode::kInvalidArgument);
  EXPECT_TRUE(camera.MarkOutlierById({0.5, 0, 0, 4}).code() == StatusCode::kInvalidArgument);            // camera.cpp:281-288
  EXPECT_OK(camera.MarkOutlierById(m.id));
  EXPECT_TRUE(camera.NumberOfMeasurements() == 1);
  sensors::Gyroscope gyro;
  EXPECT_TRUE(gyro.SetIntrinsics(VectorXd(4, 1.0)).code() == StatusCode::kInvalidArgument);
  EXPECT_OK(gyro.SetModel(sensors::GyroscopeIntrinsicsModel::kGyroscopeScaleAndBias));
  EXPECT_TRUE(gyro.GetIntrinsics().size() == 4);
  WorldModel wm;
  RigidBody body;
  EXPECT_OK(wm.AddRigidBody(&body, false));
  EXPECT_TRUE(wm.AddRigidBody(&body, false).code() == StatusCode::kInvalidArgument);                     // world_model.cpp:29-38
  Trajectory traj;
  EXPECT_TRUE(traj.FitSpline({}).code() == StatusCode::kInvalidArgument);                                // bspline.hpp:304-306
  return 0;
}

int main(int argc, char** argv) {
  const int samples_per_segment = argc > 1 ? std::atoi(argv[1]) : 10;      // 10 = the reference fixture (240 poses); smaller for the emulated run
  const int max_iterations = argc > 2 ? std::atoi(argv[2]) : 50;
  if (SetterErrors()) return 1;
  DefaultSyntheticTest fixture(samples_per_segment);
  const std::vector<double>& stamps = fixture.stamps;

  // Construct a world model consisting of a single planar object.
  RigidBody planar_target;
  planar_target.world_pose_is_constant = true;
  planar_target.model_definition_is_constant = true;
  for (size_t i = 0; i < fixture.t_world_points.size(); ++i) planar_target.model_definition[int(i)] = fixture.t_world_points[i];
  WorldModel* world_model = new WorldModel;
  const Vector3d true_gravity = world_model->gravity();
  EXPECT_OK(world_model->AddRigidBody(&planar_target, /*take_ownership=*/false));
  // Construct the sensorrig trajectory.
  Trajectory* trajectory_world_sensorrig = new Trajectory;
  ASSERT_OK(trajectory_world_sensorrig->FitSpline(fixture.trajectory_world_sensorrig));

  // Construct ground truth cameras and measurements.
  const sensors::CameraIntrinsicsModel kCameraModel = sensors::CameraIntrinsicsModel::kOpenCv5;
  const double kStereoRotationAngle = 2.0 * M_PI / 180.0, kStereoBaseline = 0.05, kRightCameraLatency = 0.01;
  const VectorXd true_camera_intrinsics{785, 640, 400, -3.149e-1, 1.069e-1, 1.616e-4, 1.141e-4, -1.853e-2};
  Pose3d true_extrinsics_left, true_extrinsics_right;
  true_extrinsics_right.rotation() = AngleAxis(kStereoRotationAngle, Normalized(RandomVector()));
  { const Vector3d r = RandomVector(); true_extrinsics_right.translation() = {kStereoBaseline * r[0], kStereoBaseline * r[1], kStereoBaseline * r[2]}; }
  sensors::Camera true_camera_left;
  EXPECT_OK(true_camera_left.SetModel(kCameraModel));
  EXPECT_OK(true_camera_left.SetIntrinsics(true_camera_intrinsics));
  true_camera_left.SetExtrinsics(true_extrinsics_left);
  sensors::Camera true_camera_right;
  EXPECT_OK(true_camera_right.SetModel(kCameraModel));
  EXPECT_OK(true_camera_right.SetIntrinsics(true_camera_intrinsics));
  true_camera_right.SetExtrinsics(true_extrinsics_right);
  EXPECT_OK(true_camera_right.SetLatency(kRightCameraLatency));
  auto measurements_left = true_camera_left.Project(stamps, *trajectory_world_sensorrig, *world_model);
  ASSERT_OK(measurements_left.status());
  auto measurements_right = true_camera_right.Project(stamps, *trajectory_world_sensorrig, *world_model);
  ASSERT_OK(measurements_right.status());
  // Construct ground truth IMU and measurements.
  const auto kGyroscopeModel = sensors::GyroscopeIntrinsicsModel::kGyroscopeScaleAndBias;
  const auto kAccelerometerModel = sensors::AccelerometerIntrinsicsModel::kAccelerometerScaleAndBias;
  const double kImuRotationAngle = 2.0 * M_PI / 180.0, kGyroscopeLatency = 0.02, kAccelerometerLatency = 0.02;
  const VectorXd true_gyroscope_intrinsics{1.3, 0.01, -0.01, 0.01}, true_accelerometer_intrinsics{1.3, 0.01, -0.01, 0.01};
  Pose3d true_extrinsics_gyroscope, true_extrinsics_accelerometer;
  true_extrinsics_gyroscope.rotation() = AngleAxis(kImuRotationAngle, Normalized(RandomVector()));
  true_extrinsics_accelerometer.rotation() = AngleAxis(kImuRotationAngle, Normalized(RandomVector()));
  sensors::Gyroscope true_gyroscope;
  EXPECT_OK(true_gyroscope.SetModel(kGyroscopeModel));
  EXPECT_OK(true_gyroscope.SetIntrinsics(true_gyroscope_intrinsics));
  true_gyroscope.SetExtrinsics(true_extrinsics_gyroscope);
  EXPECT_OK(true_gyroscope.SetLatency(kGyroscopeLatency));
  auto measurements_gyroscope = true_gyroscope.Project(stamps, *trajectory_world_sensorrig, *world_model);
  ASSERT_OK(measurements_gyroscope.status());
  sensors::Accelerometer true_accelerometer;
  EXPECT_OK(true_accelerometer.SetModel(kAccelerometerModel));
  EXPECT_OK(true_accelerometer.SetIntrinsics(true_accelerometer_intrinsics));
  true_accelerometer.SetExtrinsics(true_extrinsics_accelerometer);
  EXPECT_OK(true_accelerometer.SetLatency(kAccelerometerLatency));
  auto measurements_accelerometer = true_accelerometer.Project(stamps, *trajectory_world_sensorrig, *world_model);
  ASSERT_OK(measurements_accelerometer.status());
  EXPECT_TRUE(measurements_gyroscope->size() == stamps.size() && measurements_left->size() > 0);

  // Create optimization sensors.
  VectorXd initial_camera_intrinsics = true_camera_intrinsics;
  for (auto& v : initial_camera_intrinsics) v *= 1.01;
  for (size_t i = 3; i < initial_camera_intrinsics.size(); ++i) initial_camera_intrinsics[i] = 0.0;
  Pose3d initial_extrinsics_right = true_extrinsics_right;
  { const Vector3d r = RandomVector(); for (int i = 0; i < 3; ++i) initial_extrinsics_right.translation()[i] += 0.01 * r[i]; }
  sensors::Camera* camera_left = new sensors::Camera();
  camera_left->SetName("Left");
  EXPECT_OK(camera_left->SetModel(kCameraModel));
  EXPECT_OK(camera_left->SetIntrinsics(initial_camera_intrinsics));
  camera_left->EnableExtrinsicsEstimation(false);
  camera_left->EnableIntrinsicsEstimation(true);
  camera_left->EnableLatencyEstimation(false);
  EXPECT_OK(camera_left->AddMeasurements(*measurements_left));
  sensors::Camera* camera_right = new sensors::Camera();
  camera_right->SetName("Right");
  EXPECT_OK(camera_right->SetModel(kCameraModel));
  EXPECT_OK(camera_right->SetIntrinsics(initial_camera_intrinsics));
  camera_right->SetExtrinsics(initial_extrinsics_right);
  camera_right->EnableExtrinsicsEstimation(true);
  camera_right->EnableIntrinsicsEstimation(true);
  camera_right->EnableLatencyEstimation(true);
  EXPECT_OK(camera_right->AddMeasurements(*measurements_right));
  VectorXd initial_gyroscope_intrinsics = true_gyroscope_intrinsics;
  for (auto& v : initial_gyroscope_intrinsics) v *= 1.01;
  sensors::Gyroscope* gyroscope = new sensors::Gyroscope();
  gyroscope->SetName("Gyroscope");
  EXPECT_OK(gyroscope->SetModel(kGyroscopeModel));
  EXPECT_OK(gyroscope->SetIntrinsics(initial_gyroscope_intrinsics));
  gyroscope->SetExtrinsics(true_extrinsics_gyroscope);
  gyroscope->EnableExtrinsicsEstimation(true);
  gyroscope->EnableIntrinsicsEstimation(true);
  gyroscope->EnableLatencyEstimation(true);
  EXPECT_OK(gyroscope->AddMeasurements(*measurements_gyroscope));
  VectorXd initial_accelerometer_intrinsics = true_accelerometer_intrinsics;
  for (auto& v : initial_accelerometer_intrinsics) v *= 1.01;
  Pose3d initial_accelerometer_extrinsics = true_extrinsics_accelerometer;
  { const Vector3d r = RandomVector(); for (int i = 0; i < 3; ++i) initial_accelerometer_extrinsics.translation()[i] += 0.05 * r[i]; }
  sensors::Accelerometer* accelerometer = new sensors::Accelerometer();
  accelerometer->SetName("Accelerometer");
  EXPECT_OK(accelerometer->SetModel(kAccelerometerModel));
  EXPECT_OK(accelerometer->SetIntrinsics(initial_accelerometer_intrinsics));
  accelerometer->SetExtrinsics(initial_accelerometer_extrinsics);
  accelerometer->EnableExtrinsicsEstimation(true);
  accelerometer->EnableIntrinsicsEstimation(true);
  accelerometer->EnableLatencyEstimation(true);
  EXPECT_OK(accelerometer->AddMeasurements(*measurements_accelerometer));

  // Construct optimization problem and optimize.
  BatchOptimizer optimizer;
  optimizer.AddSensor(camera_left);
  optimizer.AddSensor(camera_right);
  optimizer.AddSensor(gyroscope);
  optimizer.AddSensor(accelerometer);
  optimizer.AddWorldModel(world_model);
  optimizer.AddTrajectory(trajectory_world_sensorrig);
  SolverOptions options = DefaultSolverOptions();
  options.minimizer_progress_to_stdout = false;
  options.max_num_iterations = max_iterations;
  auto summary = optimizer.Optimize(options);
  ASSERT_OK(summary.status());
  std::printf("%s\n", summary->FullReport().c_str());
  if (max_iterations < 50) {      // emulated smoke variant: the cost must have dropped; acceptance is checked on the GPU run
    EXPECT_TRUE(summary->final_cost < 0.5 * summary->initial_cost && summary->num_successful_steps >= 1);
    std::printf(g_failures ? "TEST FAILED\n" : "TEST PASSED (emulated smoke variant)\n");
    return g_failures ? 1 : 0;
  }

  // Expect near perfect calibration results due to perfect data.
  const double kSmallNumber = 1e-7;
  EXPECT_TRUE(summary->termination_type == CONVERGENCE);
  EXPECT_TRUE(summary->final_cost < kSmallNumber);
  EXPECT_TRUE(MaxAbsDiff(true_camera_intrinsics, camera_left->GetIntrinsics()) < kSmallNumber);
  EXPECT_TRUE(MaxAbsDiff(true_camera_intrinsics, camera_right->GetIntrinsics()) < kSmallNumber);
  EXPECT_TRUE(PoseDiff(true_extrinsics_right, camera_right->GetExtrinsics()) < kSmallNumber);
  EXPECT_NEAR(kRightCameraLatency, camera_right->GetLatency(), kSmallNumber);
  EXPECT_TRUE(MaxAbsDiff(true_gyroscope_intrinsics, gyroscope->GetIntrinsics()) < kSmallNumber);
  EXPECT_TRUE(PoseDiff(true_extrinsics_gyroscope, gyroscope->GetExtrinsics()) < kSmallNumber);
  EXPECT_NEAR(kGyroscopeLatency, gyroscope->GetLatency(), kSmallNumber);
  EXPECT_TRUE(MaxAbsDiff(true_accelerometer_intrinsics, accelerometer->GetIntrinsics()) < kSmallNumber);
  EXPECT_TRUE(PoseDiff(true_extrinsics_accelerometer, accelerometer->GetExtrinsics()) < kSmallNumber);
  EXPECT_NEAR(kAccelerometerLatency, accelerometer->GetLatency(), kSmallNumber);
  for (int i = 0; i < 3; ++i) EXPECT_NEAR(world_model->gravity()[i], true_gravity[i], kSmallNumber);
  // Residual refresh (Sensor::UpdateResiduals, camera.cpp:70-80): one residual per measurement, all tiny at the optimum.
  auto pairs = camera_right->GetMeasurementResidualPairs();
  ASSERT_OK(pairs.status());
  EXPECT_TRUE(pairs->size() == measurements_right->size());
  double rmax = 0;
  for (const auto& pr : *pairs) rmax = std::max({rmax, std::fabs(pr.second[0]), std::fabs(pr.second[1])});
  EXPECT_TRUE(rmax < 1e-3);
  std::printf(g_failures ? "TEST FAILED\n" : "TEST PASSED\n");
  return g_failures ? 1 : 0;
}
