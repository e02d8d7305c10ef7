/* calico_b200 — C ABI of the B200-native batch calibration optimizer.
 *
 * This is the drop-in boundary for ONE path of yangjames/Calico: calico::BatchOptimizer::Optimize()
 * (reference calico/batch_optimizer.h:62-63, batch_optimizer.cpp:53-81) and what it pulls in through the
 * Sensor plugin interface (calico/sensors/sensor_base.h:22-102): parameter registration, one residual block
 * per observation, the Ceres LM solve, and the residual refresh. Everything below is plain C: opaque handle,
 * pointers and sizes; no C++/torch/Eigen types cross it.  All floating point data is FP64.
 *
 * Conventions shared with the reference:
 *   - quaternions are stored x,y,z,w (Eigen coeffs() order, calico/typedefs.h:69-81);
 *   - poses are T_sensorrig_sensor / T_world_rigidbody exactly as in the reference;
 *   - enum integer values are those of the reference enums (cited beside each).
 *   - return value = absl::StatusCode value the reference would have produced:
 *     0 OK, 3 InvalidArgument, 9 FailedPrecondition, 12 Unimplemented, 13 Internal (also CUDA/NCCL faults);
 *     cb2_last_error() returns the message.
 */
#ifndef CALICO_B200_H_
#define CALICO_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cb2_problem cb2_problem;

enum { CB2_OK = 0, CB2_INVALID_ARGUMENT = 3, CB2_FAILED_PRECONDITION = 9, CB2_UNIMPLEMENTED = 12, CB2_INTERNAL = 13 };

/* sensor kind — which reference class the sensor replaces: sensors/camera.h, gyroscope.h, accelerometer.h */
enum { CB2_CAMERA = 0, CB2_GYROSCOPE = 1, CB2_ACCELEROMETER = 2 };
/* calico::sensors::CameraIntrinsicsModel, sensors/camera_models.h:16-33 */
enum { CB2_CAM_NONE = 0, CB2_CAM_OPENCV5 = 1, CB2_CAM_OPENCV8 = 2, CB2_CAM_KANNALA_BRANDT = 3, CB2_CAM_DOUBLE_SPHERE = 4,
       CB2_CAM_FIELD_OF_VIEW = 5, CB2_CAM_UNIFIED = 6, CB2_CAM_EXTENDED_UNIFIED = 7 };
/* calico::sensors::{Accelerometer,Gyroscope}IntrinsicsModel, sensors/accelerometer_models.h:16-25, gyroscope_models.h:16-25 */
enum { CB2_IMU_NONE = 0, CB2_IMU_SCALE_ONLY = 1, CB2_IMU_SCALE_AND_BIAS = 2, CB2_IMU_VECTORNAV = 3 };
/* calico::utils::LossFunctionType, optimization_utils.h:15-22 */
enum { CB2_LOSS_NONE = 0, CB2_LOSS_HUBER = 1, CB2_LOSS_CAUCHY = 2 };
/* ceres::TerminationType (Ceres external) as read by batch_optimizer_test.cpp:186 */
enum { CB2_CONVERGENCE = 0, CB2_NO_CONVERGENCE = 1, CB2_FAILURE = 2 };

/* The subset of ceres::Solver::Options the reference exposes (calico/calico.cpp:378-394) plus the trust-region
 * constants Ceres applies by default. cb2_default_options() == calico::DefaultSolverOptions()
 * (batch_optimizer.cpp:10-17): DENSE_SCHUR, function_tolerance 1e-8, parameter_tolerance 1e-10. */
typedef struct cb2_options {
  int32_t max_num_iterations;            /* 50 */
  double function_tolerance;             /* 1e-8  (batch_optimizer.cpp:14) */
  double gradient_tolerance;             /* 1e-10 */
  double parameter_tolerance;            /* 1e-10 (batch_optimizer.cpp:15) */
  double initial_trust_region_radius;    /* 1e4 */
  double max_trust_region_radius;        /* 1e16 */
  double min_trust_region_radius;        /* 1e-32 */
  double min_relative_decrease;          /* 1e-3 */
  double min_lm_diagonal;                /* 1e-6 */
  double max_lm_diagonal;                /* 1e32 */
  int32_t max_num_consecutive_invalid_steps; /* 5 */
  int32_t jacobi_scaling;                /* 1 */
  int32_t num_threads;                   /* host threads for packing; the solve runs on the GPU */
  int32_t minimizer_progress_to_stdout;  /* 1 in DefaultSolverOptions (batch_optimizer.cpp:13) */
  int32_t linear_solver;                 /* ignored: always block-banded Schur + dense reduced solve on device */
  int32_t use_cuda_graph;                /* reserved (0) */
} cb2_options;

/* One row of Ceres's minimizer progress table (the columns of demos/imu_camera_calibration.ipynb:350). */
typedef struct cb2_iteration {
  int32_t iteration;
  double cost, cost_change, gradient_max_norm, gradient_norm, step_norm, relative_decrease, trust_region_radius;
  int32_t step_is_valid, step_is_successful;
  double iteration_time;                 /* seconds, host clock around the iteration */
} cb2_iteration;

/* The ceres::Solver::Summary fields the reference exposes (calico/calico.cpp:352-375) + termination_type. */
typedef struct cb2_summary {
  int32_t termination_type;
  double initial_cost, final_cost, fixed_cost;
  int32_t num_successful_steps, num_unsuccessful_steps, num_iterations;
  int32_t num_parameter_blocks, num_parameters, num_effective_parameters, num_residual_blocks, num_residuals;
  int32_t num_parameter_blocks_reduced, num_parameters_reduced, num_effective_parameters_reduced,
      num_residual_blocks_reduced, num_residuals_reduced;
  double jacobian_time, linear_solver_time, total_time;   /* seconds; device phases timed with CUDA events */
  char message[256];
} cb2_summary;

/* Device-side accounting for bench.py (not part of the reference surface). */
typedef struct cb2_stats {
  int64_t kernel_launches;        /* kernels of this library launched since cb2_stats_reset */
  int64_t jacobian_sweeps;        /* residual+Jacobian sweeps (K1-K3 over all observations) */
  int64_t jacobian_blocks;        /* residual blocks evaluated with Jacobians */
  double jacobian_kernel_ms;      /* CUDA-event time spent in K1-K3 */
  double jacobian_bytes;          /* algorithmic bytes moved by K1-K3 (SURVEY §8d definition) */
  double normal_eq_ms, schur_ms, cost_eval_ms;   /* CUDA-event time in K4, K5-K7 (+ step update), K8 */
  int64_t h2d_bytes, d2h_bytes;
  double lm_loop_ms;              /* CUDA-event time from the first to the last kernel of the LM loop(s) */
  int64_t lm_iterations;          /* LM iterations (accepted + rejected) run since reset */
  double camera_kernel_ms;        /* CUDA-event time of the camera residual+Jacobian kernel alone (the dominant kernel of the sweep) */
  double camera_kernel_bytes;     /* its algorithmic bytes (camera blocks only), same definition as jacobian_bytes */
  double camera_kernel_gram_bytes; /* what the same kernel writes BESIDES the Jacobian: compact per-image Gram slots + per-warp calibration partials */
} cb2_stats;

void cb2_default_options(cb2_options* out);

/* ---- problem assembly: replaces BatchOptimizer::Add{Sensor,WorldModel,Trajectory} (batch_optimizer.h:33,40,54)
 *      + the Add{Parameters,Residuals}ToProblem loops (camera.cpp:92-153, gyroscope.cpp:10-54,
 *      accelerometer.cpp:10-56, world_model.cpp:40-77, bspline.hpp:10-17) by one SoA hand-over. ---- */
int cb2_problem_create(cb2_problem** out);
void cb2_problem_destroy(cb2_problem* p);
const char* cb2_last_error(cb2_problem* p);

/* Trajectory (trajectory.h:27-120, bspline.h): full knot vector (n_knots = n_cp + spline_order, bspline.hpp:164-180)
 * and control points [n_cp][6] = [axis-angle phi_world_rig ; t_world_rig]. Basis matrices are derived inside by the
 * reference's recursion (bspline.hpp:192-244). Control points are always estimated (bspline.hpp:10-17). */
int cb2_set_trajectory(cb2_problem* p, int spline_order, int n_knots, const double* knots, int n_cp, const double* ctrl);
/* WorldModel gravity (world_model.h:78,111); never estimated (world_model.cpp:79-81). */
int cb2_set_gravity(cb2_problem* p, const double* g3);
/* WorldModel::AddRigidBody (world_model.cpp:29-38, struct world_model.h:41-69). With a false constness flag the block is estimated
 * (world_model.cpp:52-70): the pose as translation + quaternion (EigenQuaternionManifold, optimization_utils.h:51-61), each model point
 * as a 3-vector — extra unknowns of the reduced (calibration) system shared by all cameras. */
int cb2_add_rigid_body(cb2_problem* p, int id, const double* q_xyzw, const double* t3, int n_pts, const int* feature_ids,
                       const double* pts_xyz, int world_pose_is_constant, int model_definition_is_constant);
/* A sensor with its model, state and flags: Camera/Gyroscope/Accelerometer setters of sensor_base.h:27-101
 * (SetModel, SetIntrinsics, SetExtrinsics, SetLatency, SetMeasurementNoise, SetLossFunction, Enable*Estimation). */
int cb2_add_sensor(cb2_problem* p, int kind, int model, const char* name, int n_intr, const double* intr, const double* q_xyzw,
                   const double* t3, double latency, double sigma, int loss_type, double loss_scale, int enable_intrinsics,
                   int enable_extrinsics, int enable_latency, int* sensor_id);
/* Camera::AddMeasurements (camera.cpp:224-254) with the outlier set (camera.cpp:281-299) as a mask. id fields are
 * CameraObservationId (camera.h:24-50). model_id must name a rigid body (camera.cpp:125-131).
 * Threading: a handle is single-caller, with two exceptions made for large problems — once every sensor and rigid body has been
 * added, cb2_add_camera_observations / cb2_add_imu_observations may be called concurrently for DIFFERENT sensor ids, and after
 * cb2_optimize cb2_get_residuals may be called concurrently (plain copies out of the result block). */
int cb2_add_camera_observations(cb2_problem* p, int sensor_id, int n, const double* stamp, const int* image_id, const int* model_id,
                                const int* feature_id, const double* pixel_xy, const uint8_t* outlier_mask);
/* Gyroscope/Accelerometer::AddMeasurements (gyroscope.cpp, accelerometer.cpp); ids are {stamp, sequence}. */
int cb2_add_imu_observations(cb2_problem* p, int sensor_id, int n, const double* stamp, const int* sequence, const double* xyz);

/* ---- the hot path ---- */
/* calico::BatchOptimizer::Optimize (batch_optimizer.cpp:53-81): packs and uploads the problem (if changed), runs the LM
 * loop on the device, writes optimised parameters back into the handle, refreshes per-observation residuals
 * (Sensor::UpdateResiduals, camera.cpp:70-80). log may be NULL; *n_log receives the number of iterations recorded. */
int cb2_optimize(cb2_problem* p, const cb2_options* opts, cb2_summary* summary, cb2_iteration* log, int log_cap, int* n_log);

/* Analogue of ceres::Problem::Evaluate as the reference tests use it (accelerometer_test.cpp:179-203): per sensor,
 * un-robustified residuals [n_obs][m] and Jacobians [n_obs][m][W] in the canonical column order
 * [control points 6k | intrinsics | extrinsic rotation (tangent 3) | extrinsic translation 3 | latency], W = 6k+n_intr+7,
 * valid[n_obs] != 0 = functor returned true; for cameras bit 1 (value 2) is set in addition when the point lies behind the image plane
 * (p_c.z <= 0), which Camera::Project skips (camera.cpp:172-174,186-188). Any output may be NULL. Outlier observations are left untouched. */
int cb2_evaluate_sensor(cb2_problem* p, int sensor_id, double* residuals, double* jacobians, uint8_t* valid);
/* 1/2 sum rho(|r|^2) over all non-outlier residual blocks at the current state; *ok = 0 if any functor fails. */
int cb2_cost(cb2_problem* p, double* cost, int* ok);

/* ---- write-back: the reference mutates the user's objects in place through raw pointers (camera.cpp:98-101) ---- */
int cb2_get_sensor(cb2_problem* p, int sensor_id, double* intr, double* q_xyzw, double* t3, double* latency);
int cb2_set_sensor(cb2_problem* p, int sensor_id, const double* intr, const double* q_xyzw, const double* t3, double latency);
int cb2_get_trajectory(cb2_problem* p, double* ctrl);
/* RigidBody write-back when world_pose_is_constant / model_definition_is_constant are false (world_model.cpp:52-70: the reference estimates
 * those blocks and Ceres mutates them in place): q_xyzw[4], t3[3], pts_xyz[n_pts][3] in the order given to cb2_add_rigid_body. Any may be NULL. */
int cb2_get_rigid_body(cb2_problem* p, int id, double* q_xyzw, double* t3, double* pts_xyz);
/* Camera::GetMeasurementResidualPairs source (camera.cpp:258-279): residuals [n_obs][m] in observation order, valid mask. */
int cb2_get_residuals(cb2_problem* p, int sensor_id, double* residuals, uint8_t* valid);

/* ---- multi-GPU: observations are sharded by contiguous spline-segment ranges, one process per GPU; the only exchange
 *      is one allreduce of the reduced system per LM iteration. unique_id is ncclUniqueId bytes (128). ---- */
int cb2_comm_unique_id(uint8_t* id128);
int cb2_comm_init(cb2_problem* p, int world_size, int rank, const uint8_t* id128);
int cb2_set_device(int device);
/* A second handle of the same process re-uses the first one's communicator (communicator creation is one-time setup). */
int cb2_comm_clone(cb2_problem* dst, cb2_problem* src);
/* Host-side shard plan (no device needed): the chunks [chunk_lo, chunk_hi) of n_chunks and the spline segments
 * [seg_lo, seg_hi) whose observations rank `rank` of `world_size` evaluates. Every rank is handed the whole problem and keeps
 * its shard; after cb2_optimize every rank holds EVERY observation's residual (cb2_get_residuals: one cross-rank sum of the scattered
 * residual arrays, as Sensor::UpdateResiduals fills every measurement, camera.cpp:70-80); cb2_evaluate_sensor covers the local shard only. */
int cb2_shard_plan(cb2_problem* p, int world_size, int rank, int* n_chunks, int* chunk_lo, int* chunk_hi, int* seg_lo, int* seg_hi);

/* ---- trajectory spline fit, the step before the hot path (no problem handle; errors through cb2_fit_last_error) ----
 * BSpline<6, double>::FitToData (bspline.hpp:20-38): validation of CheckDataForSplineFit (bspline.hpp:299-327), knot vector of
 * ComputeKnotVector (bspline.hpp:164-180), basis matrices (bspline.hpp:192-244) and the least-squares fit of FitSpline
 * (bspline.hpp:247-297) — solved on the device as the banded SPD system the reference's TODO (bspline.hpp:287-289) describes.
 * cb2_fit_spline_size returns the sizes the caller must allocate: n_knots = n_valid + 2 (order - 1), n_cp = n_knots - order. */
int cb2_fit_spline_size(int n, const double* times, int spline_order, double knot_frequency, int* n_knots, int* n_cp);
/* times[n] ascending, data6[n][6]; knots_out[n_knots], ctrl_out[n_cp][6]. */
int cb2_fit_spline(int n, const double* times, const double* data6, int spline_order, double knot_frequency, int n_knots, double* knots_out,
                   int n_cp, double* ctrl_out);
/* Trajectory::FitSpline (trajectory.cpp:14-49): poses (stamp, q_xyzw world<-rig, t_world_rig) in any order -> sorted by stamp,
 * rotation -> axis * angle (Eigen::AngleAxisd), UnwrapPhaseLogMap (trajectory.cpp:81-93), then cb2_fit_spline on [phi ; t].
 * Sizes from cb2_fit_spline_size on the sorted stamps (only first and last matter). */
int cb2_fit_trajectory(int n, const double* stamps, const double* q_xyzw, const double* t3, int spline_order, double knot_frequency, int n_knots,
                       double* knots_out, int n_cp, double* ctrl_out);
const char* cb2_fit_last_error(void);

/* ---- bench accounting ---- */
int cb2_stats_reset(cb2_problem* p);
int cb2_stats_get(cb2_problem* p, cb2_stats* out);
/* Restores every estimated parameter to the values it had when the problem was last uploaded (device-side copy). */
int cb2_reset_parameters(cb2_problem* p);
/* Upload now (otherwise done lazily by the first optimize/evaluate). */
int cb2_upload(cb2_problem* p);
const char* cb2_version(void);
/* Number of intrinsics of a sensor model — CameraModel::NumberOfParameters (camera_models.h:79,231,395,596,716,848,961: 8, 11, 7, 5, 4,
 * 4, 5) and the IMU models (accelerometer_models.h / gyroscope_models.h: 1, 4, 12); -1 for kNone / an unknown kind or model. */
int cb2_num_intrinsics(int kind, int model);

#ifdef __cplusplus
}
#endif
#endif /* CALICO_B200_H_ */
