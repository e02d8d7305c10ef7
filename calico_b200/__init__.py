"""calico_b200 — B200-native drop-in for the hot path of yangjames/Calico's `calico::BatchOptimizer::Optimize()`.

The product is `libcalico_b200.so` (CUDA for sm_100a behind the C ABI of include/calico_b200.h); this package is the thin
host-side mirror used by the tests and the benchmark. There is no CPU fallback: `_capi.CApi()` raises if the library has
not been built, and every device entry point fails with status 13 (Internal) when no GPU is present.
"""
# Every entry point include/calico_b200.h declares (checked by tests/test_capi_symbols.py against the header itself).
C_ABI_SYMBOLS = [
    "cb2_default_options", "cb2_problem_create", "cb2_problem_destroy", "cb2_last_error", "cb2_set_trajectory", "cb2_set_gravity",
    "cb2_add_rigid_body", "cb2_add_sensor", "cb2_add_camera_observations", "cb2_add_imu_observations", "cb2_optimize",
    "cb2_evaluate_sensor", "cb2_cost", "cb2_get_sensor", "cb2_set_sensor", "cb2_get_trajectory", "cb2_get_rigid_body", "cb2_get_residuals",
    "cb2_comm_unique_id", "cb2_comm_init", "cb2_comm_clone", "cb2_shard_plan", "cb2_set_device", "cb2_stats_reset", "cb2_stats_get", "cb2_reset_parameters",
    "cb2_upload", "cb2_version", "cb2_num_intrinsics", "cb2_fit_spline_size", "cb2_fit_spline", "cb2_fit_trajectory", "cb2_fit_last_error",
]
