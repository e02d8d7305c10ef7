"""GPU parity on the BASELINE.json shapes themselves (the shapes the metric is quoted on), through the C ABI, against the CPU oracle
(banded-Schur solver: same LM step as the Ceres-ordered dense Schur up to rounding, tests/test_oracle_pins.py).

  C2   1 OpenCv5 camera + IMU, 500 frames            full LM run
  C3   4 KannalaBrandt cameras, 2000 frames          full LM run
  C4   8 OpenCv5 cameras + IMU, 5000 frames, 1.1 M   first 3 LM iterations (cost, accept/reject, radius, gradient norms)
  C5_like  16 OpenCv5 cameras + 2 IMUs, Huber + 2 % outliers, N_c = 279 > 208: the calibration border is too wide for the DMMA Gram kernel
       and the shared-memory reduced solve, so border_gram_kernel (SIMT) and reduced_solve_kernel (global memory) are the ones compared
  small_cauchy  ceres::CauchyLoss (optimization_utils.h:40-41) with outliers
Tolerance: 1e-6 relative on per-iteration cost and converged parameters (north star)."""
import os

import numpy as np
import pytest

from calico_b200 import _capi, synthetic
from test_gpu_parity import REL, _compare_runs

pytestmark = pytest.mark.gpu


def _run_both(cfg, oracle, product_lib, **opts):
    truth, prob = synthetic.generate(cfg, oracle.oracle_api, noise=True)
    a, o = _capi.CApi(product_lib), oracle.oracle_api()
    ids_a, ids_o = prob.clone().push(a), prob.clone().push(o)
    sum_a, log_a = a.optimize(_capi.Options(minimizer_progress_to_stdout=0, **opts))
    sum_o, log_o = o.optimize(oracle.OracleOptions(linear_solver=1, num_threads=os.cpu_count() or 1, **opts))
    return prob, a, o, ids_a, ids_o, sum_a, log_a, sum_o, log_o


@pytest.mark.parametrize("cfg", ["C2", "C3", "C5_like", "small_cauchy"])
def test_full_lm_run_matches_oracle(cfg, oracle, product_lib):
    prob, a, o, ids_a, ids_o, sum_a, log_a, sum_o, log_o = _run_both(cfg, oracle, product_lib)
    assert sum_a.termination_type == _capi.CONVERGENCE
    assert sum_a.message == sum_o.message or sum_a.message.split(b":")[0] == sum_o.message.split(b":")[0]
    _compare_runs(a, o, prob, ids_a, ids_o, log_a, log_o, sum_a, sum_o)
    if cfg == "C5_like":
        n_c = sum(int(s.en_intr) * s.intr.size + 6 * int(s.en_extr) + int(s.en_lat) for s in prob.sensors)
        assert n_c > 208   # the point of this shape: the wide-border fallback kernels ran


def test_c4_first_iterations_match_oracle(oracle, product_lib):
    """The config the metric is quoted on: 1 099 982 residual blocks, 2 505 control points. Three LM iterations: cost, acceptance, radius
    and gradient norms of every iteration against the oracle; then the un-robustified residuals of the final point."""
    prob, a, o, ids_a, ids_o, sum_a, log_a, sum_o, log_o = _run_both("C4", oracle, product_lib, max_num_iterations=3)
    assert sum_a.num_residual_blocks == sum_o.num_residual_blocks == 1099982
    assert sum_a.termination_type == sum_o.termination_type == _capi.NO_CONVERGENCE
    assert len(log_a) == len(log_o) == 4
    for x, y in zip(log_a, log_o):
        assert x.step_is_successful == y.step_is_successful
        assert abs(x.cost - y.cost) <= REL * abs(y.cost)
        assert abs(x.gradient_max_norm - y.gradient_max_norm) <= 1e-5 * y.gradient_max_norm
        assert abs(x.gradient_norm - y.gradient_norm) <= 1e-5 * y.gradient_norm
        assert abs(x.step_norm - y.step_norm) <= 1e-5 * max(y.step_norm, 1e-12)
        assert abs(x.trust_region_radius - y.trust_region_radius) <= 1e-4 * y.trust_region_radius
    _compare_runs(a, o, prob, ids_a, ids_o, log_a, log_o, sum_a, sum_o)


def test_start_at_the_optimum_terminates_like_ceres(oracle, product_lib):
    """A problem that starts at a converged point: the first step is tiny, and Ceres's ParameterToleranceReached / FunctionToleranceReached
    (trust_region_minimizer.cc) carry no "at least one successful step" guard, so the solve stops at iteration 1."""
    truth, prob = synthetic.generate("tiny", oracle.oracle_api, noise=True)
    a, o = _capi.CApi(product_lib), oracle.oracle_api()
    pa, po = prob.clone(), prob.clone()
    ids_a, ids_o = pa.push(a), po.push(o)
    a.optimize(_capi.Options(minimizer_progress_to_stdout=0))
    pa.pull(a, ids_a)
    # second call warm-starts at the optimum (batch_optimizer.cpp:57: a new problem per call, values from the objects)
    a2, o2 = _capi.CApi(product_lib), oracle.oracle_api()
    ids_a2, ids_o2 = pa.clone().push(a2), pa.clone().push(o2)
    sum_a, log_a = a2.optimize(_capi.Options(minimizer_progress_to_stdout=0))
    sum_o, log_o = o2.optimize(oracle.OracleOptions(linear_solver=1))
    assert sum_a.termination_type == sum_o.termination_type == _capi.CONVERGENCE
    assert len(log_a) == len(log_o) <= 3
    assert sum_a.message.split(b":")[0] == sum_o.message.split(b":")[0]
