"""Kernel logic without a GPU: the SAME kernel sources compiled by g++ against tests/emul/cuda_emul.h (one OS thread per CUDA
thread) and driven through the same C ABI, compared with the oracle on a micro problem. This exercises indexing and
synchronisation of every kernel (sweep, accumulate/assemble, chunked band factor, Gram, level 2/3, back-substitution) — the
GPU parity tests proper are in test_gpu_parity.py. The emulated library is a test artefact and is never loaded by the package."""
import os
import subprocess
import sys

import numpy as np
import pytest

from calico_b200 import _capi, synthetic

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emul"))


@pytest.fixture(scope="module")
def emul_lib():
    import build as emul_build
    return emul_build.build()


def test_analytic_jacobians_match_dual_numbers_on_host():
    """cb2_functors.cuh compiled for the host against the oracle's dual numbers: all 7 camera models, 3 gyroscope and 3
    accelerometer models, generic and small-angle regimes (tests/emul/check_functors.cpp)."""
    import build as emul_build
    exe = emul_build.build_functor_check()
    out = subprocess.run([exe], capture_output=True, text=True)
    worst = [ln for ln in out.stdout.splitlines() if ln.startswith("WORST")]
    assert worst, out.stdout[-2000:]
    # generic regime: 1e-11; the theta ~ 1e-9 cases are limited by cancellation in the ORACLE's dual-number path
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("kind") and "theta_scale 1:" in ln]
    assert len(lines) == 13
    for ln in lines:
        assert float(ln.split()[-1]) < 1e-10, ln
    assert float(worst[0].split()[1]) < 1e-5


@pytest.mark.timeout(600)
def test_emulated_sweep_matches_oracle(emul_lib, oracle):
    truth, prob = synthetic.generate("micro", oracle.oracle_api, noise=True)
    a, o = _capi.CApi(emul_lib), oracle.oracle_api()
    ids_a, ids_o = prob.clone().push(a), prob.clone().push(o)
    for sa, so in zip(ids_a, ids_o):
        r1, J1, v1 = a.evaluate_sensor(sa)
        r2, J2, v2 = o.evaluate_sensor(so)
        assert (v1 == v2).all()
        assert np.abs(r1 - r2).max() <= 1e-9 * max(1.0, np.abs(r2).max())
        assert np.abs(J1 - J2).max() <= 1e-10 * np.abs(J2).max()
    c1, _ = a.cost()
    c2, _ = o.cost()
    assert abs(c1 - c2) <= 1e-12 * c2


@pytest.mark.timeout(900)
@pytest.mark.parametrize("schur,chunk", [("cr", None), ("cr", "6"), ("band", None), ("band", "6"), ("cr+jreread", None), ("cr+plainloop", None)])
def test_emulated_lm_iterations_match_oracle(schur, chunk, emul_lib, oracle, monkeypatch):
    """Three LM iterations through every kernel: level 1 by block cyclic reduction (default) and by the chunked band factor,
    single chunk and two chunks (+ separator level)."""
    if schur.endswith("+jreread"):         # every sensor through accumulate_kernel (J re-read) instead of the sweep's compact Gram slots
        schur = schur.split("+")[0]
        monkeypatch.setenv("CB2_NO_SWEEP_GRAM", "1")
    if schur.endswith("+plainloop"):       # no speculative trial sweep, no deferred host round trip: the textbook two-sync LM loop
        schur = schur.split("+")[0]
        monkeypatch.setenv("CB2_NO_SPECULATIVE_SWEEP", "1")
        monkeypatch.setenv("CB2_NO_DEFERRED_SYNC", "1")
    monkeypatch.setenv("CB2_SCHUR", schur)
    if chunk:
        monkeypatch.setenv("CB2_CHUNK_CPS", chunk)
    truth, prob = synthetic.generate("micro", oracle.oracle_api, noise=True)
    a, o = _capi.CApi(emul_lib), oracle.oracle_api()
    ids_a, ids_o = prob.clone().push(a), prob.clone().push(o)
    sum_a, log_a = a.optimize(_capi.Options(minimizer_progress_to_stdout=0, max_num_iterations=3))
    sum_o, log_o = o.optimize(oracle.OracleOptions(linear_solver=1, max_num_iterations=3))
    assert len(log_a) == len(log_o) == 4
    for x, y in zip(log_a, log_o):
        assert abs(x.cost - y.cost) <= 1e-9 * abs(y.cost)
        assert abs(x.step_norm - y.step_norm) <= 1e-7 * max(y.step_norm, 1e-12)
        assert abs(x.gradient_max_norm - y.gradient_max_norm) <= 1e-7 * y.gradient_max_norm
    pa, po = prob.clone(), prob.clone()
    pa.pull(a, ids_a)
    po.pull(o, ids_o)
    np.testing.assert_allclose(pa.spline.ctrl, po.spline.ctrl, rtol=1e-8, atol=1e-9)
    for s1, s2 in zip(pa.sensors, po.sensors):
        np.testing.assert_allclose(s1.intr, s2.intr, rtol=1e-8, atol=1e-10)


@pytest.mark.timeout(900)
@pytest.mark.parametrize("no_early_gram", [False, True])
def test_emulated_single_plain_sensor_and_gram_variants(no_early_gram, emul_lib, oracle, monkeypatch):
    """Camera + gyroscope only: accumulate_kernel packs FOUR segments into a CTA (one warp per (segment, sensor) pair; micro itself
    runs two per CTA), and the border Gram product once split by reduction level (default) and once as a single launch."""
    if no_early_gram:
        monkeypatch.setenv("CB2_NO_EARLY_GRAM", "1")
    truth, prob = synthetic.generate("micro", oracle.oracle_api, noise=True)
    prob.sensors = prob.sensors[:-1]
    assert [s.kind for s in prob.sensors] == [0, 1]
    a, o = _capi.CApi(emul_lib), oracle.oracle_api()
    prob.clone().push(a)
    prob.clone().push(o)
    sum_a, log_a = a.optimize(_capi.Options(minimizer_progress_to_stdout=0, max_num_iterations=3))
    sum_o, log_o = o.optimize(oracle.OracleOptions(linear_solver=1, max_num_iterations=3))
    assert len(log_a) == len(log_o) == 4
    for x, y in zip(log_a, log_o):
        assert abs(x.cost - y.cost) <= 1e-9 * abs(y.cost)
        assert abs(x.step_norm - y.step_norm) <= 1e-7 * max(y.step_norm, 1e-12)
        assert abs(x.gradient_max_norm - y.gradient_max_norm) <= 1e-7 * y.gradient_max_norm


@pytest.mark.timeout(900)
def test_host_packing_is_independent_of_its_thread_split(emul_lib, oracle, monkeypatch):
    """upload() packs a large sensor with several host threads (ranges of observations): same packed problem — hence bitwise the same
    cost, residuals (caller's order) and Jacobians — whatever the split, on shuffled input with flagged outliers, and through
    concurrent cb2_add_*_observations calls (ProblemSpec.push with one thread per sensor)."""
    from calico_b200 import spec as spec_mod
    truth, prob = synthetic.generate("tiny", oracle.oracle_api, noise=True)
    rng = np.random.default_rng(5)
    for s in prob.sensors:                       # caller's order is arbitrary (absl::flat_hash_map iteration in the reference, camera.cpp:120)
        order = rng.permutation(s.n_obs)
        s.stamp, s.meas = s.stamp[order], s.meas[order]
        if s.kind == 0:
            s.image_id, s.model_id, s.feature_id = s.image_id[order], s.model_id[order], s.feature_id[order]
            s.outlier = (rng.random(s.n_obs) < 0.1).astype(np.uint8)
        elif s.seq is not None:
            s.seq = s.seq[order]
    results = []
    for split, threads, par in (("1000000000", "1", 10**9), ("1", "4", 10**9), ("1", "3", 1)):
        monkeypatch.setenv("CB2_PACK_SPLIT_MIN", split)
        monkeypatch.setenv("CB2_PACK_SUB_THREADS", threads)
        monkeypatch.setattr(spec_mod, "PARALLEL_OBS", par)
        a = _capi.CApi(emul_lib)
        p = prob.clone()
        ids = p.push(a)
        ev = [a.evaluate_sensor(sid) for sid in ids]
        summ, log = a.optimize(_capi.Options(minimizer_progress_to_stdout=0, max_num_iterations=2))
        res = p.residuals(a, ids)
        results.append((ev, [x.cost for x in log], res, summ.num_residual_blocks))
        a.close()
    ev0, cost0, res0, nb0 = results[0]
    assert nb0 == sum(int(s.n_obs - (s.outlier.sum() if s.outlier is not None else 0)) for s in prob.sensors)
    for ev, cost, res, nb in results[1:]:
        assert nb == nb0 and cost == cost0
        for (r0, J0, v0), (r1, J1, v1) in zip(ev0, ev):
            assert np.array_equal(r0, r1) and np.array_equal(J0, J1) and np.array_equal(v0, v1)
        for (r0, v0), (r1, v1) in zip(res0, res):
            assert np.array_equal(r0, r1) and np.array_equal(v0, v1)


@pytest.mark.timeout(900)
def test_emulated_lm_with_rejected_steps_matches_oracle(emul_lib, oracle):
    """The reference's toy stereo + IMU problem starts with three rejected steps (as in the stored Ceres log): exercises the
    reject -> accept transitions of the speculative trial sweep (the candidate buffer is reused by every new candidate)."""
    truth, prob = synthetic.toy_stereo_imu_problem(oracle.oracle_api, seed=3)
    a, o = _capi.CApi(emul_lib), oracle.oracle_api()
    prob.clone().push(a)
    prob.clone().push(o)
    sum_a, log_a = a.optimize(_capi.Options(minimizer_progress_to_stdout=0, max_num_iterations=5))
    sum_o, log_o = o.optimize(oracle.OracleOptions(linear_solver=1, max_num_iterations=5))
    assert [x.step_is_successful for x in log_a] == [x.step_is_successful for x in log_o] == [1, 0, 0, 0, 1, 0]
    for x, y in zip(log_a, log_o):
        assert abs(x.cost - y.cost) <= 1e-9 * abs(y.cost)
