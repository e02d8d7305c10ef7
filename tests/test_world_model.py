"""Freed world-model blocks: RigidBody::world_pose_is_constant / model_definition_is_constant = false (world_model.h:41-69). The reference
registers the chart pose (translation + quaternion with EigenQuaternionManifold) and every model point as free parameter blocks
(world_model.cpp:52-70, camera_cost_functor.h parameters t_model_point, q_world_model, t_world_model); here they join the calibration vector
of the reduced system. Same LM trajectory and converged values as the oracle, through the C ABI: on the SIMT emulation (CPU) and on the GPU."""
import os
import sys
import threading

import numpy as np
import pytest

from calico_b200 import _capi, synthetic

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emul"))


def _free_world(prob, mode, seed=5):
    b = prob.bodies[0]
    rng = np.random.default_rng(seed)
    if "pose" in mode:
        b.pose_const = False
        b.t = b.t + 0.01 * rng.standard_normal(3)
    if "pts" in mode:
        b.model_const = False
        b.pts = b.pts + 0.002 * rng.standard_normal(np.asarray(b.pts).shape)
    return prob


def _check(lib, oracle, cfg, mode, iters, rel):
    truth, prob = synthetic.generate(cfg, oracle.oracle_api, noise=True)
    prob = _free_world(prob, mode)
    a, o = _capi.CApi(lib), oracle.oracle_api()
    pa, po = prob.clone(), prob.clone()
    ia, io = pa.push(a), po.push(o)
    sa, la = a.optimize(_capi.Options(minimizer_progress_to_stdout=0, max_num_iterations=iters))
    so, lo = o.optimize(oracle.OracleOptions(linear_solver=0, max_num_iterations=iters, num_threads=os.cpu_count() or 1))
    assert sa.termination_type == so.termination_type
    assert len(la) == len(lo)
    for x, y in zip(la, lo):
        assert x.step_is_successful == y.step_is_successful
        assert abs(x.cost - y.cost) <= rel * abs(y.cost)
        assert abs(x.gradient_max_norm - y.gradient_max_norm) <= 1e-5 * y.gradient_max_norm
    for f in ("num_parameter_blocks_reduced", "num_parameters_reduced", "num_effective_parameters_reduced", "num_residual_blocks"):
        assert getattr(sa, f) == getattr(so, f), f
    pa.pull(a, ia)
    po.pull(o, io)
    ba, bo = pa.bodies[0], po.bodies[0]
    np.testing.assert_allclose(ba.t, bo.t, rtol=0, atol=1e-7)
    np.testing.assert_allclose(ba.q_xyzw, bo.q_xyzw, rtol=0, atol=1e-7)
    np.testing.assert_allclose(ba.pts, bo.pts, rtol=0, atol=1e-7)
    if "pose" in mode:
        assert np.abs(np.asarray(ba.t) - np.asarray(prob.bodies[0].t)).max() > 1e-6      # the block really moved
    if "pts" in mode:
        assert np.abs(np.asarray(ba.pts) - np.asarray(prob.bodies[0].pts)).max() > 1e-6
    for s1, s2 in zip(pa.sensors, po.sensors):
        np.testing.assert_allclose(s1.intr, s2.intr, rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(pa.spline.ctrl, po.spline.ctrl, rtol=1e-6, atol=1e-7)


@pytest.mark.timeout(900)
@pytest.mark.parametrize("mode", ["pose", "pts", "pose+pts"])
def test_emulated_freed_world_blocks_match_oracle(mode, oracle):
    import build as emul_build
    _check(emul_build.build(), oracle, "micro", mode, 3 if mode == "pose" else 2, 1e-8)


@pytest.mark.timeout(900)
def test_emulated_freed_pose_two_ranks(oracle, monkeypatch):
    """The world unknowns are shared by every rank like the calibration blocks."""
    import build as emul_build
    lib = emul_build.build()
    monkeypatch.setenv("CB2_CHUNK_CPS", "6")
    truth, prob = synthetic.generate("tiny", oracle.oracle_api, noise=True)
    prob = _free_world(prob, "pose")
    o = oracle.oracle_api()
    po = prob.clone()
    io = po.push(o)
    so, lo = o.optimize(oracle.OracleOptions(linear_solver=0, max_num_iterations=3))
    po.pull(o, io)
    res, errors = [None, None], []

    def run(rank):
        try:
            a = _capi.CApi(lib)
            pa = prob.clone()
            ids = pa.push(a)
            a.comm_init_local(2, rank, "world2")
            s, lg = a.optimize(_capi.Options(minimizer_progress_to_stdout=0, max_num_iterations=3))
            pa.pull(a, ids)
            res[rank] = (lg, pa)
        except Exception as e:   # noqa: BLE001
            errors.append(e)
    th = [threading.Thread(target=run, args=(r,)) for r in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors, errors
    for lg, pa in res:
        assert len(lg) == len(lo)
        for x, y in zip(lg, lo):
            assert abs(x.cost - y.cost) <= 1e-8 * y.cost
        np.testing.assert_allclose(pa.bodies[0].t, po.bodies[0].t, rtol=0, atol=1e-8)
        np.testing.assert_allclose(pa.bodies[0].q_xyzw, po.bodies[0].q_xyzw, rtol=0, atol=1e-8)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,mode", [("tiny", "pose"), ("tiny", "pose+pts"), ("small", "pose")])
def test_freed_world_blocks_match_oracle(cfg, mode, oracle, product_lib):
    _check(product_lib, oracle, cfg, mode, 50, 1e-6)
