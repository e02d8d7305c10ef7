"""Edge cases of the hot path the reference handles (or rejects) explicitly, on the SIMT emulation of the kernel sources (CPU) and on
the GPU: ragged observation sets (segments with no observations, sensors with none), IMU-only and camera-only problems, an all-outlier
camera, the shortest admissible trajectory, and the error statuses of problem assembly."""
import os
import sys

import numpy as np
import pytest

from calico_b200 import _capi, synthetic

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emul"))


def _lib(kind, product_lib=None):
    if kind == "gpu":
        return product_lib
    import build as emul_build
    return emul_build.build()


def _compare(lib, oracle, prob, iters=3, rtol=1e-8):
    a, o = _capi.CApi(lib), oracle.oracle_api()
    prob.clone().push(a)
    prob.clone().push(o)
    sa, la = a.optimize(_capi.Options(minimizer_progress_to_stdout=0, max_num_iterations=iters))
    so, lo = o.optimize(oracle.OracleOptions(linear_solver=1, max_num_iterations=iters))
    assert len(la) == len(lo)
    assert [x.step_is_successful for x in la] == [x.step_is_successful for x in lo]
    for x, y in zip(la, lo):
        assert abs(x.cost - y.cost) <= rtol * abs(y.cost) + 1e-18
    assert sa.num_residual_blocks == so.num_residual_blocks and sa.num_parameters_reduced == so.num_parameters_reduced
    return sa


def _cases(oracle):
    truth, prob = synthetic.generate("micro", oracle.oracle_api, noise=True)
    cases = {}
    # ragged: drop every observation of the middle third of the time span (whole spline segments become empty)
    p = prob.clone()
    t_lo, t_hi = 0.4, 0.8
    for s in p.sensors:
        keep = (np.asarray(s.stamp) < t_lo) | (np.asarray(s.stamp) > t_hi)
        for name in ("stamp", "meas", "image_id", "model_id", "feature_id", "seq", "outlier"):
            v = getattr(s, name)
            if v is not None:
                setattr(s, name, np.asarray(v)[keep])
    cases["ragged"] = p
    # IMU only / camera only
    p = prob.clone(); p.sensors = [s for s in p.sensors if s.kind != synthetic.CAMERA]; cases["imu_only"] = p
    p = prob.clone(); p.sensors = [s for s in p.sensors if s.kind == synthetic.CAMERA]; cases["camera_only"] = p
    # a sensor without observations stays in the problem (its blocks are simply never referenced)
    p = prob.clone()
    for s in p.sensors:
        if s.kind == synthetic.GYROSCOPE:
            for name in ("stamp", "meas", "seq"):
                v = getattr(s, name)
                if v is not None:
                    setattr(s, name, np.asarray(v)[:0])
    cases["empty_sensor"] = p
    # every second camera observation marked as an outlier (camera.cpp:121-124)
    p = prob.clone()
    for s in p.sensors:
        if s.kind == synthetic.CAMERA:
            s.outlier = (np.arange(s.n_obs) % 2 == 0).astype(np.uint8)
    cases["half_outliers"] = p
    return cases


@pytest.mark.timeout(900)
@pytest.mark.parametrize("case", ["ragged", "imu_only", "camera_only", "empty_sensor", "half_outliers"])
def test_edge_cases_emulated(case, oracle):
    _compare(_lib("emul"), oracle, _cases(oracle)[case], iters=2)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["ragged", "imu_only", "camera_only", "empty_sensor", "half_outliers"])
def test_edge_cases_gpu(case, oracle, product_lib):
    _compare(_lib("gpu", product_lib), oracle, _cases(oracle)[case], iters=8)


def test_assembly_errors_emulated(oracle):
    """Statuses of problem assembly: unknown rigid body -> kFailedPrecondition (camera.cpp:125-131), wrong intrinsics size ->
    kInvalidArgument (camera.cpp:26-33), stamp outside the valid knots -> kInvalidArgument, no trajectory -> kFailedPrecondition;
    an all-outlier problem has nothing to optimise and converges immediately with zero residual blocks."""
    lib = _lib("emul")
    truth, prob = synthetic.generate("micro", oracle.oracle_api, noise=True)
    cam = next(s for s in prob.sensors if s.kind == synthetic.CAMERA)
    p = prob.clone(); c = next(s for s in p.sensors if s.kind == synthetic.CAMERA); c.model_id = np.asarray(c.model_id) + 7
    a = _capi.CApi(lib); p.push(a)
    with pytest.raises(_capi.CalicoError) as e:
        a.optimize(_capi.Options(minimizer_progress_to_stdout=0))
    assert e.value.code == _capi.FAILED_PRECONDITION
    p = prob.clone(); c = next(s for s in p.sensors if s.kind == synthetic.CAMERA); c.intr = np.asarray(c.intr)[:-1]
    a = _capi.CApi(lib)
    with pytest.raises(_capi.CalicoError) as e:
        p.push(a); a.optimize(_capi.Options(minimizer_progress_to_stdout=0))
    assert e.value.code == _capi.INVALID_ARGUMENT
    p = prob.clone(); g = next(s for s in p.sensors if s.kind == synthetic.GYROSCOPE); g.stamp = np.asarray(g.stamp) + 1e3
    a = _capi.CApi(lib); p.push(a)
    with pytest.raises(_capi.CalicoError) as e:
        a.optimize(_capi.Options(minimizer_progress_to_stdout=0))
    assert e.value.code == _capi.INVALID_ARGUMENT
    a = _capi.CApi(lib)
    with pytest.raises(_capi.CalicoError) as e:
        a.optimize(_capi.Options(minimizer_progress_to_stdout=0))
    assert e.value.code == _capi.FAILED_PRECONDITION
    p = prob.clone(); p.sensors = [cam.__class__(**{**cam.__dict__})]; p.sensors[0].outlier = np.ones(cam.n_obs, dtype=np.uint8)
    a = _capi.CApi(lib); p.push(a)
    s, log = a.optimize(_capi.Options(minimizer_progress_to_stdout=0))
    assert s.termination_type == _capi.CONVERGENCE and s.num_residual_blocks == 0
