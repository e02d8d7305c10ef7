"""Trajectory spline fit: oracle pinned on the reference's acceptance test, the kernels through the SIMT emulation
on CPU, and the CUDA path through the C ABI on a GPU."""
import os
import sys

import numpy as np
import pytest

from calico_b200 import _capi

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emul"))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import spline_fit as ofit  # noqa: E402


def _fixture():
    """calico/test/bspline_test.cpp:14-31: 10 s sampled at 0.01 s of (cos t, sin 1.5 t, t cos t); order 6, 5 Hz knots."""
    t = np.arange(0.0, 10.0, 0.01)
    data = np.stack([np.cos(t), np.sin(1.5 * t), t * np.cos(t)], axis=1)
    return t, data


def _expected(t, d):
    return [np.stack([np.cos(t), np.sin(1.5 * t), t * np.cos(t)], 1),
            np.stack([-np.sin(t), 1.5 * np.cos(1.5 * t), np.cos(t) - t * np.sin(t)], 1),
            np.stack([-np.cos(t), -2.25 * np.sin(1.5 * t), -2.0 * np.sin(t) - t * np.cos(t)], 1),
            np.stack([np.sin(t), -3.375 * np.cos(1.5 * t), t * np.sin(t) - 3.0 * np.cos(t)], 1)][d]


def test_oracle_fit_meets_reference_acceptance():
    """bspline_test.cpp:52-94 InterpolationPrecision3DOF on the restated dense fit."""
    t, data = _fixture()
    knots, valid, basis, ctrl = ofit.fit_spline(t, data, 6, 5.0)
    ti = (t[-1] - t[0]) / 201 * np.arange(201)
    for d, tol in enumerate([1e-6, 1e-5, 1e-4, 1e-2]):
        got = ofit.interpolate(knots, valid, basis, ctrl, ti, d, 6)
        assert np.abs(got - _expected(ti, d)).max() < tol


def test_oracle_fit_rejects_bad_input():
    """bspline.hpp:299-327."""
    t, data = _fixture()
    for args in [(t[:0], data[:0], 6, 5.0), (t, data[:-1], 6, 5.0), (t[::-1], data, 6, 5.0), (t, data, 1, 5.0), (t, data, 6, 0.0)]:
        with pytest.raises(ValueError):
            ofit.fit_spline(*args)


def _check_against_oracle(lib_path, n=600, freq=7.0, seed=3):
    rng = np.random.default_rng(seed)
    t = np.sort(rng.uniform(0.0, 6.0, n))
    t[0], t[-1] = 0.0, 6.0
    data = np.stack([np.sin((0.5 + 0.3 * c) * t + c) + 0.01 * rng.standard_normal(n) for c in range(6)], axis=1)
    knots, ctrl = _capi.fit_spline(t, data, 6, freq, lib_path=lib_path)
    ok, _, _, oc = ofit.fit_spline(t, data, 6, freq)
    np.testing.assert_allclose(knots, ok, rtol=0, atol=1e-12)
    np.testing.assert_allclose(ctrl, oc, rtol=1e-7, atol=1e-7)
    # trajectory entry point: shuffled stamps, rotations through +-pi (phase unwrapping)
    ang = np.linspace(0.0, 3.0 * np.pi, n)[:, None] * np.array([[0.0, 0.0, 1.0]]) + 0.2 * np.sin(t)[:, None] * np.array([[1.0, 0.0, 0.0]])
    th = np.linalg.norm(ang, axis=1)
    q = np.concatenate([ang * (np.sin(0.5 * th) / np.where(th > 0, th, 1.0))[:, None], np.cos(0.5 * th)[:, None]], axis=1)
    pos = np.stack([np.cos(t), np.sin(t), 0.1 * t], 1)
    perm = rng.permutation(n)
    k2, c2 = _capi.fit_trajectory(t[perm], q[perm], pos[perm], freq, 6, lib_path=lib_path)
    _, _, _, oc2 = ofit.fit_trajectory(t[perm], q[perm], pos[perm], freq, 6)
    np.testing.assert_allclose(c2, oc2, rtol=1e-7, atol=1e-7)
    assert np.abs(np.diff(c2[5:-5, 2])).max() < 1.0   # unwrapped: no 2 pi jumps along the (interior) control polygon
    # error behaviour mirrors CheckDataForSplineFit
    for bad in [(t[::-1], data, 6, freq), (t, data, 1, freq), (t, data, 6, 0.0)]:
        with pytest.raises(_capi.CalicoError) as e:
            _capi.fit_spline(*bad, lib_path=lib_path)
        assert e.value.code == _capi.INVALID_ARGUMENT


@pytest.mark.timeout(600)
def test_emulated_fit_matches_oracle():
    import build as emul_build
    _check_against_oracle(emul_build.build())


@pytest.mark.gpu
def test_gpu_fit_matches_oracle(product_lib):
    _check_against_oracle(product_lib)


@pytest.mark.gpu
def test_gpu_fit_meets_reference_acceptance_and_scales(product_lib):
    """bspline_test.cpp:52-94 through the CUDA path, then a C5-sized fit (10 000 poses, 5 005 control points) checked by its residual
    (size-independent property: samples of a spline of the same knot vector are reproduced)."""
    t, data = _fixture()
    data6 = np.concatenate([data, data], axis=1)
    knots, ctrl = _capi.fit_spline(t, data6, 6, 5.0, lib_path=product_lib)
    kn, valid = ofit.knot_vector(t[0], t[-1], 5.0, 6)
    basis = [ofit.basis_matrix(kn, 6, i + 5) for i in range(valid.size - 1)]
    ti = (t[-1] - t[0]) / 201 * np.arange(201)
    for d, tol in enumerate([1e-6, 1e-5, 1e-4, 1e-2]):
        got = ofit.interpolate(knots, valid, basis, ctrl, ti, d, 6)[:, :3]
        assert np.abs(got - _expected(ti, d)).max() < tol
    n = 10000
    tt = np.arange(n) / 20.0
    truth_k, truth_c = _capi.fit_spline(tt, np.stack([np.sin(0.3 * tt + c) for c in range(6)], 1), 6, 10.0, lib_path=product_lib)
    assert truth_c.shape == (5005, 6)
    from calico_b200 import spline as sp
    s = sp.Spline(6, truth_k, truth_c)
    samples = s.evaluate(tt, 0)
    k2, c2 = _capi.fit_spline(tt, samples, 6, 10.0, lib_path=product_lib)
    np.testing.assert_allclose(sp.Spline(6, k2, c2).evaluate(tt, 0), samples, rtol=0, atol=1e-7)   # cond(X^T X) ~ 1e6 at the ends of the span
