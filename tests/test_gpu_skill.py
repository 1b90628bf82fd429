"""GPU parity of the skill reductions (RPS / RPSS / CC / ACC / ensemble mean / MME combine)."""
import ctypes as C

import numpy as np
import pytest

from oracle import skill as so

pytestmark = pytest.mark.gpu


def dev(arr, stream, dtype=np.float32):
    from s2s_ismr_unet_b200.runtime import DeviceBuffer
    return DeviceBuffer.from_array(np.ascontiguousarray(arr, dtype), stream)


def call(name, *a):
    from s2s_ismr_unet_b200._lib import call as c
    c(name, *a)


def P(b):
    return C.c_void_p(b.ptr)


@pytest.mark.parametrize("T,Y,X", [(261, 64, 64), (21, 24, 24), (7, 5, 3), (300, 33, 31)])
def test_rps_and_rpss(T, Y, X, stream):
    from s2s_ismr_unet_b200.runtime import DeviceBuffer
    rng = np.random.default_rng(0)
    p = rng.dirichlet([1, 1, 1], size=(T, Y, X)).astype(np.float32)
    lab = rng.integers(0, 3, size=(T, Y, X)).astype(np.float64)
    lab[rng.random((T, Y, X)) < 0.05] = np.nan
    lab[:, 0, 0] = np.nan                                  # an all-missing gridpoint -> NaN
    o = so.onehot_obs(lab).astype(np.float32)
    clim = so.climo_forecast((T, Y, X)).astype(np.float32)
    d_p, d_o, d_c, d_out = dev(p, stream), dev(o, stream), dev(clim, stream), DeviceBuffer(4 * Y * X)
    sp = C.c_void_p(stream.ptr)
    call("s2s_rps_map", P(d_p), P(d_o), T, Y, X, P(d_out), sp)
    got = d_out.download((Y, X), np.float32, stream)
    ref = so.rps(o, p)
    assert np.isnan(got[0, 0]) and np.isnan(ref[0, 0])
    np.testing.assert_allclose(got, ref, atol=1e-5, equal_nan=True)
    call("s2s_rpss_map", P(d_p), P(d_c), P(d_o), T, Y, X, P(d_out), sp)
    got = d_out.download((Y, X), np.float32, stream)
    np.testing.assert_allclose(got, so.rpss(clim, p, o), atol=1e-4, equal_nan=True)


def test_rps_known_answers(stream):
    from s2s_ismr_unet_b200.runtime import DeviceBuffer
    T, Y, X = 9, 2, 3
    lab = np.tile(np.array([0, 1, 2] * 3, np.float64)[:, None, None], (1, Y, X))
    o = so.onehot_obs(lab).astype(np.float32)
    clim = so.climo_forecast((T, Y, X)).astype(np.float32)
    d_o, d_c, d_out = dev(o, stream), dev(clim, stream), DeviceBuffer(4 * Y * X)
    sp = C.c_void_p(stream.ptr)
    call("s2s_rps_map", P(d_o), P(d_o), T, Y, X, P(d_out), sp)        # perfect forecast
    np.testing.assert_allclose(d_out.download((Y, X), np.float32, stream), 0.0, atol=1e-7)
    call("s2s_rps_map", P(d_c), P(d_o), T, Y, X, P(d_out), sp)        # climatology: (5/9 + 2/9 + 5/9)/3
    np.testing.assert_allclose(d_out.download((Y, X), np.float32, stream), 4.0 / 9.0, atol=1e-6)
    call("s2s_rpss_map", P(d_c), P(d_c), P(d_o), T, Y, X, P(d_out), sp)  # climo vs climo -> 0
    np.testing.assert_allclose(d_out.download((Y, X), np.float32, stream), 0.0, atol=1e-6)


# (128, 256, 256): BASELINE.json configs[4], the 0.25-degree archive grid, with NaN pairs
@pytest.mark.parametrize("T,Y,X", [(261, 64, 64), (44, 24, 24), (500, 17, 9), (128, 256, 256)])
def test_acc_and_cc(T, Y, X, stream):
    from s2s_ismr_unet_b200.runtime import DeviceBuffer
    rng = np.random.default_rng(1)
    week = rng.integers(18, 40, size=T)
    x = rng.gamma(2.0, 3.0, size=(T, Y, X)).astype(np.float32)
    y = (0.5 * x + 0.5 * rng.gamma(2.0, 3.0, size=(T, Y, X))).astype(np.float32)
    x[rng.random((T, Y, X)) < 0.02] = np.nan
    y[rng.random((T, Y, X)) < 0.02] = np.nan
    acc_ref, cc_ref = so.acc_cc(x, y, week)
    weeks, gid = np.unique(week, return_inverse=True)
    order = np.argsort(gid, kind="stable").astype(np.int32)
    gstart = np.concatenate([[0], np.cumsum(np.bincount(gid, minlength=len(weeks)))]).astype(np.int32)
    d_x, d_y = dev(x, stream), dev(y, stream)
    d_o, d_g = dev(order, stream, np.int32), dev(gstart, stream, np.int32)
    d_acc, d_cc = DeviceBuffer(4 * Y * X), DeviceBuffer(4 * Y * X)
    call("s2s_acc_map", P(d_x), P(d_y), P(d_o), P(d_g), len(weeks), T, Y, X, P(d_acc), P(d_cc), C.c_void_p(stream.ptr))
    np.testing.assert_allclose(d_acc.download((Y, X), np.float32, stream), acc_ref, atol=1e-4)   # BASELINE: ACC <= 1e-4 abs
    np.testing.assert_allclose(d_cc.download((Y, X), np.float32, stream), cc_ref, atol=1e-4)


def test_ensemble_mean_and_mme_combine(stream):
    from s2s_ismr_unet_b200.runtime import DeviceBuffer
    rng = np.random.default_rng(2)
    T, M, Y, X = 13, 11, 9, 7
    x = rng.normal(size=(T, M, Y, X)).astype(np.float32)
    x[rng.random(x.shape) < 0.1] = np.nan
    d_x, d_out = dev(x, stream), DeviceBuffer(4 * T * Y * X)
    sp = C.c_void_p(stream.ptr)
    call("s2s_ensemble_mean", P(d_x), T, M, Y, X, P(d_out), sp)
    np.testing.assert_allclose(d_out.download((T, Y, X), np.float32, stream), so.ensemble_mean(x), atol=1e-6, equal_nan=True)
    probs = rng.dirichlet([1, 1, 1], size=(3, T, Y, X)).astype(np.float32)
    d_p, d_o = dev(probs, stream), DeviceBuffer(4 * T * Y * X * 3)
    call("s2s_mme_combine", P(d_p), 3, C.c_int64(T * Y * X), P(d_o), sp)
    np.testing.assert_allclose(d_o.download((T, Y, X, 3), np.float32, stream), so.mme_combine(list(probs)), atol=1e-6)
