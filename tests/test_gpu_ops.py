"""GPU parity of every single-operator C-ABI entry point against the torch-CPU oracle (fp64).
Tolerance: rel-L2 <= 1e-5 (BASELINE.json: fp32 forward outputs), tighter where possible."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import rel_l2
from oracle import keras_unet as ko

pytestmark = pytest.mark.gpu
TOL = 1e-5


def dev(arr, stream):
    from s2s_ismr_unet_b200.runtime import DeviceBuffer
    return DeviceBuffer.from_array(np.ascontiguousarray(arr, np.float32), stream)


def empty(n, stream):
    from s2s_ismr_unet_b200.runtime import DeviceBuffer
    return DeviceBuffer(4 * n)


def P(b):
    return C.c_void_p(b.ptr)


def call(name, *a):
    from s2s_ismr_unet_b200._lib import call as c
    c(name, *a)


def nchw(x):
    return torch.tensor(x, dtype=torch.float64).permute(0, 3, 1, 2)


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous().numpy()


CONV_SHAPES = [  # N, H, W, Cin, Cout
    (2, 64, 64, 1, 8), (2, 64, 64, 3, 8), (16, 64, 64, 8, 8), (2, 32, 32, 8, 16), (2, 32, 32, 16, 16),
    (2, 16, 16, 16, 32), (3, 8, 8, 32, 64), (16, 8, 8, 64, 64), (2, 24, 24, 1, 12), (2, 24, 24, 12, 12),
    (1, 6, 6, 24, 48), (2, 3, 3, 48, 96), (2, 4, 4, 128, 128), (1, 2, 2, 128, 256), (2, 16, 16, 64, 32),
    (1, 256, 256, 3, 8), (5, 12, 20, 24, 24),
    # throughput regime (>= 222 tiles of 16x32 -> the 16x32-pixel tile instantiation of gconv_kernel), incl. ragged
    # edges, a 12-channel contraction (chunks 8 + 4) and a 12-channel output (groups 8 + 4)
    (32, 64, 64, 8, 8), (28, 64, 64, 16, 8), (30, 64, 64, 12, 12), (2, 256, 256, 8, 16), (8, 88, 150, 4, 8),
]


@pytest.mark.parametrize("shape", CONV_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_conv3x3_forward(shape, stream):
    N, H, W, Cin, Cout = shape
    rng = np.random.default_rng(1)
    x = rng.normal(size=(N, H, W, Cin)).astype(np.float32)
    w = (rng.normal(size=(3, 3, Cin, Cout)) / np.sqrt(9 * Cin)).astype(np.float32)
    b = rng.normal(size=(Cout,)).astype(np.float32) * 0.1
    ref = nhwc(ko.elu(ko.conv3x3_same(nchw(x), torch.tensor(w, dtype=torch.float64), torch.tensor(b, dtype=torch.float64))))
    dx, dw, db, dy = dev(x, stream), dev(w, stream), dev(b, stream), empty(N * H * W * Cout, stream)
    call("s2s_op_conv3x3_fwd", P(dx), P(dw), P(db), P(dy), N, H, W, Cin, Cout, 1, C.c_void_p(stream.ptr))
    got = dy.download((N, H, W, Cout), np.float32, stream)
    assert rel_l2(got, ref) <= TOL, f"conv fwd {shape}: rel-L2 {rel_l2(got, ref):.3e}"


@pytest.mark.parametrize("shape", [s for s in CONV_SHAPES if s[3] % 4 == 0], ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("with_act", [False, True])
def test_conv3x3_dgrad(shape, with_act, stream):
    N, H, W, Cin, Cout = shape
    rng = np.random.default_rng(2)
    dz = rng.normal(size=(N, H, W, Cout)).astype(np.float32)
    w = (rng.normal(size=(3, 3, Cin, Cout)) / np.sqrt(9 * Cin)).astype(np.float32)
    act = ko.elu(torch.tensor(rng.normal(size=(N, H, W, Cin)), dtype=torch.float64)).numpy().astype(np.float32)
    xt = torch.zeros((N, Cin, H, W), dtype=torch.float64, requires_grad=True)
    y = ko.conv3x3_same(xt, torch.tensor(w, dtype=torch.float64), torch.zeros(Cout, dtype=torch.float64))
    (g,) = torch.autograd.grad(y, xt, nchw(dz))
    ref = nhwc(g)
    if with_act:
        a64 = act.astype(np.float64)
        ref = ref * np.where(a64 > 0, 1.0, a64 + 1.0)
    d_dz, d_w, d_act, d_dx = dev(dz, stream), dev(w, stream), dev(act, stream), empty(N * H * W * Cin, stream)
    call("s2s_op_conv3x3_dgrad", P(d_dz), P(d_w), P(d_act) if with_act else None, P(d_dx), N, H, W, Cin, Cout, C.c_void_p(stream.ptr))
    got = d_dx.download((N, H, W, Cin), np.float32, stream)
    assert rel_l2(got, ref) <= TOL, f"conv dgrad {shape} act={with_act}: rel-L2 {rel_l2(got, ref):.3e}"


@pytest.mark.parametrize("shape", CONV_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_conv3x3_wgrad(shape, stream):
    N, H, W, Cin, Cout = shape
    rng = np.random.default_rng(3)
    x = rng.normal(size=(N, H, W, Cin)).astype(np.float32)
    dz = rng.normal(size=(N, H, W, Cout)).astype(np.float32)
    wt = torch.zeros((3, 3, Cin, Cout), dtype=torch.float64, requires_grad=True)
    bt = torch.zeros(Cout, dtype=torch.float64, requires_grad=True)
    y = ko.conv3x3_same(nchw(x), wt, bt)
    gw, gb = torch.autograd.grad(y, (wt, bt), nchw(dz))
    d_x, d_dz, d_dw, d_db = dev(x, stream), dev(dz, stream), empty(9 * Cin * Cout, stream), empty(Cout, stream)
    call("s2s_op_conv3x3_wgrad", P(d_x), P(d_dz), P(d_dw), P(d_db), N, H, W, Cin, Cout, C.c_void_p(stream.ptr))
    got_w = d_dw.download((3, 3, Cin, Cout), np.float32, stream)
    got_b = d_db.download((Cout,), np.float32, stream)
    assert rel_l2(got_w, gw.numpy()) <= TOL, f"wgrad {shape}: rel-L2 {rel_l2(got_w, gw.numpy()):.3e}"
    assert rel_l2(got_b, gb.numpy()) <= TOL, f"bgrad {shape}: rel-L2 {rel_l2(got_b, gb.numpy()):.3e}"


CT_SHAPES = [  # N, h, w, Cin, Cout
    (2, 8, 8, 64, 32), (2, 16, 16, 32, 16), (2, 32, 32, 16, 8), (1, 3, 3, 48, 24), (16, 8, 8, 64, 32),
    (2, 2, 2, 256, 128), (1, 128, 128, 16, 8), (3, 6, 10, 24, 12),
]


@pytest.mark.parametrize("k", [2, 3, 5])
@pytest.mark.parametrize("shape", CT_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_conv_transpose(shape, k, stream):
    N, h, w, Cin, Cout = shape
    rng = np.random.default_rng(4)
    x = rng.normal(size=(N, h, w, Cin)).astype(np.float32)
    wt = (rng.normal(size=(k, k, Cout, Cin)) / np.sqrt(k * k * Cin / 4)).astype(np.float32)
    b = rng.normal(size=(Cout,)).astype(np.float32) * 0.1
    dy = rng.normal(size=(N, 2 * h, 2 * w, Cout)).astype(np.float32)
    xt = torch.tensor(x, dtype=torch.float64).permute(0, 3, 1, 2).requires_grad_(True)
    wtt = torch.tensor(wt, dtype=torch.float64, requires_grad=True)
    bt = torch.tensor(b, dtype=torch.float64, requires_grad=True)
    y = ko.conv_transpose_same_s2(xt, wtt, bt)
    gx, gw, gb = torch.autograd.grad(y, (xt, wtt, bt), nchw(dy))
    sp = C.c_void_p(stream.ptr)
    d_x, d_w, d_b, d_dy = dev(x, stream), dev(wt, stream), dev(b, stream), dev(dy, stream)
    d_y, d_dx, d_dw, d_db = empty(dy.size, stream), empty(x.size, stream), empty(wt.size, stream), empty(Cout, stream)
    call("s2s_op_convt_fwd", P(d_x), P(d_w), P(d_b), P(d_y), N, h, w, Cin, Cout, k, sp)
    got = d_y.download(dy.shape, np.float32, stream)
    assert rel_l2(got, nhwc(y.detach())) <= TOL, f"convT fwd {shape} k={k}: {rel_l2(got, nhwc(y.detach())):.3e}"
    call("s2s_op_convt_dgrad", P(d_dy), P(d_w), P(d_dx), N, h, w, Cin, Cout, k, sp)
    got = d_dx.download(x.shape, np.float32, stream)
    assert rel_l2(got, nhwc(gx)) <= TOL, f"convT dgrad {shape} k={k}: {rel_l2(got, nhwc(gx)):.3e}"
    call("s2s_op_convt_wgrad", P(d_x), P(d_dy), P(d_dw), P(d_db), N, h, w, Cin, Cout, k, sp)
    got_w = d_dw.download(wt.shape, np.float32, stream)
    got_b = d_db.download((Cout,), np.float32, stream)
    assert rel_l2(got_w, gw.numpy()) <= TOL, f"convT wgrad {shape} k={k}: {rel_l2(got_w, gw.numpy()):.3e}"
    assert rel_l2(got_b, gb.numpy()) <= TOL, f"convT bgrad {shape} k={k}: {rel_l2(got_b, gb.numpy()):.3e}"


def test_adam_keras_form(stream):
    from s2s_ismr_unet_b200._lib import AdamCfg
    rng = np.random.default_rng(5)
    n = 10007
    p = rng.normal(size=n).astype(np.float32)
    m = np.zeros(n, np.float32)
    v = np.zeros(n, np.float32)
    d_p, d_m, d_v = dev(p, stream), dev(m, stream), dev(v, stream)
    cfg = AdamCfg(1e-3, 0.9, 0.999, 1e-7)
    p64, m64, v64 = p.astype(np.float64), m.astype(np.float64), v.astype(np.float64)
    for t in range(1, 6):
        g = rng.normal(size=n).astype(np.float32)
        d_g = dev(g, stream)
        call("s2s_adam_step", P(d_p), P(d_g), P(d_m), P(d_v), C.c_size_t(n), C.byref(cfg), C.c_int64(t), C.c_void_p(stream.ptr))
        alpha = 1e-3 * np.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t)
        m64 += (g - m64) * (1 - 0.9)
        v64 += (g.astype(np.float64) ** 2 - v64) * (1 - 0.999)
        p64 -= alpha * m64 / (np.sqrt(v64) + 1e-7)
    got = d_p.download((n,), np.float32, stream)
    assert np.max(np.abs(got - p64)) <= 2e-6, f"adam: max abs err {np.max(np.abs(got - p64)):.3e}"


TC_SHAPES = [(2, 16, 16, 32, 32), (2, 8, 8, 96, 48), (2, 16, 16, 64, 32), (16, 8, 8, 64, 64), (1, 32, 24, 128, 64), (3, 20, 12, 64, 16), (2, 8, 8, 192, 96), (1, 4, 4, 256, 256)]


@pytest.mark.parametrize("shape", TC_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_conv3x3_forward_tensor_cores(shape, stream):
    """tcgen05/TMEM/TMA implicit-GEMM conv (bf16 operands, fp32 accumulate).  Checked (a) against the oracle on
    the SAME bf16-rounded operands (only fp32 accumulation order differs: rel-L2 <= 1e-5) and (b) against the
    full-precision oracle within BASELINE's bf16 tolerance (rel-L2 <= 1e-2)."""
    N, H, W, Cin, Cout = shape
    rng = np.random.default_rng(6)
    x = rng.normal(size=(N, H, W, Cin)).astype(np.float32)
    w = (rng.normal(size=(3, 3, Cin, Cout)) / np.sqrt(9 * Cin)).astype(np.float32)
    b = rng.normal(size=(Cout,)).astype(np.float32) * 0.1
    t64 = lambda a: torch.tensor(a, dtype=torch.float64)
    ref = nhwc(ko.elu(ko.conv3x3_same(nchw(x), t64(w), t64(b))))
    bf = lambda a: torch.tensor(a).to(torch.bfloat16).to(torch.float64)
    ref_bf = nhwc(ko.elu(ko.conv3x3_same(bf(x).permute(0, 3, 1, 2), bf(w), t64(b))))
    dx, dw, db, dy = dev(x, stream), dev(w, stream), dev(b, stream), empty(N * H * W * Cout, stream)
    call("s2s_op_conv3x3_fwd_tc", P(dx), P(dw), P(db), P(dy), N, H, W, Cin, Cout, 1, C.c_void_p(stream.ptr))
    got = dy.download((N, H, W, Cout), np.float32, stream)
    assert rel_l2(got, ref_bf) <= 1e-5, f"tc conv {shape} vs bf16-rounded oracle: rel-L2 {rel_l2(got, ref_bf):.3e}"
    assert rel_l2(got, ref) <= 1e-2, f"tc conv {shape}: rel-L2 {rel_l2(got, ref):.3e}"


# tcgen05 tf32 implicit GEMM on the fp32 tensors (csrc/tc3conv.cuh): N, H, W, Cin, Cout.  Ragged grids (24x24, 20x12),
# one-tile images (8x8), a channel count that is no multiple of 16 (24 -> 12), N-chunked (48 -> 96) and multi-stage (192).
TF32_SHAPES = [(16, 64, 64, 8, 8), (2, 32, 32, 8, 16), (2, 32, 32, 16, 16), (3, 16, 16, 32, 32), (3, 16, 16, 64, 32),
               (2, 24, 24, 32, 16), (2, 16, 16, 24, 12), (2, 16, 16, 48, 96), (3, 20, 12, 16, 8), (4, 8, 8, 96, 96),
               (1, 16, 16, 192, 192), (1, 256, 256, 8, 8),
               # flat geometry (whole zero-padded small images per tile): 2x2 .. 6x6, 1x1, ragged batches, chunked channels
               (16, 2, 2, 192, 384), (16, 4, 4, 96, 192), (5, 4, 4, 32, 32), (3, 6, 6, 16, 16), (16, 1, 1, 64, 64), (7, 3, 3, 24, 12),
               # contracted channel count that is no multiple of 8 (padded chunk, TMA zero fill): the f=3 grid points
               (2, 64, 64, 12, 12), (2, 16, 16, 20, 12)]


@pytest.mark.parametrize("npass", [1, 3])
@pytest.mark.parametrize("shape", TF32_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_conv3x3_forward_tf32_tensor_cores(shape, npass, stream):
    """npass = 3 (error-compensated 3xTF32 split) must meet the fp32 bar (rel-L2 <= 1e-5 against the fp64 oracle);
    npass = 1 (single tf32 pass) the reduced-precision bar of BASELINE.json (<= 1e-2; measured ~4e-4)."""
    N, H, W, Cin, Cout = shape
    rng = np.random.default_rng(7)
    x = rng.normal(size=(N, H, W, Cin)).astype(np.float32)
    w = (rng.normal(size=(3, 3, Cin, Cout)) / np.sqrt(9 * Cin)).astype(np.float32)
    b = rng.normal(size=(Cout,)).astype(np.float32) * 0.1
    ref = nhwc(ko.elu(ko.conv3x3_same(nchw(x), torch.tensor(w, dtype=torch.float64), torch.tensor(b, dtype=torch.float64))))
    dx, dw, db, dy = dev(x, stream), dev(w, stream), dev(b, stream), empty(N * H * W * Cout, stream)
    call("s2s_op_conv3x3_fwd_tf32", P(dx), P(dw), P(db), P(dy), N, H, W, Cin, Cout, 1, npass, C.c_void_p(stream.ptr))
    got = dy.download((N, H, W, Cout), np.float32, stream)
    tol = 1e-5 if npass == 3 else 2e-3
    assert rel_l2(got, ref) <= tol, f"tf32 x{npass} conv fwd {shape}: rel-L2 {rel_l2(got, ref):.3e}"


@pytest.mark.parametrize("npass", [1, 3])
@pytest.mark.parametrize("with_act", [False, True])
@pytest.mark.parametrize("shape", TF32_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_conv3x3_dgrad_tf32_tensor_cores(shape, with_act, npass, stream):
    N, H, W, Cin, Cout = shape
    rng = np.random.default_rng(8)
    dz = rng.normal(size=(N, H, W, Cout)).astype(np.float32)
    w = (rng.normal(size=(3, 3, Cin, Cout)) / np.sqrt(9 * Cin)).astype(np.float32)
    act = ko.elu(torch.tensor(rng.normal(size=(N, H, W, Cin)), dtype=torch.float64)).numpy().astype(np.float32)
    xt = torch.zeros((N, Cin, H, W), dtype=torch.float64, requires_grad=True)
    y = ko.conv3x3_same(xt, torch.tensor(w, dtype=torch.float64), torch.zeros(Cout, dtype=torch.float64))
    (g,) = torch.autograd.grad(y, xt, nchw(dz))
    ref = nhwc(g)
    if with_act:
        a64 = act.astype(np.float64)
        ref = ref * np.where(a64 > 0, 1.0, a64 + 1.0)
    d_dz, d_w, d_act, d_dx = dev(dz, stream), dev(w, stream), dev(act, stream), empty(N * H * W * Cin, stream)
    call("s2s_op_conv3x3_dgrad_tf32", P(d_dz), P(d_w), P(d_act) if with_act else None, P(d_dx), N, H, W, Cin, Cout, npass,
         C.c_void_p(stream.ptr))
    got = d_dx.download((N, H, W, Cin), np.float32, stream)
    tol = 1e-5 if npass == 3 else 2e-3
    assert rel_l2(got, ref) <= tol, f"tf32 x{npass} conv dgrad {shape} act={with_act}: rel-L2 {rel_l2(got, ref):.3e}"


# tcgen05 tf32 weight gradient (csrc/tcwgrad.cuh): N, H, W, Cin, Cout, n_max.  Tiled geometry (H*W >= 64: ragged 24x24 and
# 20x12 grids, one-tile 8x8 images), flat geometry (whole zero-padded small images per tile: 6x6, 4x4, 3x3, 2x2, 1x1), input /
# output channel chunking (96, 192, 384), channel counts that are no multiple of 32 (12, 24, 48), and a batch smaller than
# the one the plan was made for (the reference's last batch of 5: images 5.. of the buffers must not contribute).
WGRAD_TF32_SHAPES = [(16, 64, 64, 8, 8, 16), (2, 32, 32, 16, 8, 2), (2, 24, 24, 12, 24, 2), (3, 20, 12, 16, 8, 3), (3, 16, 16, 32, 32, 3),
                     (2, 8, 8, 64, 32, 2), (2, 16, 16, 96, 96, 2), (1, 16, 16, 48, 192, 1), (4, 4, 4, 32, 64, 4), (5, 4, 4, 64, 64, 8),
                     (16, 2, 2, 96, 192, 16), (5, 2, 2, 32, 32, 16), (3, 3, 3, 16, 16, 3), (16, 1, 1, 64, 64, 16), (3, 6, 6, 16, 16, 3),
                     (5, 64, 64, 8, 8, 16), (16, 2, 2, 384, 384, 16)]


@pytest.mark.parametrize("shape", WGRAD_TF32_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_conv3x3_wgrad_tf32_tensor_cores(shape, stream):
    """Single tf32 pass: the reduced-precision bar of BASELINE.json (<= 1e-2; measured ~7e-4) against the fp64 oracle."""
    N, H, W, Cin, Cout, n_max = shape
    rng = np.random.default_rng(11)
    x = rng.normal(size=(n_max, H, W, Cin)).astype(np.float32)
    dz = rng.normal(size=(n_max, H, W, Cout)).astype(np.float32)
    wt = torch.zeros((3, 3, Cin, Cout), dtype=torch.float64, requires_grad=True)
    bt = torch.zeros(Cout, dtype=torch.float64, requires_grad=True)
    y = ko.conv3x3_same(nchw(x[:N]), wt, bt)
    gw, gb = torch.autograd.grad(y, (wt, bt), nchw(dz[:N]))
    d_x, d_dz, d_dw, d_db = dev(x, stream), dev(dz, stream), empty(9 * Cin * Cout, stream), empty(Cout, stream)
    call("s2s_op_conv3x3_wgrad_tf32", P(d_x), P(d_dz), P(d_dw), P(d_db), N, H, W, Cin, Cout, n_max, C.c_void_p(stream.ptr))
    got_w = d_dw.download((3, 3, Cin, Cout), np.float32, stream)
    got_b = d_db.download((Cout,), np.float32, stream)
    assert rel_l2(got_w, gw.numpy()) <= 2e-3, f"tf32 wgrad {shape}: rel-L2 {rel_l2(got_w, gw.numpy()):.3e}"
    assert rel_l2(got_b, gb.numpy()) <= 2e-3, f"tf32 bgrad {shape}: rel-L2 {rel_l2(got_b, gb.numpy()):.3e}"


# Conv2DTranspose on the tcgen05 tensor cores: N, h, w, Cin, Cout.  Tiled and flat (small-image) geometries, chunked channels.
CONVT_TF32_SHAPES = [(2, 16, 16, 16, 8), (3, 8, 8, 64, 32), (2, 32, 32, 16, 8), (5, 4, 4, 32, 16), (16, 2, 2, 96, 48), (4, 12, 12, 24, 16),
                     (2, 8, 8, 384, 192), (16, 1, 1, 64, 32), (2, 32, 32, 24, 12)]


@pytest.mark.parametrize("k", [2, 3, 5])
@pytest.mark.parametrize("shape", CONVT_TF32_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_conv_transpose_tf32_tensor_cores(shape, k, stream):
    """Single tf32 pass: the reduced-precision bar of BASELINE.json (<= 1e-2; measured ~5e-4) against the fp64 oracle."""
    N, h, w, Cin, Cout = shape
    rng = np.random.default_rng(14)
    x = rng.normal(size=(N, h, w, Cin)).astype(np.float32)
    wt = (rng.normal(size=(k, k, Cout, Cin)) / np.sqrt(k * k * Cin / 4)).astype(np.float32)
    b = rng.normal(size=(Cout,)).astype(np.float32) * 0.1
    dy = rng.normal(size=(N, 2 * h, 2 * w, Cout)).astype(np.float32)
    xt = torch.tensor(x, dtype=torch.float64).permute(0, 3, 1, 2).requires_grad_(True)
    wtt = torch.tensor(wt, dtype=torch.float64)
    y = ko.conv_transpose_same_s2(xt, wtt, torch.tensor(b, dtype=torch.float64))
    (gx,) = torch.autograd.grad(y, xt, nchw(dy))
    sp = C.c_void_p(stream.ptr)
    d_x, d_w, d_b, d_dy = dev(x, stream), dev(wt, stream), dev(b, stream), dev(dy, stream)
    d_y, d_dx = empty(dy.size, stream), empty(x.size, stream)
    call("s2s_op_convt_fwd_tf32", P(d_x), P(d_w), P(d_b), P(d_y), N, h, w, Cin, Cout, k, sp)
    got = d_y.download(dy.shape, np.float32, stream)
    assert rel_l2(got, nhwc(y.detach())) <= 2e-3, f"tf32 convT fwd {shape} k={k}: {rel_l2(got, nhwc(y.detach())):.3e}"
    call("s2s_op_convt_dgrad_tf32", P(d_dy), P(d_w), P(d_dx), N, h, w, Cin, Cout, k, sp)
    got = d_dx.download(x.shape, np.float32, stream)
    assert rel_l2(got, nhwc(gx)) <= 2e-3, f"tf32 convT dgrad {shape} k={k}: {rel_l2(got, nhwc(gx)):.3e}"
    # weight gradient: pixel-contraction GEMM over the four parity planes of dy (csrc/tcwgrad.cuh)
    wtg = torch.tensor(wt, dtype=torch.float64, requires_grad=True)
    y2 = ko.conv_transpose_same_s2(xt.detach(), wtg, torch.tensor(b, dtype=torch.float64))
    (gw,) = torch.autograd.grad(y2, wtg, nchw(dy))
    d_dw = empty(wt.size, stream)
    call("s2s_op_convt_wgrad_tf32", P(d_x), P(d_dy), P(d_dw), N, h, w, Cin, Cout, k, N, sp)
    got_w = d_dw.download(wt.shape, np.float32, stream)
    assert rel_l2(got_w, gw.numpy()) <= 2e-3, f"tf32 convT wgrad {shape} k={k}: {rel_l2(got_w, gw.numpy()):.3e}"
