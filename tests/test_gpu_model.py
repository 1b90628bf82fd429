"""GPU parity of the full U-Net path (forward, loss, backward, Adam, fit protocol, Grad-CAM) through
the host `Model` / C ABI against the torch-CPU fp64 oracle, on identical injected weights, inputs and
batch orders.  Tolerances are BASELINE.json's: forward rel-L2 <= 1e-5, per-step loss <= 1e-4 relative."""
import numpy as np
import pytest
import torch

from conftest import rel_l2
from oracle import keras_unet as ko

pytestmark = pytest.mark.gpu


def make_data(N, H, W, C, seed=0, nc=3):
    rng = np.random.default_rng(seed)
    x = rng.gamma(2.0, 3.0, size=(N, H, W, C)).astype(np.float32) / 6.0
    lab = rng.integers(0, 3, size=(N, H, W))
    y = np.eye(3, dtype=np.float32)[lab] if nc == 3 else rng.gamma(2.0, 3.0, size=(N, H, W, 1)).astype(np.float32) / 6.0
    return x, y


CONFIGS = {
    "default": dict(H=64, W=64, Cin=1, filters=2, n_blocks=3, ct_kernel=3),
    "mme_c3_ct2": dict(H=32, W=32, Cin=3, filters=2, n_blocks=3, ct_kernel=2),
    "ecmwf24_f3_ct5": dict(H=24, W=24, Cin=1, filters=3, n_blocks=3, ct_kernel=5),
    "nb4": dict(H=32, W=32, Cin=1, filters=2, n_blocks=4, ct_kernel=3),
    "nb5_f3": dict(H=64, W=64, Cin=1, filters=3, n_blocks=5, ct_kernel=5),
    "maxpool": dict(H=32, W=32, Cin=1, filters=2, n_blocks=3, ct_kernel=3, apool=False),
    "nobn": dict(H=32, W=32, Cin=2, filters=2, n_blocks=3, ct_kernel=3, bn=False),
    # BASELINE.json configs[2] as benchmarked: MME, 3 stacked model channels, 64x64, ct 3 (bench.py's workload)
    "mme_c3_64": dict(H=64, W=64, Cin=3, filters=2, n_blocks=3, ct_kernel=3),
    "relu": dict(H=32, W=32, Cin=3, filters=2, n_blocks=3, ct_kernel=3, activation="relu"),
}


def build_pair(name, N, seed=0, head="proba", precision="fp32"):
    from s2s_ismr_unet_b200.model import Model
    kw = dict(CONFIGS[name])
    cfg = ko.UnetConfig(head=head, **kw)
    w = ko.random_init(cfg, seed)
    oracle = ko.UnetOracle(cfg, w, dtype=torch.float64)
    m = Model((cfg.H, cfg.W, cfg.Cin), filters=cfg.filters, n_blocks=cfg.n_blocks, ct_kernel=cfg.ct_kernel, apool=cfg.apool,
              bn=cfg.bn, output=head, max_batch=N, weights=w, activation=cfg.activation, precision=precision)
    return cfg, w, oracle, m


@pytest.mark.parametrize("name", list(CONFIGS))
def test_param_layout_matches_oracle(name):
    cfg, w, oracle, m = build_pair(name, 2)
    assert [(d["name"], d["arena"], d["shape"]) for d in m.layout] == [(n, a, tuple(s)) for n, a, s in ko.param_specs(cfg)]
    got = m.get_weights()
    for k in w:
        np.testing.assert_array_equal(got[k], w[k])


@pytest.mark.parametrize("name", list(CONFIGS))
def test_predict_matches_oracle(name):
    cfg, w, oracle, m = build_pair(name, 4)
    x, _ = make_data(6, cfg.H, cfg.W, cfg.Cin, seed=1)
    ref = oracle.predict(x, batch_size=4)
    got = m.predict(x, batch_size=4)
    e = rel_l2(got, ref)
    assert got.shape == ref.shape and e <= 1e-5, f"{name}: predict rel-L2 {e:.3e}"
    np.testing.assert_allclose(got.sum(-1), 1.0, atol=1e-5)


# Gradient tolerance per parameter tensor (rel-L2 against the fp64 oracle).  A weight gradient is a sum of N*H*W
# (up to 65 536) signed fp32 products whose partial sums are ~sqrt(K) larger than each term and whose total nearly
# cancels for the deep BatchNorm-ed layers (d beta / d gamma of a normalised activation are differences of large sums), so
# the relative error of the SUM is amplified by the cancellation ratio, not bounded by sqrt(K) * 2^-24 as for the
# forward activations (tolerance 1e-5).  5e-5 holds for every tensor of every config with the fixed-order partial
# reductions; the forward / loss tolerances stay at BASELINE.json's 1e-5 / 1e-4.
GRAD_TOL = 5e-5


@pytest.mark.parametrize("graphs", [False, True])
@pytest.mark.parametrize("name", list(CONFIGS))
def test_backward_gradients_match_oracle(name, graphs):
    N = 4
    cfg, w, oracle, m = build_pair(name, N)
    m.compile(loss="categorical_crossentropy")
    m.set_graphs(graphs)
    x, y = make_data(N, cfg.H, cfg.W, cfg.Cin, seed=2)
    loss_ref, acc_ref, g_ref = oracle.backward(x, y)
    loss, acc = m.backward_on_batch(x, y)
    assert abs(loss - loss_ref) <= 1e-5 * abs(loss_ref), f"{name}: loss {loss} vs {loss_ref}"
    assert abs(acc - acc_ref) <= 1e-6
    g = m.get_gradients()
    bad = []
    for k, v in g_ref.items():
        e = rel_l2(g[k], v.numpy())
        if e > GRAD_TOL:
            bad.append((k, e))
    assert not bad, f"{name}: gradient mismatch {bad[:6]}"
    # BN moving statistics were updated with the batch statistics (momentum 0.99, biased variance)
    after = m.get_weights()
    ref_after = oracle.get_weights()
    for k in after:
        if "moving" in k:
            assert rel_l2(after[k], ref_after[k]) <= 1e-5, f"{name}: {k}"


@pytest.mark.parametrize("name,N", [("default", 8), ("mme_c3_ct2", 8), ("ecmwf24_f3_ct5", 8), ("maxpool", 8), ("nb4", 8), ("nb5_f3", 4),
                                    ("nobn", 8), ("relu", 8),
                                    ("mme_c3_64", 16),        # the benchmarked step: batch 16 (training.py:102, batch_size=16)
                                    ("default", 5),           # the reference's last batch: 261 = 16 * 16 + 5 starts
                                    ("ecmwf24_f3_ct5", 5)])
def test_train_steps_match_oracle(name, N):
    steps = 6
    cfg, w, oracle, m = build_pair(name, N)
    oracle.compile(lr=1e-3)
    from s2s_ismr_unet_b200.keras_api.optimizers import Adam
    m.compile(optimizer=Adam(learning_rate=1e-3), loss="categorical_crossentropy")
    for s in range(steps):
        x, y = make_data(N, cfg.H, cfg.W, cfg.Cin, seed=10 + s)
        lr_, _ = oracle.train_step(x, y)
        lg, _ = m.train_on_batch(x, y)
        assert abs(lg - lr_) <= 1e-4 * abs(lr_), f"{name} step {s}: loss {lg} vs {lr_}"
    wg, wr = m.get_weights(), oracle.get_weights()
    worst = max(rel_l2(wg[k], wr[k]) for k in wg)
    assert worst <= 1e-4, f"{name}: weights after {steps} steps rel-L2 {worst:.3e}"


def test_fp32_thick_layers_take_the_3xtf32_tensor_core_kernels_at_fp32_parity(monkeypatch):
    """fp32 precision: layers with Cin * Cout >= 96 * 96 (the deep levels of the f = 3 / n_blocks = 5 grid points) run the tcgen05
    kernels with the error-compensated 3xTF32 split — forward, dgrad, wgrad, transposed convs.  Same fp32 bars as everything
    else (forward 1e-5, gradients 2e-4 against the fp64 oracle), and really another path than the FFMA kernels
    (S2S_TC3_FP32=0)."""
    N = 4
    cfg, w, oracle, m = build_pair("nb5_f3", N)
    monkeypatch.setenv("S2S_TC3_FP32", "0")
    _, _, _, m_ffma = build_pair("nb5_f3", N)
    monkeypatch.delenv("S2S_TC3_FP32")
    x, y = make_data(N, cfg.H, cfg.W, cfg.Cin, seed=3)
    ref = oracle.predict(x, batch_size=N)
    got, got_ffma = m.predict(x, batch_size=N), m_ffma.predict(x, batch_size=N)
    assert rel_l2(got, ref) <= 1e-5 and rel_l2(got_ffma, ref) <= 1e-5, (rel_l2(got, ref), rel_l2(got_ffma, ref))
    assert not np.array_equal(got, got_ffma), "the 3xTF32 tensor-core path was not taken"
    m.compile(loss="categorical_crossentropy")
    loss_ref, _, g_ref = oracle.backward(x, y)
    loss, _ = m.backward_on_batch(x, y)
    assert abs(loss - loss_ref) <= 1e-4 * abs(loss_ref)
    g = m.get_gradients()
    worst = max(rel_l2(g[k], v.numpy()) for k, v in g_ref.items())
    assert worst <= 2e-4, f"fp32 (3xTF32 thick layers) gradients rel-L2 {worst:.3e}"


@pytest.mark.parametrize("name,N", [("mme_c3_64", 16), ("maxpool", 8), ("nb4", 5), ("mme_c3_ct2", 8), ("ecmwf24_f3_ct5", 8)])
def test_bn_backward_statistics_in_the_producer_epilogue_match_the_reduce_kernel(name, N, monkeypatch):
    """The (sum dc, sum dc*xhat) partials of BatchNorm backward are written by the epilogue of the kernel that produces dc
    (transposed-conv input gradient for the un-pooled layers, the next level's first 3x3 input gradient for the pooled ones:
    average AND max pooling) instead of by bn_bwd_reduce.  Both paths against the fp64 oracle, against each other, and the
    folded one really launches fewer kernels."""
    cfg, w, oracle, m = build_pair(name, N)
    monkeypatch.setenv("S2S_NO_BN_FOLD", "1")
    _, _, _, m_ref = build_pair(name, N)
    monkeypatch.delenv("S2S_NO_BN_FOLD")
    x, y = make_data(N, cfg.H, cfg.W, cfg.Cin, seed=4)
    _, _, g_or = oracle.backward(x, y)
    got = []
    for mm in (m, m_ref):
        mm.compile(loss="categorical_crossentropy")
        mm.set_graphs(False)
        before = mm.launch_count()
        mm.backward_on_batch(x, y)
        got.append((mm.get_gradients(), mm.launch_count() - before))
    (g, n_fold), (g_ref, n_plain) = got
    # one bn_bwd_reduce per BatchNorm layer (2 n_blocks of them) goes away; where the producer runs on the tensor cores
    # (the thick layers of the filters = 3 nets) the reduce kernel stays
    if cfg.filters == 2:
        assert n_fold == n_plain - 2 * cfg.n_blocks, (n_fold, n_plain)
    else:
        assert n_plain - 2 * cfg.n_blocks <= n_fold < n_plain, (n_fold, n_plain)
    for k, v in g_or.items():
        assert rel_l2(g[k], v.numpy()) <= GRAD_TOL, (name, k, rel_l2(g[k], v.numpy()))
        assert rel_l2(g[k], g_ref[k]) <= 2e-5, (name, k, rel_l2(g[k], g_ref[k]))


def test_large_batch_train_steps_match_oracle():
    """Batch 32 at 64x64 crosses into the throughput kernels (16x32 tiles of gconv.cuh incl. its BatchNorm partials,
    batch-scaled reduction slots, pixel-split wgrad)."""
    N, steps = 32, 3
    cfg, w, oracle, m = build_pair("default", N)
    oracle.compile(lr=1e-3)
    from s2s_ismr_unet_b200.keras_api.optimizers import Adam
    m.compile(optimizer=Adam(learning_rate=1e-3), loss="categorical_crossentropy")
    for s in range(steps):
        x, y = make_data(N, cfg.H, cfg.W, cfg.Cin, seed=20 + s)
        lr_, _ = oracle.train_step(x, y)
        lg, _ = m.train_on_batch(x, y)
        assert abs(lg - lr_) <= 1e-4 * abs(lr_), f"step {s}: loss {lg} vs {lr_}"
    wg, wr = m.get_weights(), oracle.get_weights()
    worst = max(rel_l2(wg[k], wr[k]) for k in wg)
    assert worst <= 1e-4, f"weights after {steps} steps rel-L2 {worst:.3e}"
    x, _ = make_data(40, cfg.H, cfg.W, cfg.Cin, seed=31)
    e = rel_l2(m.predict(x, batch_size=32), oracle.predict(x, batch_size=32))
    assert e <= 1e-5, f"predict rel-L2 {e:.3e}"


def test_pinned_host_batches_take_the_single_call_path_with_identical_results():
    """train_on_batch on pinned host batches = one C call (s2s_unet_train_step_host: H2D, step, D2H, sync); it must give
    bit-identical losses and weights to the staged path used for pageable arrays."""
    from s2s_ismr_unet_b200.runtime import is_pinned, pinned_empty
    cfg, w, _, a = build_pair("mme_c3_ct2", 8)
    _, _, _, b = build_pair("mme_c3_ct2", 8)
    a.compile(loss="categorical_crossentropy")
    b.compile(loss="categorical_crossentropy")
    x, y = make_data(8, cfg.H, cfg.W, cfg.Cin, seed=5)
    px, py = pinned_empty(x.shape), pinned_empty(y.shape)
    px[...], py[...] = x, y
    assert is_pinned(px) and is_pinned(py) and not is_pinned(x)
    for _ in range(3):
        assert a.train_on_batch(x, y) == b.train_on_batch(px, py)
    wa, wb = a.get_weights(), b.get_weights()
    for k in wa:
        np.testing.assert_array_equal(wa[k], wb[k])


def test_streamed_host_batches_match_step_by_step_calls():
    """train_on_batches (s2s_unet_train_steps_host: the copy of batch i + 1 is staged while step i computes) must give
    bit-identical losses and weights to one train_on_batch call per batch; 5 different batches, so a stale or swapped
    staging slot would show."""
    from s2s_ismr_unet_b200.runtime import pinned_empty
    cfg, w, _, a = build_pair("mme_c3_ct2", 8)
    _, _, _, b = build_pair("mme_c3_ct2", 8)
    a.compile(loss="categorical_crossentropy")
    b.compile(loss="categorical_crossentropy")
    xs, ys = [], []
    for i in range(5):
        x, y = make_data(8, cfg.H, cfg.W, cfg.Cin, seed=50 + i)
        px, py = pinned_empty(x.shape), pinned_empty(y.shape)
        px[...], py[...] = x, y
        xs.append(px), ys.append(py)
    seq = np.array([a.train_on_batch(x, y) for x, y in zip(xs, ys)], np.float32)
    got = b.train_on_batches(xs, ys)
    np.testing.assert_array_equal(got, seq)
    wa, wb = a.get_weights(), b.get_weights()
    for k in wa:
        np.testing.assert_array_equal(wa[k], wb[k])
    # ordinary (pageable) NumPy batches — what the reference hands to model.fit — go through the pinned staging ring of the C
    # call: 9 batches (three times around the 3-slot ring), mixed with pinned ones, bit-identical to step-by-step calls
    xs2, ys2 = [], []
    for i in range(9):
        x, y = make_data(8, cfg.H, cfg.W, cfg.Cin, seed=70 + i)
        if i == 4:
            px, py = pinned_empty(x.shape), pinned_empty(y.shape)
            px[...], py[...] = x, y
            x, y = px, py
        xs2.append(x), ys2.append(y)
    seq2 = np.array([a.train_on_batch(x, y) for x, y in zip(xs2, ys2)], np.float32)
    got2 = b.train_on_batches(xs2, ys2)
    np.testing.assert_array_equal(got2, seq2)
    wa, wb = a.get_weights(), b.get_weights()
    for k in wa:
        np.testing.assert_array_equal(wa[k], wb[k])
    with pytest.raises(ValueError):
        b.train_on_batches([np.zeros((7, cfg.H, cfg.W, cfg.Cin), np.float32)], [ys[0]])       # ragged batch sizes are refused


def test_graph_replay_is_bitwise_identical_to_eager():
    N = 8
    outs = []
    for graphs in (False, True):
        cfg, w, oracle, m = build_pair("default", N)
        m.compile(loss="categorical_crossentropy")
        m.set_graphs(graphs)
        for s in range(3):
            x, y = make_data(N, cfg.H, cfg.W, cfg.Cin, seed=20 + s)
            m.train_on_batch(x, y)
        outs.append(m.get_weights())
    for k in outs[0]:
        np.testing.assert_array_equal(outs[0][k], outs[1][k], err_msg=k)


def test_fit_protocol_matches_oracle(tmp_path):
    """17-sample epochs at batch 8 (partial last batch of 1), injected orders, val pass, EarlyStopping."""
    from s2s_ismr_unet_b200.keras_api.callbacks import EarlyStopping, ModelCheckpoint
    from s2s_ismr_unet_b200.keras_api.optimizers import Adam
    from s2s_ismr_unet_b200.model import load_model
    cfg, w, oracle, m = build_pair("mme_c3_ct2", 8)
    x, y = make_data(17, cfg.H, cfg.W, cfg.Cin, seed=3)
    xv, yv = make_data(9, cfg.H, cfg.W, cfg.Cin, seed=4)
    epochs = 4
    rng = np.random.default_rng(7)
    orders = [rng.permutation(17) for _ in range(epochs)]
    oracle.compile(lr=1e-3)
    href = oracle.fit(x, y, (xv, yv), epochs, 8, orders, patience=10)
    m.compile(optimizer=Adam(1e-3), loss="categorical_crossentropy", metrics=["accuracy"])
    ck = tmp_path / "models" / "best.keras"
    h = m.fit(x=x, y=y, validation_data=(xv, yv), epochs=epochs, batch_size=8, shuffle=True, verbose=0, _orders=orders,
              callbacks=[ModelCheckpoint(str(ck), save_best_only=True, monitor="val_loss", mode="min"),
                         EarlyStopping(monitor="val_loss", patience=10, restore_best_weights=True)])
    np.testing.assert_allclose(h.history["loss"], href["loss"], rtol=1e-4)
    np.testing.assert_allclose(h.history["val_loss"], href["val_loss"], rtol=1e-4)
    assert ck.exists()
    best = load_model(str(ck))
    xt, _ = make_data(5, cfg.H, cfg.W, cfg.Cin, seed=5)
    assert rel_l2(best.predict(xt), oracle.predict(xt)) <= 1e-4      # oracle restored its best weights too


def test_masked_mse_head_matches_oracle():
    N = 4
    cfg, w, oracle, m = build_pair("default", N, head="deterministic")
    rng = np.random.default_rng(9)
    mask = (rng.random((cfg.H, cfg.W)) < 0.4)
    x, y = make_data(N, cfg.H, cfg.W, cfg.Cin, seed=6, nc=1)
    from s2s_ismr_unet_b200.runtime import DeviceBuffer
    m.compile(loss="masked_mse")
    dm = DeviceBuffer.from_array(mask.astype(np.uint8), m.stream)
    loss_ref, _, g_ref = oracle.backward(x, y, mask=mask)
    loss, _ = m.backward_on_batch(x, y, mask_ptr=dm.ptr)
    assert abs(loss - loss_ref) <= 1e-5 * abs(loss_ref)
    g = m.get_gradients()
    worst = max(rel_l2(g[k], v.numpy()) for k, v in g_ref.items())
    assert worst <= GRAD_TOL, f"masked-MSE grads rel-L2 {worst:.3e}"


@pytest.mark.parametrize("layer", ["bottleneck", "conv2d", "up_conv1_3", "up_conv2_2", "up_conv3_1", "down_conv2_2", "down_conv1_1"])
def test_gradcam_matches_oracle(layer):
    cfg, w, oracle, m = build_pair("default", 4)
    x, _ = make_data(4, cfg.H, cfg.W, cfg.Cin, seed=8)
    ref = oracle.gradcam(x, layer, cls=2)
    got = m.gradcam(x, layer, cls=2, batch_size=4)
    assert got.shape == ref.shape
    scale = max(float(np.abs(ref).max()), 1e-12)
    assert np.abs(got - ref).max() <= 1e-4 * scale + 1e-9, f"{layer}: max err {np.abs(got - ref).max():.3e} (scale {scale:.3e})"


def test_invalid_grid_is_rejected():
    from s2s_ismr_unet_b200.model import Model
    with pytest.raises(ValueError, match="not divisible"):
        Model((24, 24, 1), n_blocks=4)       # 24 / 16 = 1.5: Keras would raise on Concatenate


def test_training_is_bit_reproducible():
    N = 8
    res = []
    for _ in range(2):
        cfg, w, oracle, m = build_pair("default", N)
        m.compile(loss="categorical_crossentropy")
        for s in range(3):
            x, y = make_data(N, cfg.H, cfg.W, cfg.Cin, seed=30 + s)
            m.train_on_batch(x, y)
        res.append(m.get_weights())
    for k in res[0]:
        np.testing.assert_array_equal(res[0][k], res[1][k], err_msg=k)


@pytest.mark.parametrize("name", ["default", "nb4", "nb5_f3", "maxpool"])
def test_predict_bf16_tensor_core_mode(name):
    """precision='bf16_tc': thick layers (Cin % 32 == 0) run on tcgen05 tensor cores with bf16 operands.
    BASELINE tolerance for bf16: forward rel-L2 <= 1e-2; and the path must really differ from the fp32 one."""
    from s2s_ismr_unet_b200.model import Model
    kw = dict(CONFIGS[name])
    cfg = ko.UnetConfig(**kw)
    w = ko.random_init(cfg, 5)
    oracle = ko.UnetOracle(cfg, w, dtype=torch.float64)
    x, _ = make_data(6, cfg.H, cfg.W, cfg.Cin, seed=9)
    ref = oracle.predict(x, batch_size=4)
    common = dict(filters=cfg.filters, n_blocks=cfg.n_blocks, ct_kernel=cfg.ct_kernel, apool=cfg.apool, bn=cfg.bn, max_batch=4, weights=w)
    m_tc = Model((cfg.H, cfg.W, cfg.Cin), precision="bf16_tc", **common)
    m_32 = Model((cfg.H, cfg.W, cfg.Cin), **common)
    got, got32 = m_tc.predict(x, batch_size=4), m_32.predict(x, batch_size=4)
    e = rel_l2(got, ref)
    assert e <= 1e-2, f"{name}: bf16 tensor-core predict rel-L2 {e:.3e}"
    assert rel_l2(got32, ref) <= 1e-5
    assert not np.array_equal(got, got32), "tensor-core path was not taken"
    np.testing.assert_allclose(got.sum(-1), 1.0, atol=1e-5)


def test_predict_256x256_matches_oracle():
    """BASELINE.json configs[4]: real-time MME inference on the 0.25-degree (256 x 256) grid."""
    cfg = ko.UnetConfig(H=256, W=256, Cin=3, filters=2, n_blocks=3, ct_kernel=3)
    w = ko.random_init(cfg, 3)
    from s2s_ismr_unet_b200.model import Model
    m = Model((256, 256, 3), filters=2, n_blocks=3, ct_kernel=3, max_batch=2, weights=w)
    x, _ = make_data(3, 256, 256, 3, seed=12)
    ref = ko.UnetOracle(cfg, w, dtype=torch.float64).predict(x, batch_size=2)
    got = m.predict(x, batch_size=2)
    assert rel_l2(got, ref) <= 1e-5, f"256x256 predict rel-L2 {rel_l2(got, ref):.3e}"


# ---------------------------------------------------------------------------------------------- precision="tf32"
# BASELINE.json: reduced-precision forward outputs rel-L2 <= 1e-2.  Per-step loss: the tf32 products carry ~5e-4
# relative error per layer, which reaches the loss at ~1e-3; tolerance 5e-3 relative per step, stated here.
TF32_LOSS_RTOL = 5e-3


@pytest.mark.parametrize("name", ["default", "mme_c3_64", "nb4", "nb5_f3", "maxpool", "ecmwf24_f3_ct5"])
def test_tf32_tensor_core_training_and_predict(name):
    """precision='tf32': every 3x3 convolution with >= 8 input channels runs forward AND input-gradient on tcgen05 (kind::tf32,
    TMA-fed, accumulators in TMEM); the thick / small-image layers also their weight gradients (tcwgrad.cuh) and the thick
    transposed convolutions forward, dgrad and wgrad (nb5_f3 = the largest tuning-grid point exercises all of them, flat
    geometry included); everything else in fp32.  Trains (6 Adam steps) within the stated loss tolerance, predicts within
    1e-2, and really takes another path than fp32."""
    N, steps = 8, 6
    cfg, w, oracle, m = build_pair(name, N, precision="tf32")
    _, _, _, m32 = build_pair(name, N)
    oracle.compile(lr=1e-3)
    m.compile(loss="categorical_crossentropy")
    for s_ in range(steps):
        x, y = make_data(N, cfg.H, cfg.W, cfg.Cin, seed=40 + s_)
        lr_, _ = oracle.train_step(x, y)
        lg, _ = m.train_on_batch(x, y)
        assert abs(lg - lr_) <= TF32_LOSS_RTOL * abs(lr_), f"{name} step {s_}: tf32 loss {lg} vs {lr_}"
    x, _ = make_data(N, cfg.H, cfg.W, cfg.Cin, seed=50)
    ref = oracle.predict(x, batch_size=N)
    got = m.predict(x, batch_size=N)
    assert rel_l2(got, ref) <= 1e-2, f"{name}: tf32 predict after training rel-L2 {rel_l2(got, ref):.3e}"
    m32.set_weights(m.get_weights())
    assert not np.array_equal(m32.predict(x, batch_size=N), got), "the tensor-core path was not taken"


def test_tf32_gradients_within_reduced_precision_tolerance():
    cfg, w, oracle, m = build_pair("mme_c3_64", 8, precision="tf32")
    m.compile(loss="categorical_crossentropy")
    x, y = make_data(8, cfg.H, cfg.W, cfg.Cin, seed=2)
    loss_ref, _, g_ref = oracle.backward(x, y)
    loss, _ = m.backward_on_batch(x, y)
    assert abs(loss - loss_ref) <= TF32_LOSS_RTOL * abs(loss_ref)
    g = m.get_gradients()
    worst = max(rel_l2(g[k], v.numpy()) for k, v in g_ref.items())
    assert worst <= 1e-2, f"tf32 gradients rel-L2 {worst:.3e}"


def test_keras_archive_save_load_resumes_training_bit_identically(tmp_path):
    """model.save -> Keras-3 `.keras` archive (config.json + model.weights.h5 in Keras' layer / variable paths) ->
    load_model: same predictions, and one more Adam step on the reloaded model equals the step on the original
    (weights, moments and the iteration counter all travel through the archive)."""
    import zipfile
    from s2s_ismr_unet_b200.model import load_model
    cfg, w, _, a = build_pair("mme_c3_ct2", 32)      # max_batch 32 = load_model's default: the same reduction-slot plan
    a.compile(loss="categorical_crossentropy")
    for s_ in range(2):
        x, y = make_data(8, cfg.H, cfg.W, cfg.Cin, seed=60 + s_)
        a.train_on_batch(x, y)
    path = tmp_path / "models" / "best_model_unet_0.keras"
    a.save(str(path))
    assert {"config.json", "metadata.json", "model.weights.h5"} <= set(zipfile.ZipFile(path).namelist())
    b = load_model(str(path))
    x, y = make_data(8, cfg.H, cfg.W, cfg.Cin, seed=70)
    np.testing.assert_array_equal(a.predict(x, batch_size=8), b.predict(x, batch_size=8))
    assert a.train_on_batch(x, y) == b.train_on_batch(x, y)
    wa, wb = a.get_weights(), b.get_weights()
    for k in wa:
        np.testing.assert_array_equal(wa[k], wb[k], err_msg=k)
