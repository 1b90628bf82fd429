"""Known-answer / self-consistency pins of the oracle (the reference has no tests or golden vectors,
SURVEY §4): every Keras-specific semantic the CUDA path is checked against is itself checked here
against an independent statement (naive loops, autograd identities, closed forms)."""
import math

import numpy as np
import torch
import torch.nn.functional as F

from oracle import keras_unet as ko
from oracle import skill as so

D = torch.float64


def test_parameter_counts_match_the_reference_models():
    def counts(**k):
        sp = ko.param_specs(ko.UnetConfig(H=64, W=64, **k))
        total = sum(int(np.prod(s)) for _, _, s in sp)
        trainable = sum(int(np.prod(s)) for _, a, s in sp if a == 0)
        bn_ch = sum(int(np.prod(s)) for n, _, s in sp if n.endswith("/gamma"))
        return total, trainable, total - 4 * bn_ch
    # Keras model.count_params() at defaults = 134 811; SURVEY §8d's "trainable P" column
    # counts conv / transposed-conv kernels + biases only (total minus the 4 BN vectors per BN layer).
    assert counts() == (134_811, 134_475, 134_139)
    assert counts(filters=3, n_blocks=5, ct_kernel=5)[2] == 6_436_695
    assert counts(Cin=3)[2] == 134_283
    assert counts(Cin=3, filters=3, n_blocks=5, ct_kernel=3)[2] == 4_865_583


def test_conv3x3_same_equals_naive_cross_correlation():
    rng = np.random.default_rng(0)
    x = rng.normal(size=(1, 5, 6, 2))
    w = rng.normal(size=(3, 3, 2, 3))
    b = rng.normal(size=(3,))
    ref = np.zeros((1, 5, 6, 3))
    xp = np.pad(x, ((0, 0), (1, 1), (1, 1), (0, 0)))
    for y in range(5):
        for xx in range(6):
            for ky in range(3):
                for kx in range(3):
                    ref[0, y, xx] += xp[0, y + ky, xx + kx] @ w[ky, kx]
    ref += b
    got = ko.conv3x3_same(torch.tensor(x).permute(0, 3, 1, 2), torch.tensor(w), torch.tensor(b)).permute(0, 2, 3, 1).numpy()
    np.testing.assert_allclose(got, ref, atol=1e-12)


def test_conv_transpose_same_is_gradient_of_tf_same_strided_conv():
    for k in (2, 3, 5):
        h, w, ci, co = 5, 4, 3, 2
        x = torch.randn(2, ci, h, w, dtype=D)
        W = torch.randn(k, k, co, ci, dtype=D)
        y = ko.conv_transpose_same_s2(x, W, torch.zeros(co, dtype=D))
        z = torch.randn(2, co, 2 * h, 2 * w, dtype=D, requires_grad=True)
        pt = k - 2
        pb = pt // 2
        out = F.conv2d(F.pad(z, (pb, pt - pb, pb, pt - pb)), W.permute(3, 2, 0, 1), stride=2)   # TF SAME: extra pad at the end
        (g,) = torch.autograd.grad(out, z, x)
        assert float((g - y).abs().max()) < 1e-12
        # explicit scatter definition: y[2i+ky-pb, 2j+kx-pb, co] += x[i,j,ci] W[ky,kx,co,ci]
        ref = np.zeros((2, 2 * h, 2 * w, co))
        xn, Wn = x.permute(0, 2, 3, 1).numpy(), W.numpy()
        for i in range(h):
            for j in range(w):
                for ky in range(k):
                    for kx in range(k):
                        oy, ox = 2 * i + ky - pb, 2 * j + kx - pb
                        if 0 <= oy < 2 * h and 0 <= ox < 2 * w:
                            ref[:, oy, ox] += xn[:, i, j] @ Wn[ky, kx].T
        np.testing.assert_allclose(y.permute(0, 2, 3, 1).numpy(), ref, atol=1e-12)


def test_batchnorm_uses_biased_variance_eps_1e3_momentum_099():
    x = torch.randn(4, 3, 5, 5, dtype=D) * 2 + 1
    g, b = torch.tensor([1.5, 0.5, 1.0], dtype=D), torch.tensor([0.1, -0.2, 0.0], dtype=D)
    mm, mv = torch.zeros(3, dtype=D), torch.ones(3, dtype=D)
    y, nmm, nmv = ko.batchnorm(x, g, b, mm, mv, True, 1e-3, 0.99)
    mean = x.mean((0, 2, 3))
    var = x.var((0, 2, 3), unbiased=False)
    ref = (x - mean.view(1, -1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1) + 1e-3) * g.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)
    assert float((y - ref).abs().max()) < 1e-12
    np.testing.assert_allclose(nmm.numpy(), 0.01 * mean.numpy(), atol=1e-14)
    np.testing.assert_allclose(nmv.numpy(), 0.99 + 0.01 * var.numpy(), atol=1e-14)
    yi, _, _ = ko.batchnorm(x, g, b, nmm, nmv, False, 1e-3, 0.99)
    refi = (x - nmm.view(1, -1, 1, 1)) / torch.sqrt(nmv.view(1, -1, 1, 1) + 1e-3) * g.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)
    assert float((yi - refi).abs().max()) < 1e-12


def test_cce_known_answers():
    t = torch.zeros(2, 3, 4, 4, dtype=D)
    t[:, 1] = 1
    assert abs(float(ko.keras_cce(torch.full((2, 3, 4, 4), 1 / 3, dtype=D), t)) - math.log(3)) < 1e-12
    p = torch.zeros(2, 3, 4, 4, dtype=D)
    p[:, 1] = 1                                                  # perfect forecast: clipped at 1 - 1e-7
    assert abs(float(ko.keras_cce(p, t)) + math.log(1 - 1e-7)) < 1e-12
    p = torch.zeros(2, 3, 4, 4, dtype=D)
    p[:, 0] = 1                                                  # certain and wrong: -log(1e-7)
    assert abs(float(ko.keras_cce(p, t)) + math.log(1e-7)) < 1e-9


def test_keras_adam_closed_form_on_constant_gradient():
    """With a constant gradient g: m_t = g(1-b1^t), v_t = g^2(1-b2^t) so the bias-corrected ratio is
    sign(g) up to eps and every step moves the weight by ~lr."""
    cfg = ko.UnetConfig(H=8, W=8, n_blocks=1)
    o = ko.UnetOracle(cfg, ko.glorot_uniform_init(cfg, 0))
    o.compile(lr=1e-3)
    name = o.trainable[0]
    w0 = o.p[name].detach().clone()
    grads = {n: torch.full_like(o.p[n], 0.5) for n in o.trainable}
    for t in range(1, 6):
        o.apply_adam(grads)
        alpha = 1e-3 * math.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t)
        step = alpha * (0.5 * (1 - 0.9 ** t)) / (math.sqrt(0.25 * (1 - 0.999 ** t)) + 1e-7)
        w0 = w0 - step
        assert float((o.p[name].detach() - w0).abs().max()) < 1e-15
    assert abs(float((o.p[name].detach() - w0).abs().max())) < 1e-12


def test_forward_shapes_and_softmax_rows_for_the_tuning_grid():
    rng = np.random.default_rng(0)
    for nb, f, k, hw in [(3, 2, 2, 24), (3, 3, 5, 32), (4, 2, 3, 32), (5, 2, 3, 32), (5, 3, 5, 64)]:
        cfg = ko.UnetConfig(H=hw, W=hw, n_blocks=nb, filters=f, ct_kernel=k)
        o = ko.UnetOracle(cfg, ko.glorot_uniform_init(cfg, 1))
        out = o.predict(rng.normal(size=(2, hw, hw)).astype(np.float32))       # rank-3 input auto-expands
        assert out.shape == (2, hw, hw, 3)
        np.testing.assert_allclose(out.sum(-1), 1.0, atol=1e-6)
    import pytest
    with pytest.raises(ValueError):
        ko.UnetOracle(ko.UnetConfig(H=24, W=24, n_blocks=4), {})


def test_autograd_gradient_matches_finite_differences():
    cfg = ko.UnetConfig(H=8, W=8, Cin=2, n_blocks=2, ct_kernel=3)
    w = ko.random_init(cfg, 3)
    o = ko.UnetOracle(cfg, w)
    rng = np.random.default_rng(1)
    x = rng.normal(size=(3, 8, 8, 2))
    y = np.eye(3)[rng.integers(0, 3, (3, 8, 8))]
    _, _, g = o.backward(x, y)
    for name in ["down_conv1_1/kernel", "up_conv2_1/kernel", "batch_normalization/gamma", "conv2d_1/bias"]:
        idx = tuple(0 for _ in w[name].shape)
        eps = 1e-5
        vals = []
        for s in (+1, -1):
            w2 = {k: v.copy().astype(np.float64) for k, v in w.items()}
            w2[name][idx] += s * eps
            o2 = ko.UnetOracle(cfg, w2)
            out = o2.forward(x, training=True, update_moving=False)
            vals.append(float(o2.loss_and_acc(out, y)[0]))
        fd = (vals[0] - vals[1]) / (2 * eps)
        assert abs(fd - float(g[name][idx])) < 1e-6 * max(1.0, abs(fd)), name


def test_rps_pearson_and_acc_known_answers():
    T, Y, X = 9, 2, 2
    lab = np.tile(np.array([0, 1, 2] * 3, float)[:, None, None], (1, Y, X))
    o = so.onehot_obs(lab)
    np.testing.assert_allclose(so.rps(o, o), 0.0)
    clim = so.climo_forecast((T, Y, X))
    np.testing.assert_allclose(so.rps(o, clim), (5 / 9 + 2 / 9 + 5 / 9) / 3)
    np.testing.assert_allclose(so.rpss(clim, clim, o), 0.0, atol=1e-15)
    np.testing.assert_allclose(so.rpss(clim, o, o), 1.0)
    rng = np.random.default_rng(0)
    x = rng.normal(size=(40, 3, 3))
    np.testing.assert_allclose(so.pearson_t(x, 2.5 * x + 1), 1.0)
    np.testing.assert_allclose(so.pearson_t(x, -0.5 * x + 4), -1.0)
    week = rng.integers(20, 30, size=40)
    y = 0.3 * x + rng.normal(size=x.shape)
    acc0, _ = so.acc_cc(x, y, week)
    off = rng.normal(size=31)[week][:, None, None] * 7.0          # per-week constant offsets do not change the ACC
    acc1, cc1 = so.acc_cc(x + off, y - 2 * off, week)
    np.testing.assert_allclose(acc1, acc0, atol=1e-12)
    assert np.abs(cc1 - acc0).max() > 1e-3


def test_to_categorical_and_mme_combine():
    y = np.array([[0, 1], [2, np.nan]])
    oh = so.to_categorical(y, 3)
    assert oh.shape == (2, 2, 3) and oh[1, 0, 2] == 1 and oh[1, 1, 0] == 1       # NaN -> class 0 (cast semantics)
    p = [np.array([[0.2, 0.3, 0.5]]), np.array([[0.6, 0.2, 0.2]])]
    np.testing.assert_allclose(so.mme_combine(p), [[0.4, 0.25, 0.35]])


def test_elr_oracle_satisfies_the_score_equations_and_known_answers():
    """Analytic pins of oracle/elr.py (statsmodels itself cannot be imported): at the IRLS fixed point the binomial
    score X^T (y - mu) vanishes; an intercept-only model returns logit(mean y); a separable-free two-group design returns
    the group log-odds; linear-interpolated tercile edges of an arithmetic sequence are known in closed form."""
    from oracle import elr as eo
    rng = np.random.default_rng(0)
    n = 600
    x = rng.normal(size=n)
    q = np.r_[np.full(n // 2, 33.0), np.full(n // 2, 67.0)]
    X = np.column_stack([np.ones(n), x, q])
    y = (rng.random(n) < 1 / (1 + np.exp(-(-2.0 + 0.8 * x + 0.04 * q)))).astype(float)
    beta, it = eo.glm_binomial_irls(X, y)
    mu = 1 / (1 + np.exp(-X @ beta))
    assert it < 15 and np.abs(X.T @ (y - mu)).max() < 1e-6
    b0, _ = eo.glm_binomial_irls(np.ones((n, 1)), y)
    assert abs(b0[0] - np.log(y.mean() / (1 - y.mean()))) < 1e-9
    g = (np.arange(n) % 2).astype(float)
    yy = np.where(g == 1, rng.random(n) < 0.7, rng.random(n) < 0.2).astype(float)
    bg, _ = eo.glm_binomial_irls(np.column_stack([np.ones(n), g]), yy)
    p0, p1 = yy[g == 0].mean(), yy[g == 1].mean()
    np.testing.assert_allclose(bg, [np.log(p0 / (1 - p0)), np.log(p1 / (1 - p1)) - np.log(p0 / (1 - p0))], atol=1e-8)
    # tercile edges of 0..9 (n = 10): virtual indices 3 and 6 -> exactly 3 and 6; of 0..10 (n = 11): 10/3 and 20/3
    e = so.rolling_tercile_edges(np.arange(10.0)[:, None], np.full(10, 25), window=0)[25]
    np.testing.assert_array_equal(e[:, 0], [3.0, 6.0])
    e = so.rolling_tercile_edges(np.arange(11.0)[:, None], np.full(11, 25), window=0)[25]
    np.testing.assert_allclose(e[:, 0], [10 / 3, 20 / 3], rtol=1e-15)
    lab = so.apply_tercile_labels(np.arange(10.0)[:, None], np.full(10, 25), {25: np.array([[3.0], [6.0]])})
    np.testing.assert_array_equal(lab[:, 0], [0, 0, 0, 1, 1, 1, 1, 2, 2, 2])       # y == edge stays in the middle class
