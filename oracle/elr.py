"""ORACLE (test infrastructure, never on the product path): NumPy restatement of the reference's extended logistic
regression baseline, utils/training.py:402-530 (train_single_bootstrap_ELR) with preprocessing.py:270-333
(rolling_labeler_ELR), including the third-party arithmetic it calls:

  statsmodels 0.14.4 (dependencies-windows.txt:427; not installed here) — sm.GLM(y, add_constant(X), Binomial()).fit():
  IRLS (GLM._fit_irls) from mu0 = (y + 0.5)/2, weights mu(1-mu) and working response eta + (y-mu)/(mu(1-mu)) with
  p clipped to [eps, 1-eps] inside the link / variance, WLS solved by numpy.linalg.lstsq on sqrt(w)-scaled rows,
  deviance 2*sum[y log(clip(y/(mu+1e-20))) + (1-y) log(clip((1-y)/(1-mu+1e-20)))], convergence when successive deviances
  differ by <= 1e-8, at most 100 iterations; predict = 1/(1+exp(-X beta)).

Parity unpinned: neither statsmodels nor xarray can be imported here and the reference has no fixtures for this path;
pinned by analytic properties in tests/ (probabilities sum to one, monotone in the threshold, recovery of a known
logistic model).  Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np

from oracle import skill as so

EPS = np.finfo(float).eps


def glm_binomial_irls(X: np.ndarray, y: np.ndarray, maxiter: int = 100, tol: float = 1e-8):
    """-> (params, n_iterations).  X already holds the constant column."""
    X = np.asarray(X, np.float64)
    y = np.asarray(y, np.float64)
    clip = lambda p: np.clip(p, EPS, 1.0 - EPS)

    def deviance(mu):
        a = np.clip(y / (mu + 1e-20), EPS, np.inf)
        b = np.clip((1.0 - y) / (1.0 - mu + 1e-20), EPS, np.inf)
        return float(np.sum(2.0 * (y * np.log(a) + (1.0 - y) * np.log(b))))

    mu = (y + 0.5) / 2.0
    pc = clip(mu)
    eta = np.log(pc / (1.0 - pc))
    dev = [deviance(mu)]
    params = np.zeros(X.shape[1])
    it = 0
    for it in range(1, maxiter + 1):
        pc = clip(mu)
        w = pc * (1.0 - pc)
        z = eta + (y - mu) / (pc * (1.0 - pc))
        sw = np.sqrt(w)
        params = np.linalg.lstsq(X * sw[:, None], z * sw, rcond=-1)[0]
        eta = X @ params
        mu = 1.0 / (1.0 + np.exp(-eta))
        dev.append(deviance(mu))
        if abs(dev[-1] - dev[-2]) <= tol:
            break
    return params, it


def elr_masks(edges: dict, weeks_y: np.ndarray):
    """Per start: the (2,Y,X) edges of the nearest training week and the drop mask of rolling_labeler_ELR
    (preprocessing.py:304-308): an edge is NaN, e0 == 0 or e0 == e1."""
    wk = np.array(sorted(edges))
    e_t = []
    for w in weeks_y:
        d = np.abs(wk - w)
        e_t.append(edges[int(wk[np.nonzero(d == d.min())[0][-1]])])
    e_t = np.stack(e_t)                                   # (T, 2, Y, X)
    mask = np.isnan(e_t).any(1) | (e_t[:, 0] == 0) | (e_t[:, 0] == e_t[:, 1])
    return e_t, mask


def train_single_bootstrap_elr(x_train, y_train, weeks_train, x_test, weeks_test, window=1):
    """x_*: ensemble-mean predictor (T,Y,X); y_train (T,Y,X).  -> (p_train (T,Y,X,3), p_test (Tt,Y,X,3), iterations (Y,X))."""
    T, Y, X = y_train.shape
    Tt = x_test.shape[0]
    edges = so.rolling_tercile_edges(y_train, weeks_train, window=window)
    e_tr, m_tr = elr_masks(edges, weeks_train)
    _, m_te = elr_masks(edges, weeks_test)
    p_train = np.full((T, Y, X, 3), np.nan)
    p_test = np.full((Tt, Y, X, 3), np.nan)
    iters = np.zeros((Y, X), np.int32)
    for i in range(Y):
        for j in range(X):
            if np.isnan(y_train[:, i, j]).any():
                continue
            vtr, vte = ~m_tr[:, i, j], ~m_te[:, i, j]
            if not vtr.any():
                continue
            xtr, xte = x_train[vtr, i, j].astype(np.float64), x_test[vte, i, j].astype(np.float64)
            if np.isnan(xtr).any() or np.isnan(xte).any():
                continue
            if 2 * vtr.sum() <= 2 or 2 * vte.sum() <= 2:
                continue
            yv = y_train[vtr, i, j]
            y33 = (yv <= e_tr[vtr, 0, i, j]).astype(np.float64)
            y66 = (yv <= e_tr[vtr, 1, i, j]).astype(np.float64)
            n = len(xtr)
            D = np.column_stack([np.ones(2 * n), np.tile(xtr, 2), np.r_[np.full(n, 33.0), np.full(n, 67.0)]])
            params, it = glm_binomial_irls(D, np.r_[y33, y66])
            iters[i, j] = it
            sig = lambda z: 1.0 / (1.0 + np.exp(-z))
            for xs, v, out in ((xtr, vtr, p_train), (xte, vte, p_test)):
                q33 = sig(params[0] + params[1] * xs + params[2] * 33.0)
                q67 = sig(params[0] + params[1] * xs + params[2] * 67.0)
                out[:, i, j, :] = 1.0 / 3.0
                out[v, i, j, 0], out[v, i, j, 1], out[v, i, j, 2] = q33, q67 - q33, 1.0 - q67
    return p_train, p_test, iters
