"""Distributional check that does not involve the oracle: for a LINEAR target equal to its auxiliary law the guided proposal
is exact, so (i) every pCN proposal is accepted (ll° == ll up to rounding, because the ll integrand G vanishes) and (ii) the
ensemble of smoothed paths must follow the exact Gaussian smoothing distribution, which scipy gives in closed form
(Kalman filter + RTS smoother on the exactly discretised state-space model)."""
import numpy as np
import pytest
from scipy.linalg import expm

import dmt_b200
from dmt_b200 import _lib, configs
from harness import make_ctx

pytestmark = pytest.mark.gpu


def exact_disc(Bm, beta, a, h):
    d = Bm.shape[0]
    A = np.zeros((d + 1, d + 1)); A[:d, :d] = Bm; A[:d, d] = beta
    E = expm(A * h)
    V = np.zeros((2 * d, 2 * d)); V[:d, :d] = -Bm; V[:d, d:] = a; V[d:, d:] = Bm.T
    EV = expm(V * h)
    return E[:d, :d], E[:d, d], EV[d:, d:].T @ EV[:d, d:]


def rts_smoother(Bm, beta, a, x0, times, obs_idx, L, Sig, vs):
    """exact smoothing mean / covariance at `times` (times[0] = 0 with known x0) given observations at indices obs_idx"""
    d = len(x0); n = len(times)
    mf = np.zeros((n, d)); Pf = np.zeros((n, d, d)); mp = np.zeros((n, d)); Pp = np.zeros((n, d, d)); Phis = [None] * n
    mf[0] = x0
    o = dict(zip(obs_idx, vs))
    for i in range(1, n):
        Phi, mu, Q = exact_disc(Bm, beta, a, times[i] - times[i - 1])
        Phis[i] = Phi
        mp[i] = Phi @ mf[i - 1] + mu; Pp[i] = Phi @ Pf[i - 1] @ Phi.T + Q
        if i in o:
            S = L @ Pp[i] @ L.T + Sig; Kg = Pp[i] @ L.T @ np.linalg.inv(S)
            mf[i] = mp[i] + Kg @ (o[i] - L @ mp[i]); Pf[i] = Pp[i] - Kg @ S @ Kg.T
        else:
            mf[i], Pf[i] = mp[i], Pp[i]
    ms, Ps = mf.copy(), Pf.copy()
    for i in range(n - 2, -1, -1):
        if i == 0:
            break
        C = Pf[i] @ Phis[i + 1].T @ np.linalg.inv(Pp[i + 1])
        ms[i] = mf[i] + C @ (ms[i + 1] - mp[i + 1]); Ps[i] = Pf[i] + C @ (Ps[i + 1] - Pp[i + 1]) @ C.T
    return ms, Ps


def test_linear_target_smoothing_distribution_is_exact():
    M, K = 4096, 4
    prob = configs.make_problem("ou2", M, P=1, K=K, dt=0.002, seed=4, rho=0.5)
    th = prob.theta
    Bm = th[:4].reshape(2, 2); beta = th[4:6]; a = np.diag(th[6:8] ** 2)
    ctx = make_ctx(prob, seed=12, ll_hist_len=8)
    # the auxiliary law must BE the target: B, beta, sigma of the OU model (not a linearisation artefact: it is linear)
    ctx.recompute_guiding_term(0, _lib.P_ONLY)
    assert ctx.init_paths(0, 0, 5) == 0
    ctx.loglikhd(0, 0, 0)
    ll0 = ctx.get_ll(0, 0)[0]
    assert np.ptp(ll0) < 1e-9 * max(1.0, abs(ll0[0]))          # ll is path independent: log h~(0, x0) for every chain
    for it in range(8):
        ctx.draw_proposal_path(0, it)
        ctx.accept_reject_path(0, it)
    assert ctx.get_accept_history(0, 0, 7).all()               # exact proposals are always accepted
    X = ctx.get_X(0)                                           # [NP, d, M]
    # exact smoother on the imputation grid of the first two intervals (observations at the interval ends)
    pt0 = np.concatenate([[0], np.cumsum(prob.n_pts)])
    times, idx_of = [0.0], {}
    obs_idx = []
    for k in range(K):
        tk = prob.tt[pt0[k]:pt0[k + 1]]
        for j in range(1, len(tk)):
            times.append(tk[j]); idx_of[(k, j)] = len(times) - 1
        obs_idx.append(len(times) - 1)
    # Euler–Maruyama on the tau-grid is not the exact flow: compare at the smoothing level with a tolerance that covers the
    # O(dt) discretisation bias (dt = 2e-3) plus the Monte-Carlo error of 4096 chains
    ms, Ps = rts_smoother(Bm, beta, a, prob.x0[:, 0], np.array(times), obs_idx, prob.L, prob.Sigma, [prob.v[k, :, 0] for k in range(K)])
    for (k, j) in [(0, 20), (1, 25), (2, 10), (3, 40)]:
        g = idx_of[(k, j)]
        xs = X[pt0[k] + j]                                     # [d, M]
        sd = np.sqrt(np.diag(Ps[g]))
        assert np.all(np.abs(xs.mean(axis=1) - ms[g]) < 5 * sd / np.sqrt(M) + 0.02 * sd), (k, j, xs.mean(axis=1), ms[g])
        assert np.all(np.abs(xs.std(axis=1) - sd) < 0.06 * sd), (k, j, xs.std(axis=1), sd)
    ctx.close()
