"""SURVEY §8f item 4 on the GPU: per-recording parameter names (mixed effects, /root/reference/src/param_names_collections.jl:268-288)
and ensembles whose recordings differ in model and time grid (/root/reference/src/sampling_ensemble.jl:26-38)."""
import numpy as np
import pytest

import dmt_b200
from dmt_b200 import _lib, configs, hetero, param_names
from dmt_b200 import host as H
from harness import OracleEnsemble, rel_err

ParamNamesAllObs = param_names.ParamNamesAllObs

pytestmark = pytest.mark.gpu


def recordings_of(prob):
    return dict(theta=prob.theta, L=prob.L, Sigma=prob.Sigma, v=prob.v, x0=prob.x0, xbar=prob.xbar)


def test_mixed_effects_parameter_update_matches_oracle(orc, olib):
    """recordings 0-2 read their gamma from θ°[0], recordings 3-5 from θ°[1]; beta (θ°[2]) is shared"""
    M = 6
    prob = configs.make_problem("fhn", M, K=6, dt=0.005, seed=11, rho=0.9)
    se = H.SamplingEnsemble(prob.model, recordings_of(prob), (prob.n_pts, prob.tt), seed=3, two_sided_laws=True)
    se.init_paths()
    be = H.BlockEnsemble(se, [(0, prob.K - 1)], 0.9, 4)
    H.recompute_guiding_term(be)
    H.loglikhd(be)
    names = ["gamma_a", "gamma_b", "beta"]
    pdr = [[("gamma_a", "gamma"), ("beta", "beta")]] * 3 + [[("gamma_b", "gamma"), ("beta", "beta")]] * 3
    pa = ParamNamesAllObs.build(be, names, pdr)
    assert H.is_critical_update(be, pa)
    theta_o = np.array([prob.theta[2] * 1.05, prob.theta[2] * 0.93, prob.theta[3] + 0.05])
    H.set_proposal_law(be, theta_o, pa)                                   # critical_change from the tree
    assert np.allclose(se.theta_o[2], [theta_o[0]] * 3 + [theta_o[1]] * 3) and np.allclose(se.theta_o[3], theta_o[2])
    assert np.array_equal(se.theta_o[[0, 1, 4]], se.theta[[0, 1, 4]])     # everything else stays shared with the accepted law
    # oracle: the same per-recording proposal laws, proposal path recomputed from the accepted noise
    ora = OracleEnsemble(orc, olib, prob, seed=3)
    X, W = se.ctx.get_X(0), se.ctx.get_W(0)
    for s in (0, 1):
        ora.set_X(s, X); ora.set_W(s, W)
    ora.recompute_guiding_term(0, sides=(0, 1)); ora.loglikhd(0)
    for c, P in enumerate(ora.pairs):
        th_c = se.theta_o[:, c].copy()
        for k in range(prob.K):
            P.set_theta(th_c, side=1, k=k)
            Bm, beta, at = orc.linearise(olib, prob.model, th_c, prob.xbar[k, :, c])
            P.set_aux(k, Bm, beta, at, side=1)
    ora.recompute_guiding_term(0, sides=(1,)); ora.recompute_path(0, 1, 0)
    assert rel_err(se.ctx.get_ll(be.layout, 1), ora.ll(0, 1)) < 1e-9
    assert rel_err(se.ctx.get_X(1), ora.X(1)) < 1e-9
    ll_o = se.ctx.get_ll(be.layout, 1)[0]
    assert abs(ll_o[0] - ll_o[3]) > 1e-6                                  # the two groups really got different parameters
    se.ctx.close()


def _rec_list(prob, model):
    out = []
    for c in range(prob.M):
        out.append(dict(model=model, theta=prob.theta, L=prob.L, Sigma=prob.Sigma, v=prob.v[:, :, c], x0=prob.x0[:, c],
                        xbar=prob.xbar[:, :, c], tts=(prob.n_pts, prob.tt)))
    return out


def test_heterogeneous_ensemble_buckets_and_matches_homogeneous_runs():
    pa = configs.make_problem("fhn", 3, K=6, dt=0.005, seed=1, rho=0.8)
    pb = configs.make_problem("fhn", 2, K=8, dt=0.004, seed=2, rho=0.8)
    pc = configs.make_problem("lorenz", 2, K=6, dt=0.01, seed=3, rho=0.8)
    ra, rb, rc = _rec_list(pa, _lib.FHN), _rec_list(pb, _lib.FHN), _rec_list(pc, _lib.LORENZ)
    recs = [ra[0], rb[0], rc[0], ra[1], rc[1], rb[1], ra[2]]              # interleaved on purpose
    he = hetero.HeterogeneousEnsemble(recs, seed=17, two_sided_laws=False)
    assert he.members == [[0, 3, 6], [1, 5], [2, 4]] and he.where[5] == (1, 1) and he.num_recordings() == 7
    he.init_paths()
    hbe = hetero.HeterogeneousBlockEnsemble(he, lambda K: [(0, K // 2 - 1), (K // 2, K - 1)], 0.8, 3)
    assert [be.n_blocks for be in hbe.parts] == [2, 2, 2] and hbe.parts[1].ranges == [(0, 3), (4, 7)]
    for i in range(3):
        hetero.blocking_sweep(hbe, i)
        hetero.accept_reject_proposal_path(hbe, i)
    tot = hetero.fetch_ll(hbe)
    assert np.isfinite(tot) and abs(tot - sum(H.fetch_ll(be) for be in hbe.parts)) == 0.0
    rates = hetero.accpt_rate(hbe, (0, 2))
    assert len(rates) == 3 and all(r.shape == (2,) for r in rates)
    Xs = hetero.paths(he)
    assert [x.shape for x in Xs] == [(Xs[0].shape[0], 2), (Xs[1].shape[0], 2), (Xs[2].shape[0], 3)] + [x.shape for x in Xs[3:]]
    assert Xs[1].shape[0] != Xs[0].shape[0]                               # different grids
    # every bucket behaves exactly like a stand-alone homogeneous ensemble with the same random streams
    off = 0
    for prob, model, b in ((pa, _lib.FHN, 0), (pb, _lib.FHN, 1), (pc, _lib.LORENZ, 2)):
        se = H.SamplingEnsemble(model, recordings_of(prob), (prob.n_pts, prob.tt), seed=17, two_sided_laws=False, chain_offset_base=off)
        off += prob.M
        se.init_paths()
        K = prob.K
        be = H.BlockEnsemble(se, [(0, K // 2 - 1), (K // 2, K - 1)], 0.8, 3)
        for i in range(3):
            H.blocking_sweep(be, i)
            H.accept_reject_proposal_path(be, i)
        assert np.array_equal(se.ctx.get_X(0), he.buckets[b].ctx.get_X(0))
        assert np.array_equal(se.ctx.get_ll(be.layout, 0), he.buckets[b].ctx.get_ll(hbe.parts[b].layout, 0))
        se.ctx.close()
    he.close()


def test_heterogeneous_parameter_update_one_theta_for_all_buckets():
    """θ° = (sigma_fhn, sigma_lorenz): each bucket's ParamNamesAllObs picks its own entry"""
    pa = configs.make_problem("fhn", 2, K=4, dt=0.005, seed=5, rho=0.8)
    pc = configs.make_problem("lorenz", 3, K=4, dt=0.01, seed=6, rho=0.8)
    recs = _rec_list(pa, _lib.FHN) + _rec_list(pc, _lib.LORENZ)
    he = hetero.HeterogeneousEnsemble(recs, seed=2, two_sided_laws=True)
    he.init_paths()
    hbe = hetero.HeterogeneousBlockEnsemble(he, lambda K: [(0, K - 1)], 0.8, 2)
    hetero.recompute_guiding_term(hbe)
    hetero.loglikhd(hbe)
    names = ["sigma_fhn", "sigma_lorenz"]
    pdr = [[("sigma_fhn", "sigma")]] * 2 + [[("sigma_lorenz", "sigma")]] * 3
    pns = hetero.param_names(hbe, names, pdr)
    theta_o = np.array([pa.theta[4] * 1.1, pc.theta[3] * 0.9])
    hetero.set_proposal_law(hbe, theta_o, pns)
    assert np.allclose(he.buckets[0].theta_o[4], theta_o[0]) and np.allclose(he.buckets[1].theta_o[3], theta_o[1])
    ll, ll_o = hetero.fetch_ll(hbe), hetero.fetch_ll_o(hbe)
    assert np.isfinite(ll) and np.isfinite(ll_o) and ll != ll_o
    he.close()
