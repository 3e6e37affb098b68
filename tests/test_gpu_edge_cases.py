"""Edge cases of the path (SURVEY §8c "cover the edge cases the reference tests": the reference tests none, so these follow
its container semantics): ragged time grids, a single interval / single chain, general pset maps, per-block rho,
non-default block flags, argument errors."""
import numpy as np
import pytest

import dmt_b200
from dmt_b200 import _lib, configs
from harness import OracleEnsemble, compare_guiding, make_ctx, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-10


def ragged_problem(name, M, n_steps, seed=3, P=None, layouts=None):
    """intervals with different numbers of steps (every partial-tile shape)"""
    K = len(n_steps)
    base = configs.make_problem(name, M, P=P, K=K, seed=seed, layouts=layouts, rho=0.6)
    obs_dt = 0.01 if name == "jr" else 0.1
    grids = [configs.tau_grid(k * obs_dt, (k + 1) * obs_dt, obs_dt / n_steps[k]) for k in range(K)]
    base.n_pts = np.array([len(g) for g in grids], dtype=np.int32)
    base.tt = np.concatenate(grids)
    return base


@pytest.mark.parametrize("name", ["lorenz", "fhn", "prok"])
def test_ragged_time_grid(orc, olib, name):
    n_steps = [9, 10, 11, 12, 13, 7, 8]   # tile remainders 1, 2, 3, 0, 1, 3, 0
    layouts = [([(0, 2), (3, 4), (5, 6)], [0.5, 0.6, 0.7])]
    prob = ragged_problem(name, 35, n_steps, layouts=layouts)
    ctx = make_ctx(prob, seed=4, n_layouts=2)
    ora = OracleEnsemble(orc, olib, prob, seed=4)
    K = prob.K
    ctx.set_blocks(1, [(0, K - 1)], 0.0)
    ctx.recompute_guiding_term(1, _lib.P_ONLY)
    assert ctx.init_paths(1, 11, 50) == 0
    X, W = ctx.get_X(0), ctx.get_W(0)
    assert X.shape[0] == sum(n_steps) + K and W.shape[0] == sum(n_steps)
    for s in (0, 1):
        ora.set_X(s, X); ora.set_W(s, W)
    ctx.set_artificial_obs(0); ora.set_artificial_obs(0)
    ctx.recompute_guiding_term(0, _lib.P_ONLY); ora.recompute_guiding_term(0)
    for (i0, i1) in layouts[0][0]:
        for k in range(i0, i1 + 1):
            compare_guiding(ctx, ora, k, 0, 1 if (k == i1 and i1 != K - 1) else 0, tol=TOL)
    ctx.find_W_loglikhd_draw(0, 2)
    ora.find_W_for_X(0); ora.set_W(0, ctx.get_W(0)); ora.loglikhd(0); ora.set_ll(0, 0, ctx.get_ll(0, 0)); ok_o = ora.draw(0, 2)
    assert np.array_equal(ctx.get_success(0), ok_o)
    assert rel_err(ctx.get_ll(0, 1), ora.ll(0, 1)) < 1e-9
    ctx.accept_reject_path(0, 2); acc_o, _ = ora.accept(0, 2)
    assert np.array_equal(ctx.get_last_accept(0), acc_o)
    assert rel_err(ctx.get_X(0), ora.X(0)) < 1e-9 and rel_err(ctx.get_W(0), ora.W(0)) < 1e-9
    ctx.close()


def test_single_chain_single_interval(orc, olib):
    prob = ragged_problem("fhn", 1, [7])
    ctx = make_ctx(prob, seed=9)
    ora = OracleEnsemble(orc, olib, prob, seed=9)
    ctx.recompute_guiding_term(0, _lib.P_ONLY); ora.recompute_guiding_term(0)
    compare_guiding(ctx, ora, 0, tol=TOL)
    rng = np.random.default_rng(0)
    W = 0.1 * rng.normal(size=(7, 1, 1))
    ctx.set_W(W, 0); ora.set_W(0, W)
    ctx.recompute_path(0, 0, 0); ora.recompute_path(0, 0, 0)
    assert rel_err(ctx.get_X(0), ora.X(0)) < TOL and rel_err(ctx.get_ll(0, 0), ora.ll(0, 0)) < TOL
    Z = rng.normal(size=(7, 1, 1))
    ctx.draw_proposal_path(0, 0, Z); ora.draw(0, 0, Z)
    assert rel_err(ctx.get_X(1), ora.X(1)) < TOL and rel_err(ctx.get_ll(0, 1), ora.ll(0, 1)) < TOL
    ctx.close()


def test_general_pset_map(orc, olib):
    """3 data sets shared by 50 chains through an arbitrary pset_of_chain (gathered guiding-term sectors)"""
    prob = configs.make_problem("lorenz", 50, P=3, K=3, dt=0.01, seed=2)
    prob.pset_of_chain = np.array([(7 * c + c // 3) % 3 for c in range(50)], dtype=np.int32)
    prob.x0 = np.repeat(np.array(configs.X0[configs.LORENZ])[:, None], 50, axis=1)
    ctx = make_ctx(prob, seed=6)
    ora = OracleEnsemble(orc, olib, prob, seed=6)
    ctx.recompute_guiding_term(0, _lib.P_ONLY); ora.recompute_guiding_term(0)
    for k in range(3):
        compare_guiding(ctx, ora, k, tol=TOL)
    assert ctx.init_paths(0, 5, 20) == 0
    X, W = ctx.get_X(0), ctx.get_W(0)
    for s in (0, 1):
        ora.set_X(s, X); ora.set_W(s, W)
    ctx.loglikhd(0, 0, 0); ora.loglikhd(0, 0, 0)
    assert rel_err(ctx.get_ll(0, 0), ora.ll(0, 0)) < TOL
    ctx.draw_proposal_path(0, 1); ora.draw(0, 1)
    assert rel_err(ctx.get_X(1), ora.X(1)) < TOL and rel_err(ctx.get_ll(0, 1), ora.ll(0, 1)) < TOL
    # blocking needs one parameter/data set per recording (the artificial observation is per recording)
    with pytest.raises(dmt_b200.DmtError) as e:
        ctx.set_blocks(0, [(0, 1), (2, 2)], 0.5)
    assert e.value.code == 4
    ctx.close()


def test_manual_block_flags_and_argument_errors():
    prob = configs.make_problem("lv", 8, K=4, dt=0.01, seed=1)
    ctx = make_ctx(prob, seed=1, n_layouts=2)
    # BiBlock(sp, range, rho, last_block=false) on a range that does NOT end the recording: allowed, needs >= 2 intervals
    ctx.set_blocks(1, [(0, 1)], 0.3, last=[0])
    with pytest.raises(dmt_b200.DmtError) as e:
        ctx.set_blocks(1, [(2, 2)], 0.3, last=[0])          # non-terminal single-interval block: the reference indexes PP[1]
    assert e.value.code == 1 and "2 intervals" in str(e.value)
    for bad in ([(0, 4)], [(-1, 2)], [(3, 2)]):
        with pytest.raises(dmt_b200.DmtError):
            ctx.set_blocks(1, bad, 0.1)
    with pytest.raises(dmt_b200.DmtError):
        ctx.set_blocks(1, [(0, 3)], 1.5)                    # |rho| > 1
    with pytest.raises(dmt_b200.DmtError) as e:
        ctx.loglikhd(0, 1, 0)                               # proposal laws were not allocated
    assert e.value.code == 3
    with pytest.raises(dmt_b200.DmtError):
        ctx.draw_proposal_path(7, 0)                        # unknown layout
    with pytest.raises(dmt_b200.DmtError):
        ctx.get_accept_history(0, 0, 0)                     # ll_hist_len == 0
    with pytest.raises((dmt_b200.DmtError, ValueError)):
        dmt_b200.Ctx(99, prob.n_pts, prob.tt, 4, obs_dim=2)  # unknown model
    with pytest.raises(dmt_b200.DmtError):
        dmt_b200.Ctx(_lib.LV, prob.n_pts, prob.tt[::-1].copy(), 4, obs_dim=2)  # decreasing grid
    ctx.close()


def test_histories_and_accept_rate_bookkeeping():
    nit = 7
    prob = configs.make_problem("lorenz", 33, K=4, dt=0.01, seed=5, layouts=[([(0, 1), (2, 3)], [0.2, 0.95])])
    ctx = make_ctx(prob, seed=2, ll_hist_len=nit, n_layouts=2)
    ctx.set_blocks(1, [(0, 3)], 0.0)
    ctx.recompute_guiding_term(1, _lib.P_ONLY)
    assert ctx.init_paths(1, 0, 20) == 0
    lls = []
    for i in range(nit):
        ctx.blocking_sweep(0, i)
        ll, llo = ctx.get_ll(0, 0).copy(), ctx.get_ll(0, 1).copy()
        ctx.accept_reject_path(0, i)
        acc = ctx.get_last_accept(0)
        # save_ll! happens BEFORE swap_ll! (src/biblock.jl:125-126)
        assert np.array_equal(ctx.get_ll_history(0, 0, i, i)[0], ll) and np.array_equal(ctx.get_ll_history(0, 1, i, i)[0], llo, equal_nan=True)
        assert np.array_equal(ctx.get_accept_history(0, i, i)[0], acc)
        assert np.array_equal(ctx.get_ll(0, 0), np.where(acc, llo, ll))
        lls.append(acc)
    hist = ctx.get_accept_history(0, 0, nit - 1)
    assert np.array_equal(hist, np.stack(lls))
    assert np.array_equal(ctx.accept_counts(0, 2, 5), hist[2:6].sum(axis=(0, 2)))
    # rho close to 1 (block 1) must accept more often than rho small (block 0)
    assert hist[:, 1].mean() > hist[:, 0].mean()
    st = ctx.allreduce_stats(0)
    assert st[2:].tolist() == hist[-1].sum(axis=1).tolist() and abs(st[0] - ctx.get_ll(0, 0).sum()) < 1e-9 * abs(st[0])
    # the peer-memory all-reduce kernel with a world of one rank is the identity (several calls: sequence numbers, both parities)
    with pytest.raises(dmt_b200.DmtError):
        ctx.p2p_init(1, 0, np.zeros((1, 64), np.uint8))              # export first
    h = ctx.p2p_export()
    assert h.shape == (64,) and h.any()
    ctx.p2p_init(1, 0, h[None, :])
    for _ in range(3):
        assert np.array_equal(ctx.allreduce_stats(0), st)
    with pytest.raises(dmt_b200.DmtError):
        ctx.p2p_init(17, 0, np.zeros((17, 64), np.uint8))
    ctx.close()


def test_relinearise_on_device_equals_fresh_upload():
    """dmt_set_aux_linearised(xbar = NULL): a parameter update re-linearises at the stored points instead of uploading them again"""
    prob = configs.make_problem("lorenz", 12, K=4, dt=0.01, seed=3, rho=0.5)
    a, b = make_ctx(prob, seed=1), make_ctx(prob, seed=1)
    th = np.repeat(prob.theta[:, None], prob.P, axis=1)
    th[1] *= 1.0 + 0.02 * np.arange(prob.P)                 # a different rho for every recording
    for ctx, xb in ((a, prob.xbar), (b, None)):
        ctx.set_params(th, side=0, stores=3)
        ctx.set_aux_linearised(xb, side=0, store=_lib.STORE_PP)
        ctx.recompute_guiding_term(0, _lib.P_ONLY)
    for k in range(prob.K):
        for x, y in zip(a.get_guiding_term(k, 0, 0), b.get_guiding_term(k, 0, 0)):
            assert np.array_equal(x, y, equal_nan=True)
    H0, _, _ = make_ctx_guiding(prob)
    assert not np.array_equal(H0[:-1], a.get_guiding_term(1, 0, 0)[0][:-1])         # the parameter change did reach the auxiliary law
    c = dmt_b200.Ctx(prob.model, prob.n_pts, prob.tt, prob.M, prob.P, obs_dim=prob.m)
    with pytest.raises(dmt_b200.DmtError):
        c.set_aux_linearised(None)                                          # nothing stored yet
    for ctx in (a, b, c):
        ctx.close()


def make_ctx_guiding(prob):
    ctx = make_ctx(prob, seed=1)
    ctx.recompute_guiding_term(0, _lib.P_ONLY)
    out = ctx.get_guiding_term(1, 0, 0)
    ctx.close()
    return out
