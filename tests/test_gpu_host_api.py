"""The reference's tutorials, run through the reference-shaped host API (dmt_b200.host) on the GPU and, call by call, on the
CPU oracle with the same counter-based random streams: accept/reject histories must be IDENTICAL, paths agree to 1e-9.

Tutorials mirrored (/root/reference/docs/src/tutorials/):
  biblock/smoothing.md:25-58                      simple_smoothing
  block_collection/inference_with_blocking.md     blocking sweep over alternating layouts
  block_ensemble/inference.md:44-75               path update + parameter update with swap_XX!/swap_PP!/save_ll!/swap_ll!
"""
import numpy as np
import pytest

import dmt_b200
from dmt_b200 import _lib, configs
from dmt_b200 import host as H
from harness import OracleEnsemble, rel_err

pytestmark = pytest.mark.gpu


def recordings_of(prob):
    return dict(theta=prob.theta, L=prob.L, Sigma=prob.Sigma, v=prob.v, x0=prob.x0, xbar=prob.xbar)


def test_simple_smoothing_tutorial(orc, olib):
    nit = 12
    prob = configs.make_problem("fhn", 33, K=10, dt=0.005, seed=100, rho=0.96)
    se = H.SamplingEnsemble(prob.model, recordings_of(prob), (prob.n_pts, prob.tt), seed=100, two_sided_laws=False)
    se.init_paths()
    bb = H.BlockEnsemble(se, [(0, prob.K - 1)], 0.96, nit)                # BiBlock(sp, 1:length(recording.obs), ρ, true, num_steps)
    H.recompute_guiding_term(bb, H.P_only)
    H.loglikhd(bb)
    ora = OracleEnsemble(orc, olib, prob, seed=100)
    X, W = se.ctx.get_X(0), se.ctx.get_W(0)
    for s in (0, 1):
        ora.set_X(s, X); ora.set_W(s, W)
    ora.recompute_guiding_term(0); ora.loglikhd(0)
    for i in range(nit):
        H.draw_proposal_path(bb, i)
        H.accept_reject_proposal_path(bb, i)
        ora.draw(0, i, layout_id=bb.layout); ora.accept(0, i, layout_id=bb.layout)
    acc_dev = se.ctx.get_accept_history(bb.layout, 0, nit - 1)
    assert 0.05 < acc_dev.mean() < 0.99
    # oracle history from its ll bookkeeping: replay equality through the final state
    assert rel_err(se.ctx.get_X(0), ora.X(0)) < 1e-9 and rel_err(se.ctx.get_W(0), ora.W(0)) < 1e-9
    assert rel_err(se.ctx.get_ll(bb.layout, 0), ora.ll(0, 0)) < 1e-9
    r = H.accpt_rate(bb, (0, nit - 1))
    assert r.shape == (1,) and abs(r[0] - acc_dev.mean()) < 1e-12
    lla = H.ll_of_accepted(bb, nit - 1)
    assert lla.shape == (33, 1) and np.allclose(lla[:, 0], se.ctx.get_ll(bb.layout, 0)[0], rtol=1e-12)
    v = bb.recordings[3].blocks[0]
    assert v.rho == 0.96 and v.accpt_history.shape == (nit,) and abs(v.b.ll - se.ctx.get_ll(bb.layout, 0)[0, 3]) == 0
    se.ctx.close()


def test_blocking_tutorial(orc, olib):
    K, nit = 12, 4
    layouts = [([(0, 2), (3, 8), (9, 11)], 0.9), ([(0, 5), (6, 11)], 0.9)]      # [[1:25,26:75,76:100],[1:50,51:100]] scaled down
    prob = configs.make_problem("lorenz", 24, K=K, dt=0.01, seed=3, layouts=layouts)
    se = H.SamplingEnsemble(prob.model, recordings_of(prob), (prob.n_pts, prob.tt), seed=8, two_sided_laws=False)
    se.init_paths()
    blocks = [H.BlockEnsemble(se, r, rho, nit) for r, rho in layouts]
    ora = OracleEnsemble(orc, olib, prob, seed=8)
    X, W = se.ctx.get_X(0), se.ctx.get_W(0)
    for s in (0, 1):
        ora.set_X(s, X); ora.set_W(s, W)
    for i in range(nit):
        for l, B in enumerate(blocks):
            H.set_obs(B); ora.set_artificial_obs(l)
            H.recompute_guiding_term(B, H.P_only); ora.recompute_guiding_term(l)
            H.find_W_for_X(B); ora.find_W_for_X(l)
            ora.set_W(0, se.ctx.get_W(0))
            H.loglikhd(B); ora.loglikhd(l)
            ora.set_ll(l, 0, se.ctx.get_ll(B.layout, 0))
            H.draw_proposal_path(B, i); ora.draw(l, i, layout_id=B.layout)
            H.accept_reject_proposal_path(B, i)
            acc_o, _ = ora.accept(l, i, layout_id=B.layout)
            assert np.array_equal(se.ctx.get_last_accept(B.layout), acc_o)
            assert rel_err(se.ctx.get_X(0), ora.X(0)) < 1e-9
    assert all((H.accpt_rate(B, (0, nit - 1)) > 0).all() for B in blocks)
    se.ctx.close()


def test_inference_tutorial_parameter_update(orc, olib):
    """block_ensemble/inference.md: alternate path imputation and a random-walk update of gamma shared by all recordings"""
    nit = 6
    prob = configs.make_problem("fhn", 16, K=6, dt=0.005, seed=5, rho=0.9)
    se = H.SamplingEnsemble(prob.model, recordings_of(prob), (prob.n_pts, prob.tt), seed=21, two_sided_laws=True)
    se.init_paths()
    be = H.BlockEnsemble(se, [(0, prob.K - 1)], 0.9, nit)
    H.recompute_guiding_term(be)
    H.loglikhd(be)
    ora = OracleEnsemble(orc, olib, prob, seed=21)
    X, W = se.ctx.get_X(0), se.ctx.get_W(0)
    for s in (0, 1):
        ora.set_X(s, X); ora.set_W(s, W)
    ora.recompute_guiding_term(0, sides=(0, 1)); ora.loglikhd(0)
    rng = np.random.default_rng(0)
    gamma = prob.theta[2]
    pnames = [(0, 2)]                                                   # θ°[1] -> :γ
    n_par_acc = 0
    for i in range(nit):
        H.draw_proposal_path(be, i); H.accept_reject_proposal_path(be, i)
        ora.draw(0, i, layout_id=be.layout); ora.accept(0, i, layout_id=be.layout)
        # parameter update
        g_o = gamma + 2 * 0.3 * (rng.random() - 0.5)
        H.set_proposal_law(be, [g_o], pnames, True)
        th_o = prob.theta.copy(); th_o[2] = g_o
        th_a = prob.theta.copy(); th_a[2] = gamma
        for c, P in enumerate(ora.pairs):
            for k in range(prob.K):   # equalize + set on the proposal side
                P.set_theta(th_o, side=1, k=k)
                Bm, beta, at = orc.linearise(olib, prob.model, th_o, prob.xbar[k, :, c])
                P.set_aux(k, Bm, beta, at, side=1)
        ora.recompute_guiding_term(0, sides=(1,)); ora.recompute_path(0, 1, 0)
        ll, ll_o = H.fetch_ll(be), H.fetch_ll_o(be)
        assert abs(ll - ora.ll(0, 0).sum()) < 1e-8 * abs(ll) and abs(ll_o - ora.ll(0, 1).sum()) < 1e-8 * max(1.0, abs(ll_o))
        accepted = rng.exponential() > -(ll_o - ll)                       # flat prior, symmetric kernel
        if accepted:
            H.swap_XX(be); H.swap_PP(be); ora.swap(0, 1 | 4)
            gamma = g_o
            n_par_acc += 1
        H.save_ll(be, i)
        if accepted:
            H.swap_ll(be); ora.swap(0, 8)
        assert np.allclose(se.theta[2], gamma)
        assert rel_err(se.ctx.get_X(0), ora.X(0)) < 1e-9 and rel_err(se.ctx.get_ll(be.layout, 0), ora.ll(0, 0)) < 1e-9
    se.ctx.close()


def test_checkpoint_resume_is_bit_exact(tmp_path):
    """save_state / load_state: a resumed run continues exactly like the uninterrupted one"""
    K = 8
    layouts = [([(0, 2), (3, 5), (6, 7)], 0.8), ([(0, 3), (4, 7)], 0.7)]
    prob = configs.make_problem("lorenz", 40, K=K, dt=0.01, seed=9, layouts=layouts)

    def fresh():
        se = H.SamplingEnsemble(prob.model, recordings_of(prob), (prob.n_pts, prob.tt), seed=5, two_sided_laws=False)
        se.init_paths()
        return se, [H.BlockEnsemble(se, r, rho, 0) for r, rho in layouts]

    def sweeps(bes, its):
        for i in its:
            for be in bes:
                H.blocking_sweep(be, i)
                H.accept_reject_proposal_path(be, i)

    se, bes = fresh()
    sweeps(bes, range(3))
    ck = str(tmp_path / "state.npz")
    H.save_state(se, ck, bes)
    sweeps(bes, range(3, 6))
    Xa, Wa, lla = se.ctx.get_X(0), se.ctx.get_W(0), [se.ctx.get_ll(be.layout, 0) for be in bes]
    se.ctx.close()
    se2, bes2 = fresh()
    H.load_state(se2, ck, bes2)
    sweeps(bes2, range(3, 6))
    assert np.array_equal(se2.ctx.get_X(0), Xa) and np.array_equal(se2.ctx.get_W(0), Wa)
    assert all(np.array_equal(se2.ctx.get_ll(be.layout, 0), l) for be, l in zip(bes2, lla))
    se2.ctx.close()


def test_history_setters_and_recompute_path_exports():
    """set_ll!, set_accepted!, recompute_path! (exports of src/DiffusionMCMCTools.jl:38-45) through the host API"""
    nit = 3
    prob = configs.make_problem("fhn", 9, K=4, dt=0.005, seed=2, rho=0.9)
    se = H.SamplingEnsemble(prob.model, recordings_of(prob), (prob.n_pts, prob.tt), seed=4, two_sided_laws=True)
    se.init_paths()
    be = H.BlockEnsemble(se, [(0, 1), (2, 3)], 0.9, nit)
    H.set_obs(be); H.recompute_guiding_term(be); H.find_W_for_X(be); H.loglikhd(be)
    H.set_accepted(be, 1, True)
    H.set_accepted(be, 2, np.arange(2 * 9).reshape(2, 9) % 2 == 0)
    H.set_ll(be, 1, -3.5)
    H.set_ll(be, 2, np.full((2, 9), 7.25), side=1)
    acc = se.ctx.get_accept_history(be.layout, 0, nit - 1)
    assert not acc[0].any() and acc[1].all() and np.array_equal(acc[2], np.arange(18).reshape(2, 9) % 2 == 0)
    assert np.all(se.ctx.get_ll_history(be.layout, 0, 1, 1) == -3.5) and np.all(se.ctx.get_ll_history(be.layout, 1, 2, 2) == 7.25)
    assert np.allclose(H.accpt_rate(be, (1, 2)), [(9 + 5) / 18.0, (9 + 4) / 18.0])
    # recompute_path!(b°, b.WW): same law on both sides => the proposal path reproduces the accepted one and its ll
    H.recompute_path(be)
    assert rel_err(se.ctx.get_X(1), se.ctx.get_X(0)) < 1e-12
    assert rel_err(se.ctx.get_ll(be.layout, 1), se.ctx.get_ll(be.layout, 0)) < 1e-10
    with pytest.raises(dmt_b200.DmtError):
        H.set_accepted(be, nit, True)
    se.ctx.close()


def test_path_saver_snapshots_are_ordered_and_asynchronous():
    """PathSaver / dmt_snapshot_paths_async: a snapshot holds the paths as they were when it was queued, even though the next
    sweeps are enqueued before anyone waits for it (tutorial: save a path every few hundred iterations)"""
    layouts = [([(0, 2), (3, 5)], 0.5), ([(0, 5)], 0.5)]
    prob = configs.make_problem("lorenz", 70, K=6, dt=0.01, seed=8, layouts=layouts)
    se = H.SamplingEnsemble(prob.model, recordings_of(prob), (prob.n_pts, prob.tt), seed=6, two_sided_laws=False)
    se.init_paths()
    bes = [H.BlockEnsemble(se, r, rho, 0) for r, rho in layouts]
    chains = [3, 69, 0, 41]
    saver = H.PathSaver(se, chains, every=2)
    truth = {}
    for i in range(6):
        for be in bes:
            H.blocking_sweep(be, i)
            H.accept_reject_proposal_path(be, i)
        took = saver(i)
        assert took == (i % 2 == 0)
        if took:                              # a synchronous copy right after the snapshot was queued, for comparison
            truth[i] = se.ctx.get_X(0)[:, :, chains].copy()
    got = saver.paths()
    assert [i for i, _ in got] == [0, 2, 4]
    for i, X in got:
        assert X.shape == (se.ctx.NP, 3, 4) and np.array_equal(X, truth[i])
    assert not np.array_equal(got[0][1], got[2][1])          # the chain moved in between
    with pytest.raises(dmt_b200.DmtError):
        se.ctx.snapshot_paths_async([70], np.empty((se.ctx.NP, 3, 1)))
    se.ctx.close()


def test_set_proposal_law_escalates_critical_change_when_the_proposal_laws_had_to_be_equalised():
    """src/biblock.jl:362-363: `GP.equalize_*!(bb) && (critical_change = true)`.  After a rejected critical update b°'s records
    AND guiding term belong to the rejected θ°.  A following update flagged non-critical must still recompute b°'s guiding term,
    otherwise ll° is computed with (B, β, ã) of one parameter and (H, F, c) of another."""
    prob = configs.make_problem("fhn", 12, K=4, dt=0.005, seed=3, rho=0.9)
    se = H.SamplingEnsemble(prob.model, recordings_of(prob), (prob.n_pts, prob.tt), seed=1, two_sided_laws=True)
    se.init_paths()
    be = H.BlockEnsemble(se, [(0, prob.K - 1)], 0.9, 0)
    H.recompute_guiding_term(be); H.loglikhd(be)
    ll = se.ctx.get_ll(be.layout, 0).copy()
    H.set_proposal_law(be, [prob.theta[2] * 1.2], [(0, 2)], True)          # critical update of gamma ... rejected (no swap)
    assert np.abs(se.ctx.get_ll(be.layout, 1) - ll).max() > 1e-6
    # now "update" gamma back to the accepted value, flagged NON-critical: equalisation changes b°, so K1 must run on b° anyway
    H.set_proposal_law(be, [prob.theta[2]], [(0, 2)], False)
    assert rel_err(se.ctx.get_ll(be.layout, 1), ll) < 1e-12
    assert rel_err(se.ctx.get_X(1), se.ctx.get_X(0)) < 1e-12
    for k in range(prob.K):
        Ha, Fa, ca = se.ctx.get_guiding_term(k, 0); Ho, Fo, co = se.ctx.get_guiding_term(k, 1)
        assert np.array_equal(Ha[:-1], Ho[:-1]) and np.array_equal(Fa[:-1], Fo[:-1]) and np.array_equal(ca[0], co[0])
    # and with nothing to equalise a non-critical update stays non-critical (no K1): the equalize call reports no change
    assert se.ctx.equalize_laws(3) is False
    se.ctx.close()


def test_observation_parameter_update_reaches_the_proposal_laws(orc, olib):
    """θ° entries that are parameters of the observations (updt_obs, src/param_names_collections.jl:125-141): the proposal-side
    (L, Σ, v) must be set AFTER the equalisation (src/biblock.jl:362-367), otherwise equalize_obs_params! wipes them out."""
    from dmt_b200 import param_names
    ParamNamesAllObs = param_names.ParamNamesAllObs
    M = 5
    prob = configs.make_problem("fhn", M, K=4, dt=0.005, seed=13, rho=0.9)
    se = H.SamplingEnsemble(prob.model, recordings_of(prob), (prob.n_pts, prob.tt), seed=2, two_sided_laws=True)
    se.init_paths()
    be = H.BlockEnsemble(se, [(0, prob.K - 1)], 0.9, 0)
    H.recompute_guiding_term(be); H.loglikhd(be)
    names = ["obs_shift"]
    pdr = [[] for _ in range(M)]
    odr = [[(("obs_shift", 0),) for _ in range(prob.K)] for _ in range(M)]       # every observation's θ[0] := obs_shift
    pa = ParamNamesAllObs.build(be, names, pdr, odr)
    assert pa.is_critical()
    seen = {}

    def hook(updates, theta_o):                                                  # obs.θ[0] shifts the observed value
        seen["updates"] = updates
        return prob.L, prob.Sigma, prob.v + theta_o[0]
    se.obs_param_hook = hook
    H.set_proposal_law(be, np.array([0.05]), pa)
    assert len(seen["updates"]) == M and all(set(u) == set(range(prob.K)) for u in seen["updates"])
    ora = OracleEnsemble(orc, olib, prob, seed=2)
    X, W = se.ctx.get_X(0), se.ctx.get_W(0)
    for s in (0, 1):
        ora.set_X(s, X); ora.set_W(s, W)
    for c, P in enumerate(ora.pairs):
        for k in range(prob.K):
            P.set_obs(k, prob.L, prob.Sigma, prob.v[k, :, c] + 0.05, side=1)
    ora.recompute_guiding_term(0, sides=(0, 1)); ora.loglikhd(0); ora.recompute_path(0, 1, 0)
    for k in range(prob.K):
        H1, F1, c1 = se.ctx.get_guiding_term(k, 1); Ho, Fo, co = ora.guiding(k, 1)
        assert rel_err(F1[:-1], Fo[:-1]) < 1e-10 and rel_err(c1[0], co[0]) < 1e-10
        F0 = se.ctx.get_guiding_term(k, 0)[1]
        assert np.abs(F1[:-1] - F0[:-1]).max() > 1e-3                            # the proposal side really saw the new observations
    assert rel_err(se.ctx.get_ll(be.layout, 1), ora.ll(0, 1)) < 1e-9 and rel_err(se.ctx.get_X(1), ora.X(1)) < 1e-9
    se.ctx.close()


def test_checkpoint_resume_with_parameter_updates_is_bit_exact(tmp_path):
    """save_state / load_state restore θ, θ°, the laws and the histories: an inference run (path update + γ update with
    swap_XX!/swap_PP!) resumed from a checkpoint continues exactly like the uninterrupted one."""
    nit = 8
    prob = configs.make_problem("fhn", 10, K=5, dt=0.005, seed=4, rho=0.9)

    def fresh():
        se = H.SamplingEnsemble(prob.model, recordings_of(prob), (prob.n_pts, prob.tt), seed=9, two_sided_laws=True)
        se.init_paths()
        be = H.BlockEnsemble(se, [(0, prob.K - 1)], 0.9, nit)
        H.recompute_guiding_term(be); H.loglikhd(be)
        return se, be

    def run(se, be, its, gamma):
        for i in its:
            rng = np.random.default_rng(1000 + i)                        # the host's RNG state is the caller's to checkpoint
            H.draw_proposal_path(be, i); H.accept_reject_proposal_path(be, i)
            g_o = gamma + 0.2 * (rng.random() - 0.5)
            H.set_proposal_law(be, [g_o], [(0, 2)], True)
            ll, ll_o = H.fetch_ll(be), H.fetch_ll_o(be)
            acc = rng.exponential() > -(ll_o - ll)
            if acc:
                H.swap_XX(be); H.swap_PP(be); gamma = g_o
            H.save_ll(be, i)
            if acc:
                H.swap_ll(be)
        return gamma

    se, be = fresh()
    g = run(se, be, range(4), prob.theta[2])
    assert g != prob.theta[2]                                            # at least one γ update was accepted before the checkpoint
    ck = str(tmp_path / "state.npz")
    H.save_state(se, ck, [be])
    g_end = run(se, be, range(4, nit), g)
    want = (se.ctx.get_X(0), se.ctx.get_X(1), se.ctx.get_W(0), se.ctx.get_ll(be.layout, 0), se.ctx.get_ll_history(be.layout, 0, 0, nit - 1),
            se.ctx.get_accept_history(be.layout, 0, nit - 1), se.theta.copy())
    se.ctx.close()
    se2, be2 = fresh()
    H.load_state(se2, ck, [be2])
    assert np.allclose(se2.theta[2], g)
    g_end2 = run(se2, be2, range(4, nit), g)
    got = (se2.ctx.get_X(0), se2.ctx.get_X(1), se2.ctx.get_W(0), se2.ctx.get_ll(be2.layout, 0), se2.ctx.get_ll_history(be2.layout, 0, 0, nit - 1),
           se2.ctx.get_accept_history(be2.layout, 0, nit - 1), se2.theta.copy())
    assert g_end2 == g_end
    for a, b in zip(want, got):
        assert np.array_equal(a, b, equal_nan=True)
    se2.ctx.close()


def test_biblock_and_block_collection_views():
    """be.recordings[r] (BlockCollection, src/block_collection.jl:17-36) and .blocks[b] (BiBlock, src/biblock.jl:17-62) as views of the
    device ensemble: XX / WW in the reference's containers (per-interval time grids, cumulative Wiener paths), the per-block and
    per-recording swaps (src/biblock.jl:148-209, src/block_collection.jl:84-118) and the history setters / readers."""
    K, nit = 6, 3
    prob = configs.make_problem("lorenz", 9, K=K, dt=0.01, seed=5, layouts=[([(0, 1), (2, 5)], 0.8)])
    se = H.SamplingEnsemble(prob.model, recordings_of(prob), (prob.n_pts, prob.tt), seed=4, two_sided_laws=True)
    se.init_paths()
    be = H.BlockEnsemble(se, [(0, 1), (2, 5)], 0.8, nit)
    ctx = se.ctx
    H.set_obs(be); H.recompute_guiding_term(be, H.P_only); H.find_W_for_X(be); H.loglikhd(be)
    for i in range(nit):
        H.draw_proposal_path(be, i); H.accept_reject_proposal_path(be, i)
    H.draw_proposal_path(be, nit)                                # leave a proposal in place: both sides differ now
    X0, X1, W0, W1 = ctx.get_X(0), ctx.get_X(1), ctx.get_W(0), ctx.get_W(1)
    ll = ctx.get_ll(be.layout, 0), ctx.get_ll(be.layout, 1)
    rec, blk = 4, 1
    bc = be.recordings[rec]; bb = bc.blocks[blk]
    # ---- containers
    XX, WW = bb.b.XX, bb.b.WW
    assert len(XX) == 4 and len(WW) == 4                         # intervals 2..5
    for j, k in enumerate(range(2, 6)):
        t, x = XX[j]
        assert np.array_equal(t, prob.tt[ctx.pt0[k]:ctx.pt0[k + 1]]) and np.array_equal(x, X0[ctx.pt0[k]:ctx.pt0[k + 1], :, rec])
        tw, w = WW[j]
        assert w.shape == (ctx.n_pts[k], prob.dw) and np.all(w[0] == 0.0) and np.array_equal(tw, t)
        assert np.allclose(np.diff(w, axis=0), W0[ctx.step0[k]:ctx.step0[k + 1], :, rec], rtol=0, atol=1e-15)
    assert np.array_equal(bb.b_o.XX[0][1], X1[ctx.pt0[2]:ctx.pt0[3], :, rec])
    assert bb.b.ll == ll[0][blk, rec] and bb.b_o.ll == ll[1][blk, rec]
    assert bc.fetch_ll() == ll[0][:, rec].sum() and bc.fetch_ll_o() == ll[1][:, rec].sum()
    # ---- BiBlock-level swaps touch exactly one block of one recording
    bb.swap_XX()
    Y0, Y1 = ctx.get_X(0), ctx.get_X(1)
    sl = slice(ctx.pt0[2], ctx.pt0[6])
    assert np.array_equal(Y0[sl, :, rec], X1[sl, :, rec]) and np.array_equal(Y1[sl, :, rec], X0[sl, :, rec])
    other = np.ones(X0.shape, bool); other[sl, :, rec] = False
    assert np.array_equal(Y0[other], X0[other]) and np.array_equal(Y1[other], X1[other])
    bb.swap_XX()
    assert np.array_equal(ctx.get_X(0), X0)
    bb.swap_paths(); bb.swap_ll()
    ws = slice(ctx.step0[2], ctx.step0[6])
    assert np.array_equal(ctx.get_W(0)[ws, :, rec], W1[ws, :, rec]) and np.array_equal(ctx.get_X(0)[sl, :, rec], X1[sl, :, rec])
    assert ctx.get_ll(be.layout, 0)[blk, rec] == ll[1][blk, rec] and ctx.get_ll(be.layout, 0)[0, rec] == ll[0][0, rec]
    bb.swap_paths(); bb.swap_ll()
    # ---- BlockCollection-level swaps: every block of one recording
    bc.swap_XX()
    Y0 = ctx.get_X(0)
    assert np.array_equal(Y0[:, :, rec], X1[:, :, rec]) and np.array_equal(np.delete(Y0, rec, axis=2), np.delete(X0, rec, axis=2))
    bc.swap_XX()
    # ---- histories
    bb.set_accepted(1, True); bb.set_ll(1, -3.5, side=1)
    assert bb.accpt_history[1] and bb.ll_of_accepted(1) == -3.5
    bb.set_accepted(1, False); bb.set_ll(1, -7.25, side=0)
    assert not bb.accpt_history[1] and bb.ll_of_accepted(1) == -7.25
    assert bb.accpt_rate((0, nit - 1)) == np.mean(bb.accpt_history)
    assert bc.accpt_rate((0, nit - 1)) == [b.accpt_rate((0, nit - 1)) for b in bc.blocks]
    assert np.allclose(H.ll_of_accepted(be, 1)[rec], bc.ll_of_accepted(1))
    ctx.close()


def test_history_streaming_is_asynchronous_and_ordered():
    """dmt_histories_async / host.HistoryStreamer: chunks of ll_history / accpt_history leave on the copy stream while later sweeps
    run; what arrives equals what the synchronous getters return at the end, row by row (ordering: a chunk holds exactly the values
    the rows had when it was queued, later iterations never leak into it)."""
    nit = 23
    prob = configs.make_problem("lv", 50, K=5, dt=0.01, seed=9, rho=0.9)
    se = H.SamplingEnsemble(prob.model, recordings_of(prob), (prob.n_pts, prob.tt), seed=2, two_sided_laws=False)
    se.init_paths()
    be = H.BlockEnsemble(se, [(0, prob.K - 1)], 0.9, nit)
    H.recompute_guiding_term(be, H.P_only); H.loglikhd(be)
    hs = H.HistoryStreamer(be, every=5)
    shipped = 0
    for i in range(nit):
        H.draw_proposal_path(be, i); H.accept_reject_proposal_path(be, i)
        shipped += bool(hs(i))
    hs(nit - 1, flush=True)
    ll, acc = hs.collect()
    assert shipped == 4 and ll.shape == (nit, 2, 1, 50) and acc.shape == (nit, 1, 50)
    ctx = se.ctx
    assert np.array_equal(acc, ctx.get_accept_history(be.layout, 0, nit - 1))
    assert np.array_equal(ll[:, 0], ctx.get_ll_history(be.layout, 0, 0, nit - 1)) and np.array_equal(ll[:, 1], ctx.get_ll_history(be.layout, 1, 0, nit - 1))
    assert acc.any()
    with pytest.raises(dmt_b200.DmtError):
        ctx.histories_async(be.layout, 0, nit, np.empty((nit + 1, 2, 1, 50)), None)        # beyond ll_hist_len
    ctx.close()
