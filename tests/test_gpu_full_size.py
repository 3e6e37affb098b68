"""BASELINE config C3 AT FULL SIZE (Lorenz, 4096 chains, 200 x 100 EM steps, 10/11 staggered blocks): the oracle cannot run
8e7 steps per sweep in test time, so the whole ensemble is checked through size-independent properties, and a slice of it
(three chains out of 4096, with their global random streams) is replayed call by call on the oracle."""
import copy

import numpy as np
import pytest

import dmt_b200
from dmt_b200 import _lib, configs
from harness import OracleEnsemble, make_ctx, note_err, rel_err

pytestmark = [pytest.mark.gpu, pytest.mark.own_lanes]

SLICE = slice(2501, 2504)


@pytest.fixture(scope="module")
def full():
    prob = configs.named_config("c3", seed=123)
    assert (prob.M, prob.P, prob.K, prob.steps_per_chain) == (4096, 4096, 200, 20000)
    ctx = make_ctx(prob, seed=123, n_layouts=3)
    ctx.set_blocks(2, [(0, prob.K - 1)], 0.0)
    ctx.recompute_guiding_term(2, _lib.P_ONLY)
    assert ctx.init_paths(2, 0, 50) == 0
    yield prob, ctx
    ctx.close()


def test_full_size_properties(full):
    prob, ctx = full
    X0, W0 = ctx.get_X(0), ctx.get_W(0)
    assert np.isfinite(X0).all() and np.isfinite(W0).all()
    # (1) K5 o K2 = identity: the noise recovered from the initial paths under the law that made them is the noise that made them
    ctx.find_W_for_X(2)
    assert rel_err(ctx.get_W(0), W0) < 1e-8
    ctx.set_W(W0, 0)
    # (2) checksum of checksums: fetch_ll (fixed-order device tree) == sum of the per-(block, chain) values
    ctx.loglikhd(2, 0, 0)
    ll = ctx.get_ll(2, 0)
    tot, per_block = ctx.fetch_ll(2, 0)
    assert abs(tot - ll.sum()) < 1e-10 * abs(tot) and np.allclose(per_block, ll.sum(axis=1), rtol=1e-10)
    # (3) rho = 1: the pCN proposal is the accepted path (same noise, same law), and so is its log-likelihood
    ctx.set_rho(2, [1.0])
    ctx.draw_proposal_path(2, 5)
    assert np.array_equal(ctx.get_W(1), W0)
    assert rel_err(ctx.get_X(1), X0) < 1e-11 and rel_err(ctx.get_ll(2, 1), ll) < 1e-10
    # (4) one blocking sweep per layout: fused pass == the three separate passes; cached K1 == full backward filter
    for lay in (0, 1):
        ctx.set_artificial_obs(lay)
        ctx.recompute_guiding_term(lay, _lib.P_ONLY)
        ctx.find_W_for_X(lay); ctx.loglikhd(lay, 0, 0); ctx.draw_proposal_path(lay, 7)
        Wa, Xp, Wp = ctx.get_W(0), ctx.get_X(1), ctx.get_W(1)
        lla, llp, ok = ctx.get_ll(lay, 0), ctx.get_ll(lay, 1), ctx.get_success(lay)
        assert ok.mean() > 0.999
        ctx.find_W_loglikhd_draw(lay, 7)
        assert np.array_equal(ctx.get_success(lay), ok)
        assert rel_err(ctx.get_W(0), Wa) < 1e-10 and rel_err(ctx.get_ll(lay, 0), lla) < 1e-10
        good = ok.all(axis=0)
        assert rel_err(ctx.get_X(1)[:, :, good], Xp[:, :, good]) < 1e-9 and rel_err(ctx.get_W(1)[:, :, good], Wp[:, :, good]) < 1e-9
        assert rel_err(ctx.get_ll(lay, 1)[:, good], llp[:, good]) < 1e-9
        # determinism: the same call again gives the same bits
        X1 = ctx.get_X(1)
        ctx.find_W_loglikhd_draw(lay, 7)
        assert np.array_equal(ctx.get_X(1), X1, equal_nan=True)
        # guiding cache: F = F0 + Psi v reproduces the backward filter
        kb = (40, 30)[lay]                                   # first interval of a block (c is only kept where it is read)
        H, F, c = ctx.get_guiding_term(kb, 0, 0)
        ctx.enable_guiding_cache(lay)
        ctx.recompute_guiding_term(lay, _lib.P_ONLY)
        Hc, Fc, cc = ctx.get_layout_guiding_term(lay, kb)
        assert rel_err(Hc[:-1], H[:-1]) < 1e-10 and rel_err(Fc[:-1], F[:-1]) < 1e-9 and np.abs(cc[0] - c[0]).max() < 1e-8 * np.abs(c[0]).max()
        ctx.enable_guiding_cache(lay, False)
        # (5) accept bookkeeping: ll after the step is ll° where accepted, else ll; decisions follow E > -(ll° - ll)
        ll0, ll1 = ctx.get_ll(lay, 0).copy(), ctx.get_ll(lay, 1).copy()
        E = np.random.default_rng(lay).exponential(size=ll0.shape)
        ctx.accept_reject_path(lay, 7, E)
        acc = ctx.get_last_accept(lay)
        assert np.array_equal(acc, E > -(ll1 - ll0))
        assert np.array_equal(ctx.get_ll(lay, 0), np.where(acc, ll1, ll0))
        assert 0.2 < acc.mean() < 0.8


def test_slice_of_the_full_ensemble_replayed_on_the_oracle(full, orc, olib):
    """chains 2501..2503 of 4096: same data, same global Philox counters -> the oracle reproduces the device's sweep"""
    prob, ctx = full
    sub = copy.copy(prob)
    sub.M = sub.P = SLICE.stop - SLICE.start
    sub.v, sub.xbar, sub.x0 = prob.v[:, :, SLICE].copy(), prob.xbar[:, :, SLICE].copy(), prob.x0[:, SLICE].copy()
    ora = OracleEnsemble(orc, olib, sub, seed=123, chain_offset=SLICE.start)
    X, W = ctx.get_X(0)[:, :, SLICE], ctx.get_W(0)[:, :, SLICE]
    for s in (0, 1):
        ora.set_X(s, np.ascontiguousarray(X)); ora.set_W(s, np.ascontiguousarray(W))
    for lay, it in ((0, 11), (1, 12)):
        ctx.blocking_sweep(lay, it)
        ora.set_artificial_obs(lay); ora.recompute_guiding_term(lay); ora.find_W_for_X(lay); ora.loglikhd(lay); ora.draw(lay, it)
        assert rel_err(ctx.get_W(0)[:, :, SLICE], ora.W(0)) < 1e-9
        assert rel_err(ctx.get_X(1)[:, :, SLICE], ora.X(1)) < 1e-9 and rel_err(ctx.get_W(1)[:, :, SLICE], ora.W(1)) < 1e-9
        assert rel_err(ctx.get_ll(lay, 0)[:, SLICE], ora.ll(lay, 0)) < 1e-9 and rel_err(ctx.get_ll(lay, 1)[:, SLICE], ora.ll(lay, 1)) < 1e-9
        ctx.accept_reject_path(lay, it)
        acc_o, _ = ora.accept(lay, it, layout_id=lay)
        assert np.array_equal(ctx.get_last_accept(lay)[:, SLICE], acc_o)
        assert rel_err(ctx.get_X(0)[:, :, SLICE], ora.X(0)) < 1e-9


# ---- the other BASELINE configs at their full sizes --------------------------------------------------------------------------
FULL = {  # config -> (expected (M, K, steps per chain), chains replayed on the oracle)
    "c2": ((1024, 50, 5000), [700, 701, 1023]),
    "c4": ((16384, 100, 10000), [9001, 16383]),
    "c5": ((8192, 500, 50000), [5000, 8191]),
}


@pytest.mark.parametrize("cfg", sorted(FULL))
def test_other_baseline_configs_at_full_size(cfg, orc, olib):
    """C2 (Lotka-Volterra, 1024 chains), C4 (Prokaryote, 16384 chains, state-dependent diffusion), C5 (Jansen-Rit, 8192 chains,
    50,000 steps each, K1 every sweep): the whole ensemble through size-independent properties, and a few of its chains — with
    their global Philox counters — replayed call by call on the oracle (paths of single recordings travel through
    dmt_get_X_chains / dmt_get_W_chains; the full arrays are tens of GB)."""
    (M, K, S), chains = FULL[cfg]
    prob = configs.named_config(cfg, seed=123, sim_sub=1)
    assert (prob.M, prob.P, prob.K, prob.steps_per_chain) == (M, M, K, S)
    ctx = make_ctx(prob, seed=123, n_layouts=1)
    ctx.recompute_guiding_term(0, _lib.P_ONLY)
    assert ctx.init_paths(0, 1 << 20, 100) == 0
    ctx.loglikhd(0, 0, 0)
    ll = ctx.get_ll(0, 0)
    assert np.isfinite(ll).all()
    tot, per_block = ctx.fetch_ll(0, 0)
    assert abs(tot - ll.sum()) < 1e-10 * abs(tot)                       # checksum of checksums (fixed-order device tree)
    # the slice on the oracle
    sub = copy.copy(prob)
    sub.M = sub.P = len(chains)
    sub.v, sub.xbar, sub.x0 = prob.v[:, :, chains].copy(), prob.xbar[:, :, chains].copy(), prob.x0[:, chains].copy()
    ora = OracleEnsemble(orc, olib, sub, seed=123, chain_ids=chains)                # global chain ids, not contiguous
    X, W = ctx.get_X_chains(chains, 0), ctx.get_W_chains(chains, 0)
    assert np.isfinite(X).all() and np.isfinite(W).all()
    for s in (0, 1):
        ora.set_X(s, X); ora.set_W(s, W)
    ora.recompute_guiding_term(0); ora.loglikhd(0)
    tag = "full/%s/" % cfg
    assert rel_err(ll[:, chains], ora.ll(0, 0), tag=tag + "ll") < 1e-10
    for k in (0, K // 2, K - 1):                                        # K1 at full size for the replayed parameter sets
        H, F, c = ctx.get_guiding_term(k, 0, 0); Ho, Fo, co = ora.guiding(k, 0, 0)
        assert rel_err(H[:-1][..., chains], Ho[:-1], tag=tag + "H") < 1e-10 and rel_err(F[:-1][..., chains], Fo[:-1], tag=tag + "F") < 1e-10
        assert rel_err(c[0][chains], co[0], tag=tag + "c") < 1e-10
        del H, F, c
    # K5 o K2 = identity on the replayed chains (invsolve of the initial path returns the noise that made it)
    ctx.find_W_for_X(0)
    assert rel_err(ctx.get_W_chains(chains, 0), W, tag=tag + "K5oK2") < 1e-7
    ctx.set_rho(0, [0.9])
    for it in (3, 4):
        if cfg == "c5" and it == 4:     # BASELINE C5: "each sweep = set_params -> K1 (P = M) -> K2": new parameters, device re-linearisation
            th = prob.theta * (1 + 1e-3)
            ctx.set_params(th, side=0, stores=1); ctx.set_aux_linearised(None, side=0, store=_lib.STORE_PP)
            ctx.recompute_guiding_term(0, _lib.P_ONLY); ctx.loglikhd(0, 0, 0)
            for c_, P in enumerate(ora.pairs):
                P.set_theta(th, side=0)
                for k in range(K):
                    B, beta, at = orc.linearise(olib, prob.model, th, sub.xbar[k, :, c_])
                    P.set_aux(k, B, beta, at, side=0, store=0)
            ora.recompute_guiding_term(0); ora.loglikhd(0)
            assert rel_err(ctx.get_ll(0, 0)[:, chains], ora.ll(0, 0), tag=tag + "ll_newtheta") < 1e-10
        ctx.draw_proposal_path(0, it); ok_o = ora.draw(0, it)
        assert np.array_equal(ctx.get_success(0)[:, chains], ok_o)
        good = ok_o[0]
        if good.any():
            sel = [c for c, g in zip(chains, good) if g]
            assert rel_err(ctx.get_X_chains(sel, 1), ora.X(1)[:, :, good], tag=tag + "X_prop") < 1e-10
            assert rel_err(ctx.get_W_chains(sel, 1), ora.W(1)[:, :, good], tag=tag + "W_prop") < 1e-10
        assert rel_err(ctx.get_ll(0, 1)[:, chains], ora.ll(0, 1), tag=tag + "ll_prop") < 1e-10
        ll0, ll1 = ctx.get_ll(0, 0).copy(), ctx.get_ll(0, 1).copy()
        ctx.accept_reject_path(0, it); acc_o, _ = ora.accept(0, it)
        acc = ctx.get_last_accept(0)
        assert np.array_equal(acc[:, chains], acc_o)
        assert np.array_equal(ctx.get_ll(0, 0), np.where(acc, ll1, ll0))  # bookkeeping over the WHOLE ensemble
        note_err(tag + "accept_frac_it%d" % it, acc.mean())
        assert acc.mean() < 0.98 and (cfg == "c5" or acc.mean() > 0.02)   # (C5's 50,000-step paths accept rarely at rho = 0.9)
        assert rel_err(ctx.get_X_chains(chains, 0), ora.X(0), tag=tag + "X") < 1e-10
    ctx.close()


# ---- the configuration bench.py measures: lazy noise + guiding cache, at the per-GPU sizes of the 1 / 8 / 16-GPU splits ---------------
@pytest.mark.parametrize("chains,lo,kernel", [(4096, 0, "sweep_pipe_kernel<lanes=1, lazy>"),            # 1 GPU
                                              (1536, 1536, "fwd_kernel<op=6, lanes=2, lazy>"),          # (a 3-GPU-like shard: two lanes per (chain, block))
                                              (512, 1536, "sweep_sp_kernel<step lanes=4, lazy>"),       # rank 3 of 8: step-parallel lanes
                                              (256, 3840, "sweep_ws_kernel<wide, lazy>")])              # rank 15 of 16: one round of CTAs
def test_bench_configuration_slice_replay(chains, lo, kernel, orc, olib):
    """C3 exactly as bench.py runs it (lazy noise, guiding cache, automatic kernel choice) on the slice of the 4096-chain ensemble a rank of
    an N-GPU job holds: three sweeps per layout with accept steps; three chains of the slice — global Philox counters — are replayed call by
    call on the oracle.  Also pins WHICH kernel the automatic choice takes at each size, so all four fused-pass kernels meet the oracle at
    BASELINE scale."""
    prob = configs.named_config("c3", M=chains, seed=123, chain_offset=lo)          # the rank's slice of the ensemble's synthetic data
    if chains == 512:                                                               # (checked once: a shard IS a slice of the unsharded data)
        full = configs.named_config("c3", seed=123)
        assert np.array_equal(prob.v, full.v[:, :, lo:lo + chains]) and np.array_equal(prob.x0, full.x0[:, lo:lo + chains])
    ctx = make_ctx(prob, seed=123, n_layouts=3, chain_offset=lo)
    ctx.set_blocks(2, [(0, prob.K - 1)], 0.0)
    ctx.recompute_guiding_term(2, _lib.P_ONLY)
    assert ctx.init_paths(2, 0, 50) == 0
    ctx.set_lazy_noise(True)
    for lay in (0, 1):
        ctx.enable_guiding_cache(lay)
    sel = [0, chains // 2 + 1, chains - 1]
    sub = copy.copy(prob)
    sub.M = sub.P = len(sel)
    sub.v, sub.xbar, sub.x0 = prob.v[:, :, sel].copy(), prob.xbar[:, :, sel].copy(), prob.x0[:, sel].copy()
    oras = [OracleEnsemble(orc, olib, _one(sub, i), seed=123, chain_offset=lo + c) for i, c in enumerate(sel)]
    X, W = ctx.get_X(0)[:, :, sel], ctx.get_W(0)[:, :, sel]
    for i, ora in enumerate(oras):
        for s in (0, 1):
            ora.set_X(s, np.ascontiguousarray(X[:, :, i:i + 1])); ora.set_W(s, np.ascontiguousarray(W[:, :, i:i + 1]))
    n_acc = 0
    for it in range(3):
        for lay in (0, 1):
            ctx.blocking_sweep(lay, it)
            assert ctx.last_forward_kernel() == kernel
            ctx.accept_reject_path(lay, it)
            acc = ctx.get_last_accept(lay)[:, sel]
            Xd, lld, llod = ctx.get_X(0)[:, :, sel], ctx.get_ll(lay, 0)[:, sel], ctx.get_ll(lay, 1)[:, sel]
            for i, ora in enumerate(oras):
                ora.set_artificial_obs(lay); ora.recompute_guiding_term(lay); ora.find_W_for_X(lay); ora.loglikhd(lay); ora.draw(lay, it)
                ll_o, llo_o = ora.ll(lay, 0)[:, 0].copy(), ora.ll(lay, 1)[:, 0].copy()
                acc_o, _ = ora.accept(lay, it, layout_id=lay)
                assert np.array_equal(acc[:, i], acc_o[:, 0]), (it, lay, i)
                # after the accept step ll[side] have been swapped where accepted: compare the pre-accept values through the swap
                pre_ll = np.where(acc[:, i], llod[:, i], lld[:, i]); pre_llo = np.where(acc[:, i], lld[:, i], llod[:, i])
                assert rel_err(pre_ll, ll_o) < 1e-9 and rel_err(pre_llo[np.isfinite(llo_o)], llo_o[np.isfinite(llo_o)]) < 1e-9
                assert rel_err(Xd[:, :, i:i + 1], ora.X(0)) < 1e-9
            n_acc += int(acc.sum())
    assert n_acc > 0
    ctx.close()


def _one(sub, i):
    s = copy.copy(sub)
    s.M = s.P = 1
    s.v, s.xbar, s.x0 = sub.v[:, :, i:i + 1].copy(), sub.xbar[:, :, i:i + 1].copy(), sub.x0[:, i:i + 1].copy()
    return s
