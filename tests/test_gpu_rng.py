"""Device random streams (Philox4x32-10 + custom FP64 log / sqrt / sincospi, csrc/fastmath.cuh) against the oracle's
independent Philox + libm transforms, plus distributional checks at scale."""
import numpy as np
import pytest

import dmt_b200
from dmt_b200 import _lib, configs
from harness import make_ctx

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["fhn", "lv", "lorenz", "prok"])
def test_normals_match_oracle(orc, olib, name):
    prob = configs.make_problem(name, 4, K=2, seed=1)
    ctx = make_ctx(prob, seed=0xD1FF00012345)
    nc, nt = 64, 50
    for lay in (0, 5):
        z = ctx.debug_normals(7, 1000, 3, nc, nt, layout=lay)
        zo = np.stack([[orc.tile_normals(olib, 0xD1FF00012345, 7 + c, 1000 + q, 3, prob.dw, layout=lay) for q in range(nt)] for c in range(nc)])
        assert np.abs(z - zo).max() < 1e-14 * max(1.0, np.abs(zo).max())
    ctx.close()


def test_exponentials_match_oracle(orc, olib):
    prob = configs.make_problem("lorenz", 4, K=2, seed=1)
    ctx = make_ctx(prob, seed=99)
    e = ctx.debug_exponentials(5, 17, 2, 300, 11)
    eo = np.array([[olib.orc_accept_exponential(99, 5 + c, b, 17, 2) for b in range(11)] for c in range(300)])
    assert np.abs(e - eo).max() < 1e-14 * eo.max() and (e > 0).all()
    ctx.close()


def test_normal_moments_and_tails_at_scale():
    prob = configs.make_problem("lorenz", 4, K=2, seed=1)
    ctx = make_ctx(prob, seed=2026)
    z = ctx.debug_normals(0, 0, 0, 4096, 512).ravel()          # 2.5e7 draws
    n = z.size
    assert abs(z.mean()) < 5 / np.sqrt(n) and abs(z.var() - 1) < 5 * np.sqrt(2 / n)
    assert abs((z ** 3).mean()) < 5 * np.sqrt(15 / n) and abs((z ** 4).mean() - 3) < 5 * np.sqrt(96 / n)
    from scipy.stats import norm
    for thr in (1.0, 2.0, 3.0, 4.0):
        p = 2 * norm.sf(thr)
        assert abs((np.abs(z) > thr).mean() - p) < 6 * np.sqrt(p / n)
    assert np.isfinite(z).all() and np.abs(z).max() < 8.6          # sqrt(-2 log 2^-53) = 8.57
    # lag / cross-component correlations
    zz = z.reshape(-1, 12)
    cm = np.corrcoef(zz.T)
    assert np.abs(cm - np.eye(12)).max() < 6 / np.sqrt(zz.shape[0])
    ctx.close()


def test_layouts_swept_with_the_same_iteration_get_independent_innovations():
    """The reference loop uses ONE iteration index for all layouts (docs/src/tutorials/biblock/smoothing_with_blocking.md:32-59):
    the pCN counter carries the layout id, so two sweeps of the same (chain, tile, iteration) never reuse xi."""
    prob = configs.make_problem("lorenz", 4, K=2, seed=1)
    ctx = make_ctx(prob, seed=2026)
    za = ctx.debug_normals(0, 0, 9, 256, 64, layout=0).ravel()
    zb = ctx.debug_normals(0, 0, 9, 256, 64, layout=1).ravel()
    assert not np.any(za == zb)
    assert abs(np.corrcoef(za, zb)[0, 1]) < 6 / np.sqrt(za.size)
    ctx.close()


@pytest.mark.own_lanes
def test_two_layouts_same_iteration_draws_are_independent_pcn_moves():
    """End to end: after layout A accepts at iteration i, layout B's proposal at the SAME i must not be rho W + c xi with the xi that is
    already inside W: with rho = 0 the two proposals' noise must be uncorrelated."""
    K = 4
    layouts = [([(0, K - 1)], 0.0), ([(0, K - 1)], 0.0)]
    prob = configs.make_problem("lorenz", 64, K=K, dt=0.01, seed=3, layouts=layouts)
    ctx = make_ctx(prob, seed=5)
    for l in (0, 1):
        ctx.recompute_guiding_term(l, _lib.P_ONLY)
    assert ctx.init_paths(0, 1000, 20) == 0
    ctx.draw_proposal_path(0, 7); Wa = ctx.get_W(1).copy()
    ctx.draw_proposal_path(1, 7); Wb = ctx.get_W(1).copy()
    assert not np.array_equal(Wa, Wb)
    assert abs(np.corrcoef(Wa.ravel(), Wb.ravel())[0, 1]) < 6 / np.sqrt(Wa.size)
    ctx.close()
