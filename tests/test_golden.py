"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py with the CPU oracle at fixed seeds).

CPU:  the oracle still reproduces its committed vectors (guards the checker against drift).
GPU:  the CUDA path, through the C ABI, reproduces them: paths / ll within 1e-10 relative (1e-9 after K5), decisions equal.
"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402
from harness import make_ctx, rel_err  # noqa: E402

CASES = sorted(make_golden.CASES)


def load(case):
    return dict(np.load(os.path.join(HERE, "golden", case + ".npz")))


@pytest.mark.parametrize("case", CASES)
def test_oracle_reproduces_golden(case):
    want, got = load(case), make_golden.run_oracle(case)
    assert sorted(want) == sorted(got)
    for k in want:
        if want[k].dtype == bool:
            assert np.array_equal(want[k], got[k]), k
        else:
            assert rel_err(got[k], want[k]) < 1e-12, k       # same source, same flags: identical up to libm/compiler version


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_cuda_reproduces_golden(case):
    from dmt_b200 import _lib
    want = load(case)
    prob, seed = make_golden.problem_of(case)
    nl = len(prob.layouts)
    ctx = make_ctx(prob, seed=seed, n_layouts=nl + 1)
    ctx.set_blocks(nl, [(0, prob.K - 1)], 0.0)
    ctx.recompute_guiding_term(nl, _lib.P_ONLY)
    ctx.set_W(want["W0"], 0); ctx.set_W(want["W0"], 1)
    ctx.recompute_path(nl, 0, 0)
    assert ctx.get_success(nl).all()
    X0 = ctx.get_X(0)
    assert rel_err(X0, want["X0"]) < 1e-10
    ctx.set_X(X0, 1)
    blocking = nl > 1
    tol = 1e-9 if blocking else 1e-10
    for l in range(nl):
        if blocking:
            ctx.set_artificial_obs(l)
        ctx.recompute_guiding_term(l, _lib.P_ONLY)
        if blocking:
            ctx.find_W_for_X(l)
        ctx.loglikhd(l, 0, 0)
        assert rel_err(ctx.get_ll(l, 0), want["ll_%d" % l]) < tol
        ctx.draw_proposal_path(l, l)
        assert np.array_equal(ctx.get_success(l), want["ok_%d" % l])
        assert rel_err(ctx.get_ll(l, 1), want["llo_%d" % l]) < tol
        assert rel_err(ctx.get_X(1), want["Xo_%d" % l]) < tol and rel_err(ctx.get_W(1), want["Wo_%d" % l]) < tol
        ctx.accept_reject_path(l, l)
        assert np.array_equal(ctx.get_last_accept(l), want["acc_%d" % l])
        assert rel_err(ctx.get_X(0), want["Xacc_%d" % l]) < tol
        H, F, c = ctx.get_guiding_term(prob.K - 1, 0, 0)
        n = H.shape[0]
        assert rel_err(H[:n - 1], want["H_last_%d" % l][:n - 1]) < 1e-10 and rel_err(F[:n - 1], want["F_last_%d" % l][:n - 1]) < 1e-10
        assert rel_err(c[0], want["c_last_%d" % l]) < 1e-10
    ctx.close()
