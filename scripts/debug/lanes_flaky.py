import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "tests"))
import numpy as np
import dmt_b200
from dmt_b200 import _lib, configs
from harness import make_ctx

K = 6
layouts = [([(0, 1), (2, 3), (4, 5)], 0.7), ([(0, K - 1)], 0.0)]
prob = configs.make_problem("lorenz", 41, K=K, obs_dt=0.1, dt=0.01, seed=5, layouts=layouts, rho=0.7)

def eq(a, b): return np.array_equal(a, b, equal_nan=True)

nfail = 0
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 20):
    ctxs = []
    for lanes in (1, 8):
        ctx = make_ctx(prob, seed=77, ll_hist_len=4); ctx.set_fwd_lanes(lanes); ctxs.append(ctx)
    a, b = ctxs
    stage = None
    for ctx in ctxs:
        ctx.recompute_guiding_term(1, _lib.P_ONLY)
        assert ctx.init_paths(1, iter0=900, max_tries=50) == 0
    if not (eq(a.get_X(0), b.get_X(0)) and eq(a.get_W(0), b.get_W(0))): stage = "init"
    for it in range(3):
        if stage: break
        for ctx in ctxs: ctx.set_artificial_obs(0); ctx.recompute_guiding_term(0, _lib.P_ONLY)
        for k in range(K):
            for x, y in zip(a.get_guiding_term(k, 0, 0), b.get_guiding_term(k, 0, 0)):
                if not eq(x, y): stage = "K1 it%d k%d" % (it, k)
        if stage: break
        for ctx in ctxs: ctx.find_W_loglikhd_draw(0, it)
        for nm, f in (("W_acc", lambda c: c.get_W(0)), ("X_prop", lambda c: c.get_X(1)), ("W_prop", lambda c: c.get_W(1)),
                      ("ll", lambda c: c.get_ll(0, 0)), ("ll_prop", lambda c: c.get_ll(0, 1)), ("ok", lambda c: c.get_success(0))):
            x, y = f(a), f(b)
            if not eq(x, y):
                bad = np.argwhere(~((x == y) | (np.isnan(x) & np.isnan(y)))) if x.dtype != bool else np.argwhere(x != y)
                stage = "sweep it%d %s: %d places, first %s %r vs %r, chains %s" % (it, nm, len(bad), bad[0], x[tuple(bad[0])], y[tuple(bad[0])], sorted(set(bad[:, -1].tolist()))[:12])
                break
        if stage: break
        for ctx in ctxs: ctx.accept_reject_path(0, it)
        if not eq(a.get_last_accept(0), b.get_last_accept(0)): stage = "accept it%d" % it
    print("rep", rep, "FIRST DIFFERENCE:" if stage else "identical", stage or "")
    nfail += bool(stage)
    for ctx in ctxs: ctx.close()
print("failures", nfail)
