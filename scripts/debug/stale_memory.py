import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "tests"))
import numpy as np
import dmt_b200
from dmt_b200 import _lib, configs
from harness import make_ctx

K = 6
layouts = [([(0, 1), (2, 3), (4, 5)], 0.7), ([(0, K - 1)], 0.0)]
prob = configs.make_problem("lorenz", 41, K=K, obs_dt=0.1, dt=0.01, seed=5, layouts=layouts, rho=0.7)
other = configs.make_problem("lorenz", 41, K=K, obs_dt=0.1, dt=0.01, seed=99, layouts=layouts, rho=0.3)

def run(p, lanes, seed, nit=3, stages=None):
    ctx = make_ctx(p, seed=seed, ll_hist_len=4); ctx.set_fwd_lanes(lanes)
    ctx.recompute_guiding_term(1, _lib.P_ONLY)
    assert ctx.init_paths(1, iter0=900, max_tries=50) == 0
    out = [("init X", ctx.get_X(0)), ("init W", ctx.get_W(0)), ("init Xp", ctx.get_X(1))]
    for it in range(nit):
        ctx.blocking_sweep(0, it)
        out += [("sweep%d Wacc" % it, ctx.get_W(0)), ("sweep%d Xp" % it, ctx.get_X(1)), ("sweep%d ll" % it, ctx.get_ll(0, 0)), ("sweep%d llp" % it, ctx.get_ll(0, 1))]
        ctx.accept_reject_path(0, it)
        out += [("acc%d" % it, ctx.get_last_accept(0)), ("X%d" % it, ctx.get_X(0))]
    ctx.close()
    return out

ref = run(prob, 1, 77)
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 8):
    lanes = (1, 8, 1, 2, 4, 1)[rep % 6]
    if rep % 2 == 1:
        run(other, 1, 5)          # pollute the freed memory with a different problem's data
    got = run(prob, lanes, 77)
    first = next((nm for (nm, a), (_, b) in zip(ref, got) if not np.array_equal(a, b, equal_nan=True)), None)
    print("rep", rep, "lanes", lanes, "polluted" if rep % 2 else "clean   ", "->", "identical" if first is None else "FIRST DIFFERENCE at " + first)
