#!/usr/bin/env python
"""Generates tests/golden/*.npz with the CPU oracle (oracle/dmt_oracle.c) at fixed Philox seeds.

The reference ships no golden vectors (test/runtests.jl:4-6 is empty) and cannot run here (no Julia), so these fixtures are
ORACLE-generated: they pin the oracle against drift (tests/test_golden.py, CPU) and the CUDA path against the oracle at a fixed
point in time (tests/test_golden.py, -m gpu).  Re-run only when the oracle's definition changes deliberately:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = {  # name -> (model name, M, K, nsteps/interval, layouts, seed)
    "fhn_single": ("fhn", 6, 4, 10, [([(0, 3)], 0.96)], 100),
    "lorenz_blocking": ("lorenz", 5, 6, 11, [([(0, 1), (2, 3), (4, 5)], 0.9), ([(0, 2), (3, 5)], 0.8)], 7),
    "prok_single": ("prok", 4, 3, 10, [([(0, 2)], 0.9)], 11),
    "jr_single": ("jr", 3, 3, 12, [([(0, 2)], 0.9)], 13),
    "lv_single": ("lv", 5, 3, 10, [([(0, 2)], 0.5)], 17),
}


def problem_of(case):
    from dmt_b200 import configs
    name, M, K, ns, layouts, seed = CASES[case]
    obs_dt = 0.01 if name == "jr" else 0.1
    return configs.make_problem(name, M, K=K, obs_dt=obs_dt, dt=obs_dt / ns, seed=seed, layouts=layouts), seed


def run_oracle(case):
    """the fixed scenario: whole-path K1, deterministic initial noise -> path, then per layout: [blocking: set_obs, K1, K5, K4],
    draw (Philox iter = layout), accept (Philox)."""
    from harness import OracleEnsemble
    from oracle import orc
    olib = orc.load()
    prob, seed = problem_of(case)
    ora = OracleEnsemble(orc, olib, prob, seed=seed)
    out = {}
    rng = np.random.default_rng(seed)
    dts = np.concatenate([np.diff(prob.tt[a:b]) for a, b in zip(np.cumsum(prob.n_pts) - prob.n_pts, np.cumsum(prob.n_pts))])
    W0 = 0.5 * np.sqrt(dts)[:, None, None] * rng.normal(size=(prob.steps_per_chain, prob.dw, prob.M))
    out["W0"] = W0
    for P in ora.pairs:   # whole path from W0 under the full (unblocked) guiding term
        bb = P.biblock(0, prob.K - 1, True, 0.0)
        P.recompute_guiding_term(bb, 0)
    ora.set_W(0, W0); ora.set_W(1, W0)
    for c, P in enumerate(ora.pairs):
        bb = P.biblock(0, prob.K - 1, True, 0.0)
        assert P.recompute_path(bb, 0, 0)
        for k in range(prob.K):
            P.set_X(1, k, P.get_X(0, k))
    out["X0"] = ora.X(0)
    blocking = len(prob.layouts) > 1
    for l in range(len(prob.layouts)):
        if blocking:
            ora.set_artificial_obs(l)
        ora.recompute_guiding_term(l)
        if blocking:
            ora.find_W_for_X(l)
        ora.loglikhd(l)
        out["ll_%d" % l] = ora.ll(l, 0)
        out["ok_%d" % l] = ora.draw(l, l)
        out["llo_%d" % l] = ora.ll(l, 1)
        out["Xo_%d" % l] = ora.X(1)
        out["Wo_%d" % l] = ora.W(1)
        acc, _ = ora.accept(l, l)
        out["acc_%d" % l] = acc
        out["Xacc_%d" % l] = ora.X(0)
        H, F, c = ora.guiding(prob.K - 1, 0, 0)
        out["H_last_%d" % l], out["F_last_%d" % l], out["c_last_%d" % l] = H, F, c[0]
    return out


if __name__ == "__main__":
    for case in CASES:
        res = run_oracle(case)
        path = os.path.join(HERE, case + ".npz")
        np.savez_compressed(path, **res)
        print(case, os.path.getsize(path), "bytes")
