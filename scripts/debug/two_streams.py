"""Experiment: the C3 ensemble as G independent chain groups, one context (= one CUDA stream) each, sweeps enqueued round-robin.
Chains are independent, so the HBM-bound cached K1 of one group can overlap the latency-bound fused pass of another."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np
import dmt_b200
from dmt_b200 import _lib, configs

M_total = 4096
nsweeps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
for groups in (1, 2, 4):
    Mg = M_total // groups
    ctxs = []
    for g in range(groups):
        prob = configs.named_config("c3", M=Mg, seed=123, chain_offset=g * Mg)
        ctx = dmt_b200.Ctx(prob.model, prob.n_pts, prob.tt, prob.M, prob.P, obs_dim=prob.m, two_sided_laws=False, n_layouts=3,
                           chain_offset=g * Mg, seed=123, artificial_noise=prob.eps)
        configs.upload(prob, ctx, sides=(0,))
        ctx.set_blocks(2, [(0, prob.K - 1)], 0.0)
        ctx.recompute_guiding_term(2, _lib.P_ONLY)
        assert ctx.init_paths(2, 0, 50) == 0
        for l in (0, 1):
            ctx.enable_guiding_cache(l)
        ctxs.append(ctx)
    def sweeps(n, it0):
        for i in range(n):
            for ctx in ctxs:
                l = (it0 + i) & 1
                ctx.blocking_sweep(l, it0 + i)
                ctx.accept_reject_path(l, it0 + i)
        for ctx in ctxs:
            ctx.sync()
    sweeps(6, 0)
    t0 = time.perf_counter(); sweeps(nsweeps, 6); dt = time.perf_counter() - t0
    steps = M_total * 20000 * nsweeps
    print("groups %d: %.3f ms per sweep of all %d chains, %.3e steps/s" % (groups, 1e3 * dt / nsweeps, M_total, steps / dt), flush=True)
    for ctx in ctxs:
        ctx.close()
