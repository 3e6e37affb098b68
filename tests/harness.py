"""Twin runner: the same Problem pushed into the CUDA library (through the C ABI) and into the CPU oracle, with every
reference operation mirrored on both so tests can compare after each call.

Oracle side = one `orc.Pair` per chain (the reference's per-recording object graph, pointer swaps included).
"""
import numpy as np

import dmt_b200
from dmt_b200 import _lib


import json
import os

_WORST = {}


def rel_err(a, b, floor=None, tag=None):
    """ELEMENT-WISE relative error  max_i |a_i - b_i| / max(|b_i|, floor)  of `a` against the reference values `b`.
    The absolute floor (default: the root-mean-square of the finite reference values) keeps entries that pass through zero —
    Wiener increments, state components at a crossing — from dividing by ~0; everything at or above typical size is compared
    relative to ITSELF, not to the largest entry of the array.  Equal infinities and matching NaNs count as equal; a NaN or an
    infinity on one side only is an infinite error (never ignored)."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        a, b = np.broadcast_arrays(a, b)
    if a.size == 0:
        return 0.0
    fin = np.isfinite(a) & np.isfinite(b)
    same_special = (np.isnan(a) & np.isnan(b)) | (np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b)))
    if not np.all(fin | same_special):
        return float("inf")
    if not fin.any():
        return 0.0
    bf = np.abs(b[fin])
    fl = float(floor) if floor is not None else float(np.sqrt(np.mean(bf * bf)))
    fl = max(fl, 1e-300)
    e = float(np.max(np.abs(a[fin] - b[fin]) / np.maximum(bf, fl)))
    if tag is not None:
        note_err(tag, e)
    return e


def note_err(tag, e):
    """remember the worst error seen per tag; DMT_PARITY_LOG=<file> writes the table when the process exits (the measured
    GPU-vs-oracle discrepancies quoted in BASELINE.md come from this log)"""
    _WORST[tag] = max(_WORST.get(tag, 0.0), float(e))


def _dump_worst():
    path = os.environ.get("DMT_PARITY_LOG")
    if path and _WORST:
        try:
            old = json.load(open(path)) if os.path.exists(path) else {}
        except Exception:
            old = {}
        for k, v in _WORST.items():
            old[k] = max(old.get(k, 0.0), v)
        with open(path, "w") as f:
            json.dump(old, f, indent=1, sort_keys=True)


import atexit  # noqa: E402
atexit.register(_dump_worst)


class OracleEnsemble:
    def __init__(self, orc, olib, prob, seed=0, chain_offset=0, chain_ids=None):
        """chain_ids: the GLOBAL recording index of each of prob's chains (random-stream counters); default chain_offset + c"""
        self.orc, self.olib, self.prob = orc, olib, prob
        self.seed, self.chain_offset = seed, chain_offset
        self.chain_ids = [chain_offset + c for c in range(prob.M)] if chain_ids is None else [int(c) for c in chain_ids]
        self.pairs = []
        ps_of = prob.pset_of_chain if prob.pset_of_chain is not None else (np.arange(prob.M) if prob.P == prob.M else np.zeros(prob.M, int))
        self.ps_of = ps_of
        for c in range(prob.M):
            ps = int(ps_of[c])
            P = orc.Pair(olib, prob.model, prob.n_pts, prob.tt, prob.m, prob.eps)
            P.set_theta(prob.theta)
            for k in range(prob.K):
                B, beta, at = orc.linearise(olib, prob.model, prob.theta, prob.xbar[k, :, ps])
                P.set_aux(k, B, beta, at)
                P.set_obs(k, prob.L, prob.Sigma, prob.v[k, :, ps])
            P.set_start(prob.x0[:, c])
            self.pairs.append(P)
        self.layouts = []
        for ranges, rho in prob.layouts:
            nb = len(ranges)
            rhos = np.broadcast_to(np.asarray(rho, float), (nb,))
            self.layouts.append([[P.biblock(r[0], r[1], b == nb - 1, float(rhos[b])) for b, r in enumerate(ranges)] for P in self.pairs])

    # ---- mirrored ops (layout l)
    def each(self, l):
        for c, P in enumerate(self.pairs):
            for b, bb in enumerate(self.layouts[l][c]):
                yield c, b, P, bb

    def set_artificial_obs(self, l):
        for c, b, P, bb in self.each(l):
            P.set_artificial_obs(bb)

    def recompute_guiding_term(self, l, sides=(0,)):
        for c, b, P, bb in self.each(l):
            for s in sides:
                P.recompute_guiding_term(bb, s)

    def find_W_for_X(self, l):
        for c, b, P, bb in self.each(l):
            P.find_W_for_X(bb)

    def loglikhd(self, l, side=0, skip=0):
        for c, b, P, bb in self.each(l):
            P.loglikhd(bb, side, skip)

    def draw(self, l, it, Z=None, layout_id=None):
        """Z: [S][dw][M] natural layout or None (Philox).  layout_id: the id the device library registered this layout under
        (part of the pCN stream's counter, like the accept stream's)"""
        lid = l if layout_id is None else layout_id
        ok = np.zeros((len(self.layouts[l][0]), self.prob.M), bool)
        step0 = np.concatenate([[0], np.cumsum(self.prob.n_pts - 1)])
        for c, b, P, bb in self.each(l):
            zb = None
            if Z is not None:
                zb = np.ascontiguousarray(Z[step0[bb.i0]:step0[bb.i1 + 1], :, c])
            ok[b, c] = P.draw_proposal_path(bb, zb, seed=self.seed, chain=self.chain_ids[c], it=it, layout=lid)
        return ok

    def recompute_path(self, l, law_side, w_side, skip=0):
        ok = np.zeros((len(self.layouts[l][0]), self.prob.M), bool)
        for c, b, P, bb in self.each(l):
            ok[b, c] = P.recompute_path(bb, law_side, w_side, skip)
        return ok

    def accept(self, l, it, E=None, layout_id=None):
        """layout_id: the id the device library registered this layout under (part of the accept stream's counter)"""
        nb = len(self.layouts[l][0])
        acc = np.zeros((nb, self.prob.M), bool)
        hist = np.zeros((2, nb, self.prob.M))
        lid = l if layout_id is None else layout_id
        for c, b, P, bb in self.each(l):
            e = E[b, c] if E is not None else self.olib.orc_accept_exponential(self.seed, self.chain_ids[c], b, it, lid)
            acc[b, c], h = P.accept_reject(bb, float(e))
            hist[:, b, c] = h
        return acc, hist

    def swap(self, l, what, mask=None):
        for c, b, P, bb in self.each(l):
            if mask is not None and not mask[c]:
                continue
            if what & 1: P.swap_XX(bb)
            if what & 2: P.swap_WW(bb)
            if what & 4: P.swap_PP(bb)
            if what & 8: P.swap_ll(bb)

    # ---- state in the C ABI's natural layouts
    def ll(self, l, side):
        nb = len(self.layouts[l][0])
        out = np.zeros((nb, self.prob.M))
        for c, b, P, bb in self.each(l):
            out[b, c] = bb.ll[side]
        return out

    def set_ll(self, l, side, ll):
        for c, b, P, bb in self.each(l):
            bb.ll[side] = ll[b, c]

    def X(self, side):
        return np.stack([np.concatenate([P.get_X(side, k) for k in range(self.prob.K)]) for P in self.pairs], axis=2)

    def W(self, side):
        return np.stack([np.concatenate([P.get_W(side, k) for k in range(self.prob.K)]) for P in self.pairs], axis=2)

    def set_W(self, side, W):
        step0 = np.concatenate([[0], np.cumsum(self.prob.n_pts - 1)])
        for c, P in enumerate(self.pairs):
            for k in range(self.prob.K):
                P.set_W(side, k, W[step0[k]:step0[k + 1], :, c])

    def set_X(self, side, X):
        pt0 = np.concatenate([[0], np.cumsum(self.prob.n_pts)])
        for c, P in enumerate(self.pairs):
            for k in range(self.prob.K):
                P.set_X(side, k, X[pt0[k]:pt0[k + 1], :, c])

    def guiding(self, k, side=0, store=0):
        """H [n,d,d,P], F [n,d,P], c [n,P] using, for each pset, the first chain that maps to it"""
        prob = self.prob
        n, d = int(prob.n_pts[k]), prob.d
        H = np.zeros((n, d, d, prob.P)); F = np.zeros((n, d, prob.P)); cc = np.zeros((n, prob.P))
        seen = set()
        for c, P in enumerate(self.pairs):
            ps = int(self.ps_of[c])
            if ps in seen:
                continue
            seen.add(ps)
            h, f, c_ = P.get_HFc(side, store, k)
            H[..., ps], F[..., ps], cc[:, ps] = h, f, c_
        return H, F, cc


def make_ctx(prob, seed=0, chain_offset=0, two_sided=False, ll_hist_len=0, n_layouts=None):
    ctx = dmt_b200.Ctx(prob.model, prob.n_pts, prob.tt, prob.M, prob.P, obs_dim=prob.m, two_sided_laws=two_sided,
                       ll_hist_len=ll_hist_len, n_layouts=n_layouts or max(1, len(prob.layouts)), chain_offset=chain_offset,
                       seed=seed, artificial_noise=prob.eps, pset_of_chain=prob.pset_of_chain)
    dmt_b200.configs.upload(prob, ctx, sides=(0, 1) if two_sided else (0,))
    return ctx


def compare_guiding(ctx, ora, k, side=0, store=0, tol=1e-10, tag=None, c_cancel=0.0):
    """H, F at every grid point (element-wise relative, floor = the rms of that grid point's own entries, per parameter set:
    on an exact-observation interval H spans ten orders of magnitude between its two ends) and c at the interval start.
    c_cancel: magnitude of the terms that cancel inside c when c itself is integrated from an exact observation (the Tsit5 mode
    starts at c_T = v'v / 2 eps ~ 1e13 and ends at O(100)): c cannot be reproduced below ~1e3 ulp of that, whoever computes it."""
    H, F, c = ctx.get_guiding_term(k, side, store)
    Ho, Fo, co = ora.guiding(k, side, store)
    n = H.shape[0]
    flH = np.sqrt(np.mean(Ho * Ho, axis=(1, 2), keepdims=True)); flF = np.sqrt(np.mean(Fo * Fo, axis=1, keepdims=True))
    eH = float(np.max(np.abs(H[:n - 1] - Ho[:n - 1]) / np.maximum(np.abs(Ho[:n - 1]), np.maximum(flH[:n - 1], 1e-300))))
    eF = float(np.max(np.abs(F[:n - 1] - Fo[:n - 1]) / np.maximum(np.abs(Fo[:n - 1]), np.maximum(flF[:n - 1], 1e-300))))
    ec = float(np.max(np.abs(c[0] - co[0]) / np.maximum(np.maximum(1.0, np.abs(co[0])), 1e3 * 2.2e-16 * c_cancel / tol)))
    if tag:
        note_err(tag + "/H", eH); note_err(tag + "/F", eF); note_err(tag + "/c", ec)
    assert np.isfinite(H[:n - 1]).all() and np.isfinite(F[:n - 1]).all()
    assert eH < tol and eF < tol and ec < tol, (k, side, store, eH, eF, ec)
    return max(eH, eF, ec)
