"""ctypes wrapper of the CPU oracle (oracle/dmt_oracle.c).  TEST INFRASTRUCTURE ONLY.

May be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
PARITY UNPINNED: see oracle/dmt_oracle.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)

FHN, LV, LORENZ, PROK, JR, OU2 = range(6)


class BiBlock(C.Structure):
    _fields_ = [("i0", C.c_int), ("i1", C.c_int), ("last", C.c_int), ("rho", C.c_double), ("ll", C.c_double * 2)]


def build(force=False):
    """Compile the oracle with oracle/Makefile (gcc).  Building the checker is not using it."""
    so = os.path.join(_HERE, "libdmt_oracle.so")
    if force or not os.path.exists(so) or not os.path.exists(os.path.join(_HERE, "libdmt_oracle_omp.so")):
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True, capture_output=True)
    return so


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def load(omp=False, path=None):
    """path: an explicitly built flavour of the same source (bench.py rebuilds the OpenMP one with -march=native for the host it runs on)"""
    if path is None:
        name = "libdmt_oracle_omp.so" if omp else "libdmt_oracle.so"
        path = os.path.join(_HERE, name)
        if not os.path.exists(path):
            build()
    lib = C.CDLL(path)
    vp = C.c_void_p
    bbp = C.POINTER(BiBlock)
    sig = {
        "orc_model_dims": (C.c_int, [C.c_int, _ip, _ip, _ip, _ip]),
        "orc_drift": (None, [C.c_int, _dp, _dp, _dp]),
        "orc_sigma": (None, [C.c_int, _dp, _dp, _dp]),
        "orc_jacobian": (None, [C.c_int, _dp, _dp, _dp]),
        "orc_bound_ok": (C.c_int, [C.c_int, _dp, _dp]),
        "orc_linearise": (None, [C.c_int, _dp, _dp, _dp, _dp, _dp]),
        "orc_philox4x32_10": (None, [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
        "orc_tile_normals": (None, [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, _dp]),
        "orc_accept_exponential": (C.c_double, [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]),
        "orc_pair_create": (vp, [C.c_int, C.c_int, _ip, _dp, C.c_int, C.c_double]),
        "orc_pair_destroy": (None, [vp]),
        "orc_set_theta": (None, [vp, C.c_int, C.c_int, C.c_int, _dp]),
        "orc_set_aux": (None, [vp, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp]),
        "orc_set_obs": (None, [vp, C.c_int, C.c_int, _dp, _dp, _dp]),
        "orc_set_start": (None, [vp, _dp]),
        "orc_set_W": (None, [vp, C.c_int, C.c_int, _dp]),
        "orc_set_X": (None, [vp, C.c_int, C.c_int, _dp]),
        "orc_get_W": (None, [vp, C.c_int, C.c_int, _dp]),
        "orc_get_X": (None, [vp, C.c_int, C.c_int, _dp]),
        "orc_get_HFc": (None, [vp, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp]),
        "orc_set_HFc": (None, [vp, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp]),
        "orc_set_artificial_obs": (None, [vp, bbp]),
        "orc_recompute_guiding_term": (None, [vp, bbp, C.c_int]),
        "orc_recompute_guiding_term_tsit5": (C.c_int, [vp, bbp, C.c_int, C.c_double, C.c_double]),
        "orc_tsit5_tableau": (None, [_dp, _dp, _dp, _dp]),
        "orc_find_W_for_X": (None, [vp, bbp]),
        "orc_loglikhd": (C.c_double, [vp, bbp, C.c_int, C.c_int]),
        "orc_draw_proposal_path": (C.c_int, [vp, bbp, _dp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, _ip]),
        "orc_recompute_path": (C.c_int, [vp, bbp, C.c_int, C.c_int, C.c_int]),
        "orc_accept_reject": (C.c_int, [vp, bbp, C.c_double, _dp]),
        "orc_swap_XX": (None, [vp, bbp]),
        "orc_swap_WW": (None, [vp, bbp]),
        "orc_swap_PP": (None, [vp, bbp]),
        "orc_swap_ll": (None, [bbp]),
        "orc_sweep_many": (C.c_double, [C.POINTER(vp), bbp, C.c_int, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32,
                                        C.c_uint32, _ip, C.c_int, C.c_int, _ip]),
    }
    for name, (res, args) in sig.items():
        f = getattr(lib, name)
        f.restype = res
        f.argtypes = args
    return lib


def model_dims(lib, model):
    d, dw, npar, cd = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    if lib.orc_model_dims(model, C.byref(d), C.byref(dw), C.byref(npar), C.byref(cd)):
        raise ValueError("unknown model %r" % model)
    return d.value, dw.value, npar.value, bool(cd.value)


def linearise(lib, model, theta, xbar):
    d, dw, npar, _ = model_dims(lib, model)
    B = np.zeros((d, d)); beta = np.zeros(d); at = np.zeros((d, d))
    th, thp = _d(theta); xb, xbp = _d(xbar)
    lib.orc_linearise(model, thp, xbp, B.ctypes.data_as(_dp), beta.ctypes.data_as(_dp), at.ctypes.data_as(_dp))
    return B, beta, at


def philox(lib, ctr, key):
    c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
    lib.orc_philox4x32_10(c, k, o)
    return list(o)


def tile_normals(lib, seed, chain, gtile, it, dw, layout=0):
    z = np.zeros(4 * dw)
    lib.orc_tile_normals(seed, chain, gtile, it, layout, dw, z.ctypes.data_as(_dp))
    return z


def gtile0_of(n):
    """global tile offset of each interval: tiles of 4 EM steps, never straddling an interval."""
    g = np.zeros(len(n) + 1, dtype=np.int32)
    for k, nk in enumerate(n):
        g[k + 1] = g[k] + (nk - 1 + 3) // 4
    return g


class Pair:
    """One recording (SamplingPair) of the oracle."""

    def __init__(self, lib, model, n, t, m, eps=1e-11):
        self.lib = lib
        self.model = model
        self.n = np.ascontiguousarray(n, dtype=np.int32)
        self.K = len(self.n)
        self.t = np.ascontiguousarray(t, dtype=np.float64)
        assert self.t.size == int(self.n.sum())
        self.off = np.concatenate([[0], np.cumsum(self.n)])
        self.d, self.dw, self.npar, self.constdiff = model_dims(lib, model)
        self.m = m
        self.gtile0 = gtile0_of(self.n)
        self.h = lib.orc_pair_create(model, self.K, self.n.ctypes.data_as(_ip), self.t.ctypes.data_as(_dp), m, eps)
        assert self.h

    def __del__(self):
        try:
            self.lib.orc_pair_destroy(self.h)
        except Exception:
            pass

    def tk(self, k):
        return self.t[self.off[k]:self.off[k + 1]]

    def set_theta(self, theta, side=None, store=None, k=None):
        th, p = _d(theta)
        for s in ([0, 1] if side is None else [side]):
            for st in ([0, 1] if store is None else [store]):
                for kk in (range(self.K) if k is None else [k]):
                    self.lib.orc_set_theta(self.h, s, st, kk, p)

    def set_aux(self, k, B, beta, at, side=None, store=None):
        B_, Bp = _d(B); b_, bp = _d(beta); a_, ap = _d(at)
        for s in ([0, 1] if side is None else [side]):
            for st in ([0, 1] if store is None else [store]):
                self.lib.orc_set_aux(self.h, s, st, k, Bp, bp, ap)

    def set_obs(self, k, L, Sig, v, side=None):
        L_, Lp = _d(L); S_, Sp = _d(Sig); v_, vp = _d(v)
        for s in ([0, 1] if side is None else [side]):
            self.lib.orc_set_obs(self.h, s, k, Lp, Sp, vp)

    def set_start(self, x0):
        x, p = _d(x0)
        self.lib.orc_set_start(self.h, p)

    def set_W(self, side, k, dW):
        a, p = _d(dW)
        assert a.size == (self.n[k] - 1) * self.dw
        self.lib.orc_set_W(self.h, side, k, p)

    def set_X(self, side, k, X):
        a, p = _d(X)
        assert a.size == self.n[k] * self.d
        self.lib.orc_set_X(self.h, side, k, p)

    def get_W(self, side, k):
        a = np.zeros((self.n[k] - 1, self.dw))
        self.lib.orc_get_W(self.h, side, k, a.ctypes.data_as(_dp))
        return a

    def get_X(self, side, k):
        a = np.zeros((self.n[k], self.d))
        self.lib.orc_get_X(self.h, side, k, a.ctypes.data_as(_dp))
        return a

    def get_HFc(self, side, store, k):
        n, d = int(self.n[k]), self.d
        H = np.zeros((n, d, d)); F = np.zeros((n, d)); c = np.zeros(n)
        self.lib.orc_get_HFc(self.h, side, store, k, H.ctypes.data_as(_dp), F.ctypes.data_as(_dp), c.ctypes.data_as(_dp))
        return H, F, c

    def set_HFc(self, side, store, k, H, F, c):
        H_, Hp = _d(H); F_, Fp = _d(F); c_, cp = _d(c)
        self.lib.orc_set_HFc(self.h, side, store, k, Hp, Fp, cp)

    # --- BiBlock methods -------------------------------------------------------------------------------
    def biblock(self, i0, i1, last, rho=0.0):
        bb = BiBlock()
        bb.i0, bb.i1, bb.last, bb.rho = i0, i1, int(last), rho
        bb.ll[0] = bb.ll[1] = -np.inf
        return bb

    def set_artificial_obs(self, bb):
        self.lib.orc_set_artificial_obs(self.h, C.byref(bb))

    def recompute_guiding_term(self, bb, side=0):
        self.lib.orc_recompute_guiding_term(self.h, C.byref(bb), side)

    def recompute_guiding_term_tsit5(self, bb, side=0, reltol=1e-3, abstol=1e-6):
        """upstream's solver (adaptive Tsit5, OrdinaryDiffEq default tolerances); returns the number of accepted steps"""
        return self.lib.orc_recompute_guiding_term_tsit5(self.h, C.byref(bb), side, reltol, abstol)

    def find_W_for_X(self, bb):
        self.lib.orc_find_W_for_X(self.h, C.byref(bb))

    def loglikhd(self, bb, side=0, skip=0):
        return self.lib.orc_loglikhd(self.h, C.byref(bb), side, skip)

    def draw_proposal_path(self, bb, Z=None, seed=0, chain=0, it=0, layout=0):
        if Z is not None:
            Z_, zp = _d(Z)
        else:
            zp = None
        return bool(self.lib.orc_draw_proposal_path(self.h, C.byref(bb), zp, seed, chain, it, layout,
                                                    self.gtile0.ctypes.data_as(_ip)))

    def recompute_path(self, bb, law_side=1, w_side=0, skip=0):
        return bool(self.lib.orc_recompute_path(self.h, C.byref(bb), law_side, w_side, skip))

    def accept_reject(self, bb, E):
        hist = np.zeros(2)
        acc = self.lib.orc_accept_reject(self.h, C.byref(bb), E, hist.ctypes.data_as(_dp))
        return bool(acc), hist

    def swap_XX(self, bb):
        self.lib.orc_swap_XX(self.h, C.byref(bb))

    def swap_WW(self, bb):
        self.lib.orc_swap_WW(self.h, C.byref(bb))

    def swap_PP(self, bb):
        self.lib.orc_swap_PP(self.h, C.byref(bb))

    def swap_ll(self, bb):
        self.lib.orc_swap_ll(C.byref(bb))
