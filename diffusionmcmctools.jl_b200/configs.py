"""Synthetic instances of the BASELINE.json configs: model constants, imputation grids, simulated
observations, linearisation points and block layouts.  Host-side set-up only (numpy); nothing here is on the hot path.

The tutorials' set-up it mirrors: /root/reference/docs/src/tutorials/preamble.md:77-89 (simulate the target on a fine
grid, observe every 0.1 with Gaussian noise) and docs/src/tutorials/biblock/smoothing.md:27
(`OBS.setup_time_grids(recording, dt, standard_guid_prop_time_transf)`).
"""
from dataclasses import dataclass, field

import numpy as np

from . import _lib

FHN, LV, LORENZ, PROK, JR, OU2 = range(6)

THETA = {
    FHN: [0.1, -0.8, 1.5, 0.0, 0.3],
    LV: [2 / 3, 4 / 3, 1.0, 1.0, 0.1, 0.1],
    LORENZ: [10.0, 28.0, 8 / 3, 3.0],
    PROK: [0.1, 0.7, 0.35, 0.2, 0.1, 0.9, 0.3, 0.1, 10.0],
    JR: [3.25, 100.0, 22.0, 50.0, 135.0, 5.0, 6.0, 0.56, 220.0, 2000.0],
    OU2: [-0.5, 0.3, -0.2, -1.0, 0.1, -0.3, 0.4, 0.7],
}
X0 = {FHN: [-0.9, -1.0], LV: [2.0, 0.25], LORENZ: [1.5, -1.5, 25.0], PROK: [8.0, 8.0, 8.0, 5.0],
      JR: [0.08, 18.0, 15.0, -0.5, 0.0, 0.0], OU2: [0.3, -0.2]}
OBS = {  # (L, Sigma)
    FHN: (np.array([[1.0, 0.0]]), np.array([[0.01]])),
    LV: (np.eye(2), 0.01 * np.eye(2)),
    LORENZ: (np.eye(2, 3), 0.5 * np.eye(2)),
    PROK: (np.eye(4), 2.0 * np.eye(4)),
    JR: (np.array([[0.0, 1.0, -1.0, 0.0, 0.0, 0.0]]), np.array([[1e-2]])),
    OU2: (np.eye(2), 0.02 * np.eye(2)),
}
DIMS = {FHN: (2, 1), LV: (2, 2), LORENZ: (3, 3), PROK: (4, 4), JR: (6, 1), OU2: (2, 2)}


def tau_grid(t0, t1, dt):
    """standard_guid_prop_time_transf on a uniform grid: tau(s) = t0 + (s-t0)(2 - (s-t0)/T)   """
    n = int(round((t1 - t0) / dt)) + 1
    s = np.linspace(0.0, t1 - t0, n)
    return t0 + s * (2.0 - s / (t1 - t0))


# ---- vectorised numpy drift / diffusion, used ONLY to simulate synthetic observations -------------------------------
def _sigm(th, v):
    return th[5] / (1.0 + np.exp(th[7] * (th[6] - v)))


def drift(model, th, x):
    """x: [d, M] -> b: [d, M]"""
    if model == FHN:
        return np.stack([(x[0] - x[1] - x[0] ** 3 + th[1]) / th[0], th[2] * x[0] - x[1] + th[3]])
    if model == LV:
        return np.stack([th[0] * x[0] - th[1] * x[0] * x[1], th[3] * x[0] * x[1] - th[2] * x[1]])
    if model == LORENZ:
        return np.stack([th[0] * (x[1] - x[0]), th[1] * x[0] - x[1] - x[0] * x[2], x[0] * x[1] - th[2] * x[2]])
    if model == PROK:
        h = _prok_h(th, x)
        return np.stack([h[2] - h[6], h[3] - 2 * h[4] + 2 * h[5] - h[7], -h[0] + h[1] + h[4] - h[5], -h[0] + h[1]])
    if model == JR:
        A, a, B, b, C, mu = th[0], th[1], th[2], th[3], th[4], th[8]
        return np.stack([x[3], x[4], x[5],
                         A * a * _sigm(th, x[1] - x[2]) - 2 * a * x[3] - a * a * x[0],
                         A * a * (mu + 0.8 * C * _sigm(th, C * x[0])) - 2 * a * x[4] - a * a * x[1],
                         B * b * 0.25 * C * _sigm(th, 0.25 * C * x[0]) - 2 * b * x[5] - b * b * x[2]])
    if model == OU2:
        return np.stack([th[0] * x[0] + th[1] * x[1] + th[4], th[2] * x[0] + th[3] * x[1] + th[5]])
    raise ValueError(model)


def _prok_h(c, x):
    return [c[0] * x[3] * x[2], c[1] * (c[8] - x[3]), c[2] * x[3], c[3] * x[0], c[4] * x[1] * (x[1] - 1) * 0.5,
            c[5] * x[2], c[6] * x[0], c[7] * x[1]]


def noise(model, th, x, dW):
    """sigma(x) dW for dW: [dw, M]"""
    if model == FHN:
        return np.stack([np.zeros_like(dW[0]), th[4] * dW[0]])
    if model == LV:
        return np.stack([th[4] * dW[0], th[5] * dW[1]])
    if model == LORENZ:
        return th[3] * dW
    if model == PROK:  # per-reaction noise (dW has 8 components here): S diag(sqrt h) dB has covariance a(x) dt
        h = np.maximum(np.array(_prok_h(th, x)), 0.0)
        S = np.array([[0, 0, 1, 0, 0, 0, -1, 0], [0, 0, 0, 1, -2, 2, 0, -1], [-1, 1, 0, 0, 1, -1, 0, 0], [-1, 1, 0, 0, 0, 0, 0, 0]], float)
        return S @ (np.sqrt(h) * dW)
    if model == JR:
        out = np.zeros((6,) + dW.shape[1:])
        out[4] = th[9] * dW[0]
        return out
    if model == OU2:
        return np.stack([th[6] * dW[0], th[7] * dW[1]])
    raise ValueError(model)



def clamp(model, th, x):
    """keep simulated synthetic truth inside the model's domain (LV, Prokaryote)"""
    if model == LV:
        return np.maximum(x, 0.05)
    if model == PROK:
        x = np.maximum(x, [[0.5], [1.5], [0.5], [0.5]])
        x[3] = np.minimum(x[3], th[8] - 0.5)
        return x
    return x


@dataclass
class Problem:
    name: str
    model: int
    theta: np.ndarray            # [npar]
    n_pts: np.ndarray            # [K]
    tt: np.ndarray               # [sum n_pts]
    tobs: np.ndarray             # [K]
    M: int
    P: int
    m: int
    L: np.ndarray                # [m, d]
    Sigma: np.ndarray            # [m, m]
    v: np.ndarray                # [K, m, P]
    xbar: np.ndarray             # [K, d, P]  linearisation points of the auxiliary laws
    x0: np.ndarray               # [d, M]
    layouts: list = field(default_factory=list)   # [(ranges, rho), ...]; ranges 0-based inclusive
    eps: float = 1e-11
    pset_of_chain: np.ndarray = None

    @property
    def K(self):
        return len(self.n_pts)

    @property
    def steps_per_chain(self):
        return int((self.n_pts - 1).sum())

    @property
    def d(self):
        return DIMS[self.model][0]

    @property
    def dw(self):
        return DIMS[self.model][1]


def make_problem(name, M, P=None, K=None, obs_dt=None, dt=None, seed=0, layouts=None, rho=0.9, chain_offset=0, sim_sub=2, theta=None):
    """Synthetic instance of one of the named configs.  Per-pset data are simulated with a numpy generator seeded by the
    GLOBAL pset index (chain_offset + p), so a sharded ensemble sees the same data as the unsharded one."""
    spec = {
        "fhn": (FHN, 10, 0.1, 1e-3), "lv": (LV, 50, 0.1, 1e-3), "lorenz": (LORENZ, 200, 0.1, 1e-3),
        "prok": (PROK, 100, 0.1, 1e-3), "jr": (JR, 500, 0.01, 1e-4), "ou2": (OU2, 5, 0.1, 1e-3),
    }[name]
    model = spec[0]
    K = spec[1] if K is None else K
    obs_dt = spec[2] if obs_dt is None else obs_dt
    dt = spec[3] if dt is None else dt
    P = M if P is None else P
    d, dw = DIMS[model]
    th = np.array(THETA[model] if theta is None else theta, dtype=np.float64)
    L, Sigma = OBS[model]
    m = L.shape[0]
    tobs = obs_dt * np.arange(1, K + 1)
    grids = [tau_grid(tobs[k] - obs_dt, tobs[k], dt) for k in range(K)]
    n_pts = np.array([len(g) for g in grids], dtype=np.int32)
    tt = np.concatenate(grids)

    # ---- simulate the target per pset (Euler–Maruyama on a grid sim_sub times finer than the imputation grid)
    x0p = np.repeat(np.array(X0[model], dtype=np.float64)[:, None], P, axis=1)
    if model == LV:
        rng0 = np.random.default_rng([seed, 17])
        x0p = x0p * (1.0 + 0.05 * rng0.normal(size=(P + chain_offset, d)).T[:, chain_offset:])
    h = dt / sim_sub
    nsub = int(round(obs_dt / h))
    gens = [np.random.default_rng([seed, chain_offset + p]) for p in range(P)] if P <= 64 else None
    # (large ensembles: one generator per observation interval, keyed by the interval, filled chain-major up to column chain_offset + P
    # with the shard's columns kept — so a rank of a sharded run sees exactly its slice of the unsharded ensemble's data)
    PT = chain_offset + P
    x = x0p.copy()
    v = np.empty((K, m, P))
    xbar = np.empty((K, d, P))
    Lc = np.linalg.cholesky(Sigma)
    nw_sim = 8 if model == PROK else dw
    for k in range(K):
        if gens is None:
            dW_k = np.random.default_rng([seed, 1_000_003, k]).normal(size=(PT, nsub, nw_sim))[chain_offset:] * np.sqrt(h)
        for jsub in range(nsub):
            if gens is not None:
                dWn = np.stack([g.normal(size=nw_sim) for g in gens], axis=1) * np.sqrt(h)
            else:
                dWn = dW_k[:, jsub, :].T
            x = clamp(model, th, x + drift(model, th, x) * h + noise(model, th, x, dWn))
        if gens is not None:
            eta = np.stack([g.normal(size=m) for g in gens], axis=1)
        else:
            eta = np.random.default_rng([seed, 1_000_004, k]).normal(size=(PT, m))[chain_offset:].T
        v[k] = L @ x + Lc @ eta
        if model in (LV, PROK):  # noisy observations of a positive state: keep them inside the law's domain, otherwise the guided
            v[k] = clamp(model, th, v[k]) if m == d else v[k]   # proposal is pulled across the boundary and never succeeds
        xbar[k] = x
    if model == FHN:  # the upstream FitzHughNagumoAux linearises at the observed v (x2 does not enter the Jacobian)
        xbar[:, 0, :] = v[:, 0, :]
    if P == M:
        x0 = x0p
        pset_of_chain = None
    else:
        pset_of_chain = (np.arange(M) % P).astype(np.int32)
        x0 = x0p[:, pset_of_chain]
    if layouts is None:
        layouts = [([(0, K - 1)], rho)]
    return Problem(name, model, th, n_pts, tt, tobs, M, P, m, L, Sigma, v, xbar, x0, layouts, 1e-11, pset_of_chain)


def blocking_layouts(K, block_len, rho):
    """two staggered layouts (docs/src/tutorials/biblock/smoothing_with_blocking.md:69 pattern): A tiles 0..K-1 with blocks of
    block_len intervals, B is shifted by block_len/2."""
    A = [(i, min(i + block_len, K) - 1) for i in range(0, K, block_len)]
    half = block_len // 2
    B = [(0, half - 1)] + [(i, min(i + block_len, K) - 1) for i in range(half, K, block_len)]
    return [(A, rho), (B, rho)]


def named_config(cfg, M=None, seed=0, chain_offset=0, P=None, sim_sub=2):
    """BASELINE.json configs: c1..c5 (sim_sub: refinement of the grid the synthetic truth is simulated on)"""
    cfg = cfg.lower()
    import functools
    make_problem = functools.partial(globals()["make_problem"], sim_sub=sim_sub)
    if cfg == "c1":
        return make_problem("fhn", M or 1, P, seed=seed, rho=0.96, chain_offset=chain_offset)
    if cfg == "c2":
        return make_problem("lv", M or 1024, P, seed=seed, rho=0.9, chain_offset=chain_offset)
    if cfg == "c3" and P is not None and P != (M or 4096):
        # shared data / shared guiding term (SURVEY §8d "report both P=1 and P=M"): blocking freezes a per-recording artificial
        # observation, so the P < M variant is the single terminal block 1:200 with the same pCN rho
        return make_problem("lorenz", M or 4096, P, seed=seed, rho=0.9, chain_offset=chain_offset)
    if cfg == "c3":
        return make_problem("lorenz", M or 4096, P, seed=seed, layouts=blocking_layouts(200, 20, 0.9), chain_offset=chain_offset)
    if cfg == "c4":
        return make_problem("prok", M or 16384, P, seed=seed, rho=0.9, chain_offset=chain_offset)
    if cfg == "c5":
        return make_problem("jr", M or 8192, P, seed=seed, rho=0.9, chain_offset=chain_offset)
    raise ValueError("unknown config %r" % cfg)


def upload(prob, ctx, sides=(0,)):
    """Push a Problem into a device context: parameters, observations, auxiliary laws (PP and PPb: aux_laws_blocking defaults
    to aux_laws, src/sampling_unit.jl:57), start points and block layouts."""
    for side in sides:
        ctx.set_params(prob.theta, side=side, stores=3)
        ctx.set_obs(prob.L, prob.Sigma, prob.v, side=side)
        ctx.set_aux_linearised(prob.xbar, side=side, store=_lib.STORE_PP)
        ctx.set_aux_linearised(prob.xbar, side=side, store=_lib.STORE_PPB)
    ctx.set_start(prob.x0)
    for i, (ranges, rho) in enumerate(prob.layouts):
        ctx.set_blocks(i, ranges, rho)
