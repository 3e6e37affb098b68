"""Host-side mirror of /root/reference/src/param_names_collections.jl — the caller's side of set_proposal_law!.

The reference keeps, per MCMC update step, a tree  ParamNamesAllObs -> ParamNamesRecording -> ParamNamesBlock -> ParamNamesUnit
that says, for every collection of laws of every block of every recording, which entry of the proposed vector θ° goes into which
parameter of the target law, of each auxiliary law and of each observation, and which parameters must instead be kept equal to
the accepted law's.  This stays on the host (BASELINE.json north_star); this module restates it over plain Python data and
turns it into the flat per-recording parameter array that `dmt_set_params` uploads.

Differences forced by the language, all documented where they occur: symbols are strings, indices are 0-based, and the law
objects the reference inspects with `DD.var_parameter_names` are replaced by the parameter-name tables below.
"""
from dataclasses import dataclass, field

from . import _lib

# names of the entries of each compiled model's parameter vector (csrc/models.cuh, struct Par of each model), in order
TARGET_PARAM_NAMES = {
    _lib.FHN: ("eps", "s", "gamma", "beta", "sigma"),
    _lib.LV: ("alpha", "beta", "gamma", "delta", "sigma1", "sigma2"),
    _lib.LORENZ: ("theta1", "theta2", "theta3", "sigma"),
    _lib.PROK: ("c1", "c2", "c3", "c4", "c5", "c6", "c7", "c8", "K"),
    _lib.JR: ("A", "a", "B", "b", "C", "nu_max", "v0", "r", "mu", "sigma_y"),
    _lib.OU2: ("B11", "B12", "B21", "B22", "beta1", "beta2", "sigma1", "sigma2"),
}
# the linearised auxiliary laws (aux_linearise_kernel) are functions of the same parameter vector: every target name is also an
# auxiliary-law name.  A user-defined auxiliary law (dmt_set_aux) passes its own tuple of names.
AUX_PARAM_NAMES = dict(TARGET_PARAM_NAMES)


def find_theta_names_for_MCMC_update(theta_names, pdep):
    """src/param_names_collections.jl:85-100: pdep entries (global name, law name) whose global name is being updated, as
    (index into θ°, law name), in pdep order"""
    return tuple((theta_names.index(g), loc) for g, loc in pdep if g in theta_names)


def find_theta_aux_names_for_MCMC_update(updt, aux_names_per_law):
    """:109-114: per law of the collection, the updt entries that the law's auxiliary law also holds"""
    return [tuple(p for p in updt if p[1] in names) for names in aux_names_per_law]


def find_theta_obs_idx_for_MCMC_update(theta_names, odeps):
    """:125-141: per observation, (index into θ°, index into obs.θ)"""
    return [tuple((theta_names.index(g), j) for g, j in odep if g in theta_names) for odep in odeps]


def find_var_names_not_in_MCMC_update(updt, target_names, n_laws):
    """:149-153"""
    if n_laws == 0:
        return tuple()
    in_updt = {p[1] for p in updt}
    return tuple(n for n in target_names if n not in in_updt)


def find_var_aux_names_not_in_MCMC_update(updt_aux, aux_names_per_law):
    """:162-171"""
    return [tuple(n for n in names if n not in {p[1] for p in ua}) for ua, names in zip(updt_aux, aux_names_per_law)]


@dataclass
class ParamNamesUnit:
    """src/param_names_collections.jl:49-73.  One collection of laws (PP, P_last, P_excl or Pb_excl) of one block."""
    var: tuple = ()
    var_aux: list = field(default_factory=list)
    updt: tuple = ()
    updt_aux: list = field(default_factory=list)
    updt_obs: list = field(default_factory=list)

    @classmethod
    def build(cls, n_laws, target_names, aux_names, theta_names, pdep, odeps):
        aux_per_law = [aux_names] * n_laws
        updt = find_theta_names_for_MCMC_update(theta_names, pdep)
        updt_aux = find_theta_aux_names_for_MCMC_update(updt, aux_per_law)
        updt_obs = find_theta_obs_idx_for_MCMC_update(theta_names, odeps)
        var = find_var_names_not_in_MCMC_update(updt, target_names, n_laws)
        var_aux = find_var_aux_names_not_in_MCMC_update(updt_aux, aux_per_law)
        return cls(var, var_aux, updt, updt_aux, updt_obs)

    def tuple_lengths(self):
        return len(self.var), len(self.updt)


def _idx_split(i0, i1, last):
    """src/param_names_collections.jl:218-226 with the block as an inclusive 0-based interval range [i0, i1]: a terminal block's
    PP covers all of it; a non-terminal block's PP stops one short and the last interval is P_last / P_excl (src/block.jl:66-72)"""
    if last:
        return list(range(i0, i1 + 1)), []
    return list(range(i0, i1)), [i1]


@dataclass
class ParamNamesBlock:
    """src/param_names_collections.jl:201-216"""
    PP: ParamNamesUnit
    P_last: ParamNamesUnit
    P_excl: ParamNamesUnit
    Pb_excl: ParamNamesUnit
    idx_PP: tuple = ()      # (not in the reference struct) the observation intervals PP / Pb_excl and P_last / P_excl refer to
    idx_excl: tuple = ()

    @classmethod
    def build(cls, block_range, target_names, aux_names, theta_names, pdep, odeps):
        i0, i1, last = block_range
        idx1, idx2 = _idx_split(i0, i1, last)
        mk = lambda n, od: ParamNamesUnit.build(n, target_names, aux_names, theta_names, pdep, od)
        return cls(PP=mk(len(idx1), [odeps[k] for k in idx1]), P_last=mk(len(idx2), [tuple() for _ in idx2]),
                   P_excl=mk(len(idx2), [odeps[k] for k in idx2]), Pb_excl=mk(len(idx1), [tuple() for _ in idx1]),
                   idx_PP=tuple(idx1), idx_excl=tuple(idx2))

    def tuple_lengths(self):
        return self.PP.tuple_lengths()


@dataclass
class ParamNamesRecording:
    """src/param_names_collections.jl:243-251"""
    blocks: list

    @classmethod
    def build(cls, block_ranges, target_names, aux_names, theta_names, pdep, odeps):
        return cls([ParamNamesBlock.build(r, target_names, aux_names, theta_names, pdep, odeps) for r in block_ranges])


@dataclass
class ParamNamesAllObs:
    """src/param_names_collections.jl:268-288.  `param_depend_rev[i]` / `obs_depend_rev[i]` are recording i's
    (global name, law name) pairs and, per observation, (global name, index into obs.θ) pairs — the two fields of the
    reference's `all_obs` (ObservationSchemes.AllObservations) that the constructor reads."""
    recordings: list

    @classmethod
    def build(cls, be, theta_names, param_depend_rev, obs_depend_rev=None, aux_names=None):
        """ParamNamesAllObs(be::BlockEnsemble, θnames, all_obs)"""
        return cls.from_layout(be.se.model, be.ranges, be.se.ctx.K, be.se.M_total, theta_names, param_depend_rev, obs_depend_rev, aux_names)

    @classmethod
    def from_layout(cls, model, block_ranges, K, n_rec, theta_names, param_depend_rev, obs_depend_rev=None, aux_names=None):
        tn = TARGET_PARAM_NAMES[model]
        an = tuple(aux_names) if aux_names is not None else AUX_PARAM_NAMES[model]
        nb = len(block_ranges)
        ranges = [(int(a), int(b), j == nb - 1) for j, (a, b) in enumerate(block_ranges)]  # the last block is terminal, src/block_collection.jl:29
        if len(param_depend_rev) != n_rec:
            raise ValueError("param_depend_rev must hold one entry per recording (%d), got %d" % (n_rec, len(param_depend_rev)))
        recs = []
        for i in range(n_rec):
            od = obs_depend_rev[i] if obs_depend_rev is not None else [tuple() for _ in range(K)]
            recs.append(ParamNamesRecording.build(ranges, tn, an, list(theta_names), list(param_depend_rev[i]), od))
        return cls(recs)

    # ---- what the device needs from the tree ---------------------------------------------------------------------------
    def is_critical(self):
        """An update is critical (the proposal law's guiding term must be recomputed) when any updated parameter is held by an
        auxiliary law or by an observation.  (The reference's own GP.is_critical_update(bb, pnames) reads fields that no
        ParamNames struct has, src/biblock.jl:315-317; this is what its docstring describes.)"""
        for r in self.recordings:
            for b in r.blocks:
                for u in (b.PP, b.P_last, b.P_excl, b.Pb_excl):
                    if any(len(x) for x in u.updt_aux) or any(len(x) for x in u.updt_obs):
                        return True
        return False

    def flat_updates(self, model):
        """per recording: {index into the model's parameter vector: index into θ°} (union over the recording's blocks; the
        reference applies the same pdep to every block of a recording, src/param_names_collections.jl:246)"""
        tn = TARGET_PARAM_NAMES[model]
        out = []
        for r in self.recordings:
            m = {}
            for b in r.blocks:
                for u in (b.PP, b.P_last, b.P_excl, b.Pb_excl):
                    for i_src, name in u.updt:
                        m[tn.index(name)] = i_src
            out.append(m)
        return out

    def obs_updates(self):
        """per recording: {observation interval: ((index into θ°, index into obs.θ), ...)} (PP and P_excl carry them)"""
        out = []
        for r in self.recordings:
            per_k = {}
            for b in r.blocks:
                for k, pairs in zip(b.idx_PP, b.PP.updt_obs):
                    if pairs:
                        per_k[k] = tuple(pairs)
                for k, pairs in zip(b.idx_excl, b.P_excl.updt_obs):
                    if pairs:
                        per_k[k] = tuple(pairs)
            out.append(per_k)
        return out
