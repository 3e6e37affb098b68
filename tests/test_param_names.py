"""Host logic: the Python restatement of /root/reference/src/param_names_collections.jl (no GPU, no oracle)."""
import numpy as np
import pytest

import dmt_b200
from dmt_b200 import _lib
from dmt_b200 import param_names as pn

FHN_NAMES = pn.TARGET_PARAM_NAMES[_lib.FHN]


def test_updt_follows_pdep_order_and_indexes_theta_names():
    # src/param_names_collections.jl:85-100: filter pdep by membership in θnames, replace the global name by its index in θnames
    pdep = [("g_shared", "gamma"), ("b", "beta"), ("s_rec1", "sigma")]
    updt = pn.find_theta_names_for_MCMC_update(["s_rec1", "g_shared"], pdep)
    assert updt == ((1, "gamma"), (0, "sigma"))


def test_unit_var_is_the_complement_of_updt():
    u = pn.ParamNamesUnit.build(2, FHN_NAMES, ("gamma", "sigma"), ["g"], [("g", "gamma")], [(), ()])
    assert u.updt == ((0, "gamma"),)
    assert u.var == ("eps", "s", "beta", "sigma")
    assert u.updt_aux == [((0, "gamma"),), ((0, "gamma"),)]       # gamma is held by the auxiliary law too
    assert u.var_aux == [("sigma",), ("sigma",)]
    assert u.tuple_lengths() == (4, 1)
    # a parameter the auxiliary law does not hold is not a critical one
    u2 = pn.ParamNamesUnit.build(1, FHN_NAMES, ("sigma",), ["g"], [("g", "gamma")], [()])
    assert u2.updt_aux == [()] and u2.var_aux == [("sigma",)]


def test_empty_collection_has_no_var_names():
    # find_var_names_not_in_MCMC_update returns tuple() for an empty collection (src/param_names_collections.jl:151)
    u = pn.ParamNamesUnit.build(0, FHN_NAMES, FHN_NAMES, ["g"], [("g", "gamma")], [])
    assert u.var == () and u.updt == ((0, "gamma"),) and u.updt_aux == [] and u.updt_obs == []


def test_block_index_split_terminal_and_non_terminal():
    odeps = [(("o%d" % k, 0),) for k in range(6)]
    theta_names = ["o1", "o2", "o5", "g"]
    pdep = [("g", "gamma")]
    nb = pn.ParamNamesBlock.build((0, 2, False), FHN_NAMES, FHN_NAMES, theta_names, pdep, odeps)
    # non-terminal block over intervals 0..2: PP = 0,1 ; P_last / P_excl = 2 (src/param_names_collections.jl:218-221, src/block.jl:66-72)
    assert nb.idx_PP == (0, 1) and nb.idx_excl == (2,)
    assert nb.PP.updt_obs == [(), ((0, 0),)]                      # o1 -> index 0 of θnames, obs.θ slot 0
    assert nb.P_excl.updt_obs == [((1, 0),)]                      # the real observation of interval 2
    assert nb.P_last.updt_obs == [()] and nb.Pb_excl.updt_obs == [(), ()]   # artificial observations carry no parameters
    tb = pn.ParamNamesBlock.build((3, 5, True), FHN_NAMES, FHN_NAMES, theta_names, pdep, odeps)
    assert tb.idx_PP == (3, 4, 5) and tb.idx_excl == ()
    assert tb.PP.updt_obs == [(), (), ((2, 0),)] and tb.P_last.var == () and tb.P_excl.updt_obs == []
    assert nb.tuple_lengths() == tb.tuple_lengths() == (4, 1)


def test_all_obs_mixed_effects_mapping():
    # two recordings share beta, each has its own gamma: θ° = (gamma_1, gamma_2, beta)
    theta_names = ["gamma_1", "gamma_2", "beta"]
    pdr = [[("gamma_1", "gamma"), ("beta", "beta")], [("gamma_2", "gamma"), ("beta", "beta")]]
    pa = pn.ParamNamesAllObs.from_layout(_lib.FHN, [(0, 1), (2, 3)], 4, 2, theta_names, pdr)
    assert len(pa.recordings) == 2 and len(pa.recordings[0].blocks) == 2
    assert pa.recordings[0].blocks[0].PP.updt == ((0, "gamma"), (2, "beta"))
    assert pa.recordings[1].blocks[1].PP.updt == ((1, "gamma"), (2, "beta"))
    fl = pa.flat_updates(_lib.FHN)
    assert fl == [{2: 0, 3: 2}, {2: 1, 3: 2}]                     # model vector (eps, s, gamma, beta, sigma)
    assert pa.is_critical()                                       # gamma and beta enter the linearised auxiliary law
    assert pa.obs_updates() == [{}, {}]
    # an update of a parameter no auxiliary law holds is not critical
    pb = pn.ParamNamesAllObs.from_layout(_lib.FHN, [(0, 3)], 4, 2, ["sig"], [[("sig", "sigma")]] * 2, aux_names=("eps", "s", "gamma", "beta"))
    assert not pb.is_critical()
    with pytest.raises(ValueError):
        pn.ParamNamesAllObs.from_layout(_lib.FHN, [(0, 3)], 4, 3, ["sig"], [[("sig", "sigma")]] * 2)


def test_obs_updates_are_keyed_by_interval():
    odr = [[(), (("noise", 1),), (), (("noise", 1), ("gain", 0))]]
    pa = pn.ParamNamesAllObs.from_layout(_lib.LORENZ, [(0, 1), (2, 3)], 4, 1, ["gain", "noise"], [[]], odr)
    # block 0 = intervals 0,(1 as P_last/P_excl); block 1 = 2,3
    assert pa.obs_updates() == [{1: ((1, 1),), 3: ((1, 1), (0, 0))}]
    assert pa.is_critical()


def test_heterogeneous_recordings_are_bucketed_by_model_grid_and_observation_operator():
    from dmt_b200 import hetero
    tt_a = (np.array([3, 3], np.int32), np.array([0.0, 0.05, 0.1, 0.1, 0.15, 0.2]))
    tt_b = (np.array([3, 3], np.int32), np.array([0.0, 0.04, 0.1, 0.1, 0.16, 0.2]))          # same counts, different grid
    L1, L2, S = np.array([[1.0, 0.0]]), np.array([[0.0, 1.0]]), np.array([[0.1]])
    mk = lambda model, tts, L: dict(model=model, tts=tts, L=L, Sigma=S)
    recs = [mk(_lib.FHN, tt_a, L1), mk(_lib.LV, tt_a, L1), mk(_lib.FHN, tt_b, L1), mk(_lib.FHN, tt_a, L1), mk(_lib.FHN, tt_a, L2),
            mk(_lib.LV, tt_a, L1), mk(_lib.FHN, tt_b, L1)]
    assert hetero.bucket_recordings(recs) == [[0, 3], [1, 5], [2, 6], [4]]
    assert hetero.bucket_recordings([]) == []
