"""CPU: host-side logic of the Gen-API mirror (choicemaps, address schema, parameter packing)."""
import numpy as np
import pytest

import gen_b200 as g
from oracle import closed_forms as cf


def test_choicemap_addresses_and_merge():
    cm = g.choicemap(("x_init", 1), (("chain", 1, "x"), 2))
    assert cm["x_init"] == 1 and cm[("chain", 1, "x")] == 2
    assert ("chain", 2, "x") not in cm and not cm.isempty()
    with pytest.raises(KeyError):
        cm[("chain", 2, "x")]
    merged = g.merge(cm, g.choicemap((("chain", 1, "z"), 3)))
    assert len(merged) == 3
    with pytest.raises(ValueError):
        g.merge(cm, g.choicemap(("x_init", 5)))
    cm2 = g.choicemap()
    cm2.set_value("y", 2.0)
    assert cm2.get_value("y") == 2.0


def test_hmm_address_schema_matches_reference_test():
    """test/inference/particle_filter.jl:83-94: :z_init, :x_init, :chain => t => :z / :x."""
    m = g.HMM(cf.HMM_PRIOR, cf.HMM_EMISSION, cf.HMM_TRANSITION)
    assert m.obs_address(1) == "x_init" and m.obs_address(3) == ("chain", 2, "x")
    assert m.latent_address(1, "z") == "z_init" and m.latent_address(4, "z") == ("chain", 3, "z")
    assert np.array_equal(m.extract_observations(2, g.choicemap((("chain", 1, "x"), 3))), [3.0])
    with pytest.raises(g.GsmcError):
        m.extract_observations(2, g.choicemap((("chain", 5, "x"), 3)))
    with pytest.raises(g.GsmcError):
        m.extract_observations(2, g.choicemap((("chain", 1, "x"), 3), ("bogus", 1)))


def test_hmm_param_packing_matches_oracle_layout():
    m = g.HMM(cf.HMM_PRIOR, cf.HMM_EMISSION, cf.HMM_TRANSITION)
    assert np.array_equal(m.params(), cf.hmm_params())
    p = m.params()
    K = int(p[0])
    trans = p[2 + K:2 + K + K * K].reshape(K, K)
    assert np.allclose(trans.sum(axis=1), 1.0)                      # rows indexed by z_prev are distributions
    assert np.allclose(trans[2], [0.9, 0.05, 0.05])                 # transition_dists[:, 3] of the reference test


def test_regression_binding_and_observations():
    m = g.LinearRegression().bind(cf.QUICKSTART_XS)
    assert np.array_equal(m.params(), cf.regression_params())
    obs = g.choicemap(*[("y-%d" % (i + 1), y) for i, y in enumerate(cf.QUICKSTART_YS)])
    assert np.array_equal(m.extract_observations(1, obs), cf.QUICKSTART_YS)
    with pytest.raises(g.GsmcError):
        m.extract_observations(1, g.choicemap(("y-1", 0.0)))


def test_api_surface_matches_reference_exports():
    """src/inference/particle_filter.jl:215-216 and importance.jl:110."""
    for name in ("initialize_particle_filter", "particle_filter_step_b", "maybe_resample_b", "get_traces",
                 "get_log_weights", "log_ml_estimate", "sample_unweighted_traces", "importance_sampling", "importance_resampling"):
        assert callable(getattr(g, name))
