"""CPU: the oracle against every golden value the reference's own tests hold for this path, against
closed forms, and against committed fixtures (tests/golden). No GPU needed."""
import json
import math
import os

import numpy as np
import pytest

from oracle import closed_forms as cf
from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_philox_known_answers(orc):
    """Random123 kat_vectors for philox4x32-10."""
    assert orc.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert orc.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert orc.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_forward_algorithm_hand_example():
    """test/inference/particle_filter.jl:29-48."""
    prior = np.array([0.4, 0.6])
    E = np.array([[0.1, 0.9], [0.7, 0.3]]).T
    Tm = np.array([[0.5, 0.5], [0.2, 0.8]]).T
    obs = [2, 1]
    exp = 0.0
    for z1 in (1, 2):
        for z2 in (1, 2):
            exp += prior[z1 - 1] * Tm[z2 - 1, z1 - 1] * E[obs[0] - 1, z1 - 1] * E[obs[1] - 1, z2 - 1]
    assert cf.hmm_forward_alg(prior, E, Tm, obs) == pytest.approx(exp, rel=1e-14)


def test_hmm_fixture_log_ml():
    v = math.log(cf.hmm_forward_alg(cf.HMM_PRIOR, cf.HMM_EMISSION, cf.HMM_TRANSITION, cf.HMM_OBS))
    assert v == pytest.approx(-4.87645083351704, abs=1e-13)


@pytest.mark.parametrize("prop", [0, 1])
@pytest.mark.parametrize("libm", [False, True])
def test_hmm_particle_filter_reference_test(orc, orc_libm, prop, libm):
    """test/inference/particle_filter.jl:96-168: N=10^4, ess_threshold=N, atol 0.01 on ONE fixed Julia seed. The
    estimator's standard deviation is 0.0125 (default proposal) / 0.008 (locally optimal), so a single run is inside
    0.01 for ~60% / ~80% of the seeds whatever the sampler; the check that does not depend on the luck of one seed is
    made over 24 seeds: no bias (|mean error| < 3 standard errors) and the reference's bar met by most runs."""
    o = orc_libm if libm else orc
    errs = []
    for seed in range(24):
        pf = o.particle_filter(O.HMM, cf.hmm_params(), 10000, seed=seed)
        pf.init([cf.HMM_OBS[0]], proposal=prop)
        for T in range(2, 5):
            pf.maybe_resample(ess_threshold=10000)
            pf.step([cf.HMM_OBS[T - 1]], proposal=prop)
        errs.append(pf.log_ml_estimate() - cf.HMM_LOG_ML)
    errs = np.array(errs)
    sd = 0.0125 if prop == 0 else 0.0085
    assert abs(errs.mean()) < 3 * sd / math.sqrt(len(errs)), errs.mean()
    assert np.abs(errs).max() < 4 * sd
    assert np.mean(np.abs(errs) < 0.01) >= 0.45


def test_hmm_custom_proposal_weights_are_parent_marginals(orc):
    """under the locally optimal proposal the increment is log sum_z Tr*E."""
    N = 500
    pf = orc.particle_filter(O.HMM, cf.hmm_params(), N, seed=3)
    pf.init([1], proposal=1)
    expect0 = math.log(float(np.sum(cf.HMM_PRIOR * cf.HMM_EMISSION[0, :])))
    assert np.allclose(pf.log_weights(), expect0, atol=1e-14)
    z0 = pf.state()[0].astype(int)
    pf.step([2], proposal=1)
    inc = pf.log_weights() - expect0
    want = np.log(np.array([np.sum(cf.HMM_TRANSITION[:, z - 1] * cf.HMM_EMISSION[1, :]) for z in z0]))
    assert np.allclose(inc, want, atol=1e-13)


def test_unfold_extension_weight_identity(orc):
    """test/modeling_library/unfold.jl:196-234: extending by a step whose observation is constrained
    adds logpdf(obs | new latent) only; the new latent comes from the kernel's prior."""
    alpha, beta, std, x_init = 0.2, 0.3, 1.0, 0.1
    params = [x_init, 1e-300, alpha, beta, std, 1.0, 0.5]     # s0 ~ 0: x_1 = x_init exactly
    N = 64
    pf = orc.particle_filter(O.LGSSM, params, N, seed=1)
    pf.init([0.0])
    assert np.all(pf.state()[0] == x_init)
    w0 = pf.log_weights()
    z = np.random.default_rng(0).standard_normal(N)
    pf.step([1.3], z_replay=z)
    x = pf.state()[0]
    assert np.array_equal(x, (x_init * alpha + beta) + std * z)             # normal.jl:96
    inc = pf.log_weights() - w0
    want = np.array([cf.normal_logpdf(1.3, 1.0 * xi, 0.5) for xi in x])      # logpdf of the constrained choice only
    assert np.allclose(inc, want, rtol=0, atol=1e-13)


def test_normal_logpdf_formula(orc):
    """normal.jl:56-60 and support edge cases of test/modeling_library/distributions.jl:43,215."""
    L = orc.L
    assert L.orc_logpdf_normal(0.3, -0.1, 2.0) == pytest.approx(-(0.4 ** 2) / (2 * 4.0) - 0.5 * math.log(2 * math.pi * 4.0), rel=1e-15)
    assert orc.logpdf_categorical(4, [0.2, 0.3, 0.5]) == -math.inf
    assert orc.logpdf_categorical(0, [0.2, 0.3, 0.5]) == -math.inf
    assert orc.logpdf_categorical(2, [0.2, 0.3, 0.5]) == pytest.approx(math.log(0.3), rel=1e-15)
    assert L.orc_logpdf_uniform(-0.5, 0.0, 1.0) == -math.inf
    assert L.orc_logpdf_uniform(0.5, 0.0, 2.0) == pytest.approx(-math.log(2.0))
    assert L.orc_logpdf_bernoulli(1, 0.3) == pytest.approx(math.log(0.3))
    assert L.orc_logpdf_bernoulli(0, 0.3) == pytest.approx(math.log(0.7))


def test_logsumexp_and_ess(orc):
    """inference.jl:3-11; particle_filter.jl:3-12."""
    a = np.array([-1.0, -2.0, -0.5, -30.0])
    assert orc.logsumexp(a) == pytest.approx(np.log(np.sum(np.exp(a))), rel=1e-15)
    assert orc.logsumexp([-math.inf, -math.inf]) == -math.inf
    assert orc.L.orc_logsumexp2(-1.0, -2.0) == pytest.approx(np.log(np.exp(-1) + np.exp(-2)), rel=1e-15)
    assert orc.L.orc_logsumexp2(-math.inf, -math.inf) == -math.inf
    lnw = a - orc.logsumexp(a)
    p = np.exp(lnw)
    assert orc.effective_sample_size(lnw) == pytest.approx(1.0 / np.sum(p * p), rel=1e-14)
    assert orc.effective_sample_size(np.full(8, -math.log(8))) == pytest.approx(8.0, rel=1e-14)


def test_importance_sampling_invariants(orc):
    """test/inference/importance_sampling.jl:18-34 (n=4) and the closed form log N(2; 0, sqrt 2)."""
    for prop, pp in ((0, None), (1, [0.0, 2.0])):
        lat, lnw, lml = orc.importance_sampling(O.NORMAL_NORMAL, [0, 1, 1], [2.0], 4, seed=0, proposal=prop, prop_params=pp)
        assert lat.shape == (1, 4) and lnw.shape == (4,)
        assert abs(orc.logsumexp(lnw)) < 1e-14
        assert not math.isnan(lml)
    lat, lnw, lml = orc.importance_sampling(O.NORMAL_NORMAL, [0, 1, 1], [2.0], 400000, seed=1, proposal=1, prop_params=[1.0, 1.0])
    assert lml == pytest.approx(-2.2655121234846454, abs=0.01)


def test_regression_is_against_conjugate_closed_form(orc):
    exact = cf.regression_log_ml(cf.QUICKSTART_XS, cf.QUICKSTART_YS, 2, 10, 1)
    assert exact == pytest.approx(cf.QUICKSTART_LOG_ML, abs=1e-9)
    lat, lnw, lml = orc.importance_sampling(O.REGRESSION, cf.regression_params(), cf.QUICKSTART_YS, 300000, seed=0,
                                            proposal=1, prop_params=[-2.0, 0.3, 10.0, 1.5])
    assert lml == pytest.approx(exact, abs=0.03)
    assert abs(orc.logsumexp(lnw)) < 1e-10


def test_lgssm_against_kalman(orc):
    params = [0.0, 1.0, 0.9, 0.0, 1.0, 1.0, 1.0]
    ys = cf.simulate_lgssm(30, params, 0)
    exact = cf.kalman_log_ml(ys, *params)
    for prop in (0, 1):
        pf = orc.particle_filter(O.LGSSM, params, 1 << 15, seed=5)
        pf.init([ys[0]], proposal=prop)
        for t in range(1, 30):
            pf.maybe_resample()
            pf.step([ys[t]], proposal=prop)
        assert pf.log_ml_estimate() == pytest.approx(exact, abs=0.15 if prop == 0 else 0.03)


def test_oracle_math_builds_agree(orc, orc_libm):
    """The IEEE-only math build and the glibc build give the same filter up to rounding."""
    params = [0.0, 1.0, 0.9, 0.0, 1.0, 1.0, 1.0]
    ys = cf.simulate_lgssm(10, params, 1)
    outs = []
    for o in (orc, orc_libm):
        pf = o.particle_filter(O.LGSSM, params, 4096, seed=2)
        pf.init([ys[0]])
        for t in range(1, 10):
            pf.step([ys[t]])
        outs.append((pf.log_weights(), pf.log_ml_estimate()))
    assert np.allclose(outs[0][0], outs[1][0], rtol=1e-12, atol=1e-12)
    assert outs[0][1] == pytest.approx(outs[1][1], rel=1e-13)


def test_resampling_arithmetic(orc):
    """Oracle-defined integer resampling: iid search == numpy searchsorted on the integer CDF; the
    grouped-order-statistics thresholds follow their definition; residual copies are floor(N p)."""
    rng = np.random.default_rng(0)
    N = 5000
    lw = rng.standard_normal(N) * 3
    q, m = orc.quantise_weights(lw)
    assert m == lw.max() and q.max() == 1 << orc.L.orc_weight_shift(N)
    cdf = np.cumsum(q, dtype=np.uint64)
    u = rng.random(N)
    anc = orc.search_iid(cdf, u)
    T = [(int(math.floor(x * 2 ** 53)) * int(cdf[-1])) >> 53 for x in u]
    assert np.array_equal(anc, np.searchsorted(cdf, np.array(T, dtype=np.uint64), side="right"))
    # grouped order statistics, restated here from the primitives (gaps, Philox words) with Python integers
    M = N
    anc_s = orc.search_sorted(cdf, 7, 0, M)
    assert anc_s.min() >= 0 and anc_s.max() < N
    n_groups = (M + 255) // 256
    g = [int(x) for x in orc.gaps(7, 0, M, 0, n_groups)]
    assert g[-1] == orc.gap_variate(7, 0, n_groups - 1, M - 256 * (n_groups - 1)) and orc.gaps(7, 0, M, n_groups, 1)[0] == 0
    head = orc.gap_head(7, 0, M)
    assert head == orc.gap_variate(7, 0, n_groups, 1)
    stot = head + sum(g)
    total = int(cdf[-1])
    ratio = float(total) / float(stot)
    thr = lambda x: min(int(float(x) * ratio), total - 1)
    cl = [int(c) for c in cdf]
    A, exp_anc = head, []
    for j in range(n_groups):
        TL, TH = thr(A), thr(A + g[j])
        p_lo = int(np.searchsorted(cdf, np.uint64(TL), side="right"))
        p_hi = min(int(np.searchsorted(cdf, np.uint64(TH), side="right")), N - 1)
        # keys (C_p - TL) 2^32 / (TH - TL) in integer arithmetic: D normalised to its top 32 bits, one 32-bit multiplier
        D = TH - TL
        lz = 64 - D.bit_length() if D > 0 else 0
        mul = int(9223372036854774784.0 / float((D << lz) >> 32)) if D > 0 else 0
        keys = [((((cl[p] - TL) << lz) >> 32) * mul) >> 31 for p in range(p_lo, p_hi)]
        scale = orc.L.orc_bracket_scale(TL, TH)
        assert scale == (lz << 32) | mul and all(0 <= k < 2 ** 32 for k in keys)
        assert keys == sorted(keys) and all(orc.L.orc_bracket_key(cl[p] - TL, scale) == k for p, k in zip(range(p_lo, p_hi), keys))
        assert all(abs(k - (cl[p] - TL) * 2 ** 32 // D) <= 8 for p, k in zip(range(p_lo, p_hi), keys) if D > 0)
        for k in range(256 * j, min(256 * (j + 1), M)):
            if k == 256 * j:
                exp_anc.append(p_lo)
            else:
                w = orc.philox([k >> 2, 0, 0, O.STREAM_RESAMPLE], [7, 0])[k & 3]
                exp_anc.append(p_lo + sum(1 for kk in keys if kk < w))
        A += g[j]
    assert np.array_equal(anc_s, np.array(exp_anc))
    # the draw of word w sits at TL + w (TH - TL) / 2^32: its ancestor is the exact integer search up to the rounding of one key
    A = head
    for j in range(min(n_groups, 3)):
        TL, TH = thr(A), thr(A + g[j])
        for k in range(256 * j + 1, 256 * j + 40):
            w = orc.philox([k >> 2, 0, 0, O.STREAM_RESAMPLE], [7, 0])[k & 3]
            exact = min(int(np.searchsorted(cdf, np.uint64(TL + (w * (TH - TL) >> 32)), side="right")), N - 1)
            assert abs(int(anc_s[k]) - exact) <= 1
        A += g[j]
    ga = anc_s[: (M // 256) * 256].reshape(-1, 256)
    assert np.all(ga.min(axis=1) == ga[:, 0]) and np.all(ga.max(axis=1)[:-1] <= ga[1:, 0])
    # offspring counts follow the weights
    counts = np.bincount(anc_s, minlength=N)
    p = q / q.sum()
    assert abs(np.sum(counts * p) / np.sum(N * p * p) - 1) < 0.1
    # equal weights: ESS = N up to rounding, so `ess < N` is a tie the reference decides by rounding
    pf = orc.particle_filter(O.LGSSM, [0, 1, 0.9, 0, 1, 1, 1], 1000, seed=0)
    pf.init([0.0])
    pf.set_log_weights(np.zeros(1000))
    pf.maybe_resample(1000)
    assert pf.last_ess == pytest.approx(1000.0, rel=1e-13)


def test_residual_resampling_definition(orc):
    N = 4000
    pf = orc.particle_filter(O.LGSSM, [0, 1, 0.9, 0, 1, 1, 1], N, seed=0)
    pf.init([0.5])
    lw = pf.log_weights()
    assert pf.maybe_resample(N, scheme=O.RESIDUAL) is True
    anc = pf.parents()
    p = np.exp(lw - orc.logsumexp(lw))
    floor_counts = np.floor(N * p + 1e-9).astype(int)
    counts = np.bincount(anc, minlength=N)
    assert np.all(counts >= floor_counts - 1)          # -1: fixed-point truncation at exact integers
    assert counts.sum() == N
    det = int(floor_counts.sum())
    assert np.all(np.diff(anc[:det - 5]) >= 0)          # deterministic copies are laid out in index order


def test_history_is_permuted_like_the_reference(orc):
    """particle_filter.jl:202-205 copies whole traces: history must follow the ancestors."""
    N = 300
    pf = orc.particle_filter(O.LGSSM, [0, 1, 0.9, 0, 1, 1, 1], N, seed=4, keep_history=True)
    pf.init([0.2])
    h1 = pf.history(1).copy()
    pf.step([0.1])
    assert pf.maybe_resample(N) is True
    anc = pf.parents()
    assert np.array_equal(pf.history(1), h1[:, anc])


def test_golden_fixtures(orc):
    """Committed outputs of the oracle (tests/golden/make_golden.py): guards the semantics the GPU
    tests are compared against from drifting."""
    with open(os.path.join(GOLD, "oracle_golden.json")) as fh:
        gold = json.load(fh)
    from tests.golden.make_golden import run_case
    for case in gold["cases"]:
        got = run_case(orc, case["spec"])
        for key, val in case["expect"].items():
            if isinstance(val, list):
                assert list(got[key]) == val, (case["spec"], key)
            else:
                assert got[key] == pytest.approx(val, rel=1e-13, abs=1e-13), (case["spec"], key)


def test_bernoulli_uniform_families(orc):
    """The two importance-sampling families that put bernoulli.jl:10-19 and uniform_continuous.jl:12-23 on the path."""
    lo, hi, sd, y = -1.0, 3.0, 0.7, 0.4
    lat, lnw, lml = orc.importance_sampling(O.UNIFORM_NORMAL, [lo, hi, sd], [y], 200000, seed=1)
    Phi = lambda v: 0.5 * (1 + math.erf(v / math.sqrt(2)))
    assert lml == pytest.approx(math.log((Phi((hi - y) / sd) - Phi((lo - y) / sd)) / (hi - lo)), abs=0.01)
    u = orc.uniforms(1, 1, O.STREAM_UNIFORM, 0, 5)
    assert np.array_equal(lat[0, :5], u * (hi - lo) + lo)                        # random = rand() * (high - low) + low
    assert orc.L.orc_logpdf_uniform(3.5, lo, hi) == -math.inf and orc.L.orc_logpdf_uniform(3.0, lo, hi) == -math.log(4.0)
    assert orc.L.orc_logpdf_bernoulli(1, 0.3) == math.log(0.3) and orc.L.orc_logpdf_bernoulli(0, 0.3) == math.log(1. - 0.3)
    n, ns = 40, 3000
    xs = np.linspace(-5, 5, n)
    ys = -xs + 2 + 0.3 * np.sin(7 * xs)
    lat, lnw, lml = orc.importance_sampling(O.OUTLIER_REGRESSION, np.concatenate([[n, 0.5, 2.0], xs]), ys, ns, seed=2)
    assert abs(orc.logsumexp(lnw)) < 1e-12 and np.isfinite(lml)
    z = orc.normals(2, 1, 0, 8)
    assert np.array_equal(lat[:4, 0], 2.0 * z[:4]) and np.array_equal(lat[:4, 1], 2.0 * z[4:8])
    u = orc.uniforms(2, 1, O.STREAM_UNIFORM, 0, 2 * n).reshape(2, n)
    for s in range(2):
        flags = [(int(lat[4 + (i >> 5), s]) >> (i & 31)) & 1 for i in range(n)]
        assert flags == [int(v < 0.5) for v in u[s]]
        # weight = sum of the constrained :y logpdfs with std = ifelse(z, inlier_std, outlier_std) (static_model.jl:6)
        stds = np.where(np.array(flags) == 1, math.exp(lat[0, s]), math.exp(lat[1, s]))
        w = sum(orc.L.orc_logpdf_normal(float(ys[i]), float(xs[i] * lat[2, s] + lat[3, s]), float(stds[i])) for i in range(n))
        assert lnw[s] + (lml + math.log(ns)) == pytest.approx(w, rel=1e-9)
