"""Closed-form answers used as external truth for the oracle. TEST INFRASTRUCTURE ONLY.

hmm_forward_alg restates the checker the reference's own test uses
(/root/reference/test/inference/particle_filter.jl:1-27); the others are textbook
results that the reference does not contain.
"""
import math

import numpy as np


def hmm_forward_alg(prior, emission_dists, transition_dists, emissions):
    """test/inference/particle_filter.jl:1-27. Julia layout: emission_dists[x, z] = P(x|z),
    transition_dists[z, z_prev] = P(z|z_prev); emissions are 1-based."""
    prior = np.asarray(prior, dtype=np.float64)
    E = np.asarray(emission_dists, dtype=np.float64)
    Tm = np.asarray(transition_dists, dtype=np.float64)
    marg_lik = 1.0
    alpha = prior
    for i in range(1, len(emissions)):
        prev_posterior = alpha * E[emissions[i - 1] - 1, :]
        denom = prev_posterior.sum()
        prev_posterior = prev_posterior / denom
        alpha = Tm @ prev_posterior
        marg_lik *= denom
    prev_posterior = alpha * E[emissions[-1] - 1, :]
    marg_lik *= prev_posterior.sum()
    return marg_lik


# the fixture of test/inference/particle_filter.jl:52-81 (Julia writes the matrices transposed)
HMM_PRIOR = np.array([0.2, 0.3, 0.5])
HMM_EMISSION = np.array([[0.1, 0.2, 0.7], [0.2, 0.7, 0.1], [0.7, 0.2, 0.1]]).T      # [x, z]
HMM_TRANSITION = np.array([[0.4, 0.4, 0.2], [0.2, 0.3, 0.5], [0.9, 0.05, 0.05]]).T  # [z, z_prev]
HMM_OBS = [1, 1, 2, 3]
HMM_LOG_ML = -4.87645083351704


def hmm_params(prior=HMM_PRIOR, emission=HMM_EMISSION, transition=HMM_TRANSITION):
    """Pack into the catalogue layout [K, V, prior, trans[z_prev][z], emis[z][x]]."""
    K, V = len(prior), emission.shape[0]
    return np.concatenate([[K, V], prior, np.asarray(transition).T.reshape(-1), np.asarray(emission).T.reshape(-1)])


def normal_logpdf(x, mu, std):
    return -((x - mu) ** 2) / (2.0 * std * std) - 0.5 * math.log(2.0 * math.pi * std * std)


def kalman_log_ml(ys, m0, s0, a, b, q, c, r):
    """log p(y_1:T) of x1~N(m0,s0), x_t~N(a x+b,q), y_t~N(c x_t,r) (all std)."""
    m, P = m0, s0 * s0
    ll = 0.0
    for t, y in enumerate(ys):
        if t > 0:
            m, P = a * m + b, a * a * P + q * q
        S = c * c * P + r * r
        ll += normal_logpdf(y, c * m, math.sqrt(S))
        K = c * P / S
        m, P = m + K * (y - c * m), (1.0 - K * c) * P
    return ll


def regression_log_ml(xs, ys, sd_slope, sd_intercept, sd_noise):
    """y ~ N(0, X diag(sd_slope^2, sd_intercept^2) X' + sd_noise^2 I), X = [xs 1]."""
    xs, ys = np.asarray(xs, float), np.asarray(ys, float)
    X = np.stack([xs, np.ones_like(xs)], axis=1)
    S = X @ np.diag([sd_slope ** 2, sd_intercept ** 2]) @ X.T + sd_noise ** 2 * np.eye(len(xs))
    sign, logdet = np.linalg.slogdet(S)
    return float(-0.5 * (ys @ np.linalg.solve(S, ys)) - 0.5 * logdet - 0.5 * len(xs) * math.log(2 * math.pi))


# examples/regression/quickstart.jl:26-27
QUICKSTART_XS = [1., 2., 3., 4., 5., 6., 7., 8., 9., 10.]
QUICKSTART_YS = [8.23, 5.87, 3.99, 2.59, 0.23, -0.66, -3.53, -6.91, -7.24, -9.90]
QUICKSTART_LOG_ML = -18.150487182903948


def regression_params(xs=QUICKSTART_XS, sd_slope=2.0, sd_intercept=10.0, sd_noise=1.0):
    return np.concatenate([[len(xs), sd_slope, sd_intercept, sd_noise], xs])


def simulate_lgssm(T, params, seed):
    """Synthetic observations of the LG-SSM, drawn with numpy (only used to make inputs)."""
    m0, s0, a, b, q, c, r = params
    rng = np.random.default_rng(seed)
    x = m0 + s0 * rng.standard_normal()
    ys = []
    for t in range(T):
        if t > 0:
            x = a * x + b + q * rng.standard_normal()
        ys.append(c * x + r * rng.standard_normal())
    return np.array(ys)


def simulate_sv(T, params, seed):
    mu, phi, sigma = params
    rng = np.random.default_rng(seed)
    h = mu + sigma / math.sqrt(1 - phi * phi) * rng.standard_normal()
    ys = []
    for t in range(T):
        if t > 0:
            h = mu + phi * (h - mu) + sigma * rng.standard_normal()
        ys.append(math.exp(h / 2) * rng.standard_normal())
    return np.array(ys)


BEARINGS_PARAMS = np.array([0.0, 0.0, 12.4, -0.05, 0.5, 0.005, 0.3, 0.01, 0.001, 0.005])


def simulate_bearings(T, params=BEARINGS_PARAMS, seed=0, truth=(-0.05, 0.001, 12.0, -0.055)):
    sw, st = params[8], params[9]
    rng = np.random.default_rng(seed)
    x, vx, y, vy = truth
    obs = []
    for t in range(T):
        if t > 0:
            wx, wy = sw * rng.standard_normal(2)
            x, vx, y, vy = x + vx + 0.5 * wx, vx + wx, y + vy + 0.5 * wy, vy + wy
        obs.append(math.atan2(y, x) + st * rng.standard_normal())
    return np.array(obs)
