"""Particle marginal Metropolis-Hastings on top of the device particle filter.

Mirror of /root/reference/examples/pmmh/pf.jl (`ParticleFilterCombinator`: a generative function whose score is the
particle filter's log marginal likelihood estimate, `generate` :54-62, `update` :64-75, `regenerate` :77-86) and of the
parameter moves of examples/pmmh/example.jl:62-78 (`mh(tr, select(:var_x))` = resimulation from the prior,
`mh(tr, var_x_proposal, ())` = random walk). The particle filter itself -- every `initialize_particle_filter`,
`maybe_resample!`, `particle_filter_step!` of pf.jl:40-56 -- runs in libgensmc.so; only the accept/reject of the few
scalar parameters happens here, with a host RNG (as Gen's `mh` uses Julia's).
"""
import math

import numpy as np

from .choicemap import ChoiceMap
from .inference import ParticleFilterState


class PFCombinatorTrace:
    """pf.jl:1-12: args, the generative function, the emission choices and log_ml_est as the score."""

    def __init__(self, args, gen_fn, emission_choices, log_ml_est):
        self.args, self.gen_fn, self.emission_choices, self.log_ml_est = tuple(args), gen_fn, emission_choices, float(log_ml_est)

    def get_args(self):
        return self.args

    def get_retval(self):
        return None

    def get_gen_fn(self):
        return self.gen_fn

    def get_score(self):
        return self.log_ml_est

    def get_choices(self):
        return self.emission_choices


class ParticleFilterCombinator:
    """ParticleFilterCombinator(make_model, num_particles): `make_model(*params)` returns a catalogue state-space
    model (the reference builds `Unfold(kernel)` from init/dynamics/emission, pf.jl:19-38); args = (T, *params).
    Every evaluation runs one complete filter on the device (sync-free loop, `gsmc_run_steps`) with a fresh seed."""

    def __init__(self, make_model, num_particles, seed=0, ess_threshold=None, **options):
        self.make_model, self.num_particles = make_model, int(num_particles)
        self.ess_threshold = self.num_particles / 2 if ess_threshold is None else float(ess_threshold)
        self.options = dict(options)
        self._seed, self.evaluations = int(seed), 0

    def run_particle_filter(self, args, choices):
        """pf.jl:40-56."""
        T, params = int(args[0]), args[1:]
        model = self.make_model(*params)
        ys = np.array([model.extract_observations(t, _one(model, t, choices)) for t in range(1, T + 1)], dtype=np.float64)
        opts = dict(keep_history=False)
        opts.update(self.options)
        st = ParticleFilterState(model, self.num_particles, seed=self._seed + self.evaluations, **opts)
        self.evaluations += 1
        try:
            st.init(ys[0])
            if T > 1:
                st.run_steps(ys[1:], self.ess_threshold)
            return st.log_ml_estimate()
        finally:
            st.close()

    def generate(self, args, choices):
        """pf.jl:58-62: (trace, log_ml_est)."""
        lml = self.run_particle_filter(args, choices)
        return PFCombinatorTrace(args, self, choices, lml), lml

    def update(self, trace, args, argdiff=None, choices=None):
        """pf.jl:64-75: reruns the filter with the new args; weight = new_log_ml_est - old."""
        if choices is not None and len(choices) > 0:
            raise NotImplementedError("Not implemented")           # pf.jl:65-67
        lml = self.run_particle_filter(args, trace.emission_choices)
        return PFCombinatorTrace(args, self, trace.emission_choices, lml), lml - trace.log_ml_est, None, ChoiceMap()

    def regenerate(self, trace, args, argdiff=None, selection=None):
        """pf.jl:77-86."""
        if selection:
            raise NotImplementedError("Not implemented")
        new_trace, weight, _, _ = self.update(trace, args)
        return new_trace, weight, None


def _one(model, t, choices):
    addr = model.obs_address(t)
    cm = ChoiceMap()
    cm[addr] = choices[addr]
    return cm


def pmmh(gen_fn, T, observations, log_prior, init_params, iters, step_sd, prior_sampler=None, seed=0, callback=None):
    """Metropolis-Hastings over the parameters with the particle filter's log-ML estimate as the likelihood
    (examples/pmmh/example.jl:62-78). Per iteration and parameter: a random-walk move
    `theta_k' ~ normal(theta_k, step_sd[k])` (example.jl:43-51,71-72), accepted with probability
    min(1, exp(log_prior' + lml' - log_prior - lml)); when `prior_sampler` is given, also the reference's
    resimulation move `mh(tr, select(k))` (proposal = prior, acceptance ratio = the likelihood-estimate ratio alone).
    Returns (samples[iters][n_params], log_ml_estimates[iters], acceptance_rate)."""
    rng = np.random.default_rng(seed)
    theta = np.array(init_params, dtype=np.float64)
    trace, _ = gen_fn.generate((T, *theta), observations)
    lp = log_prior(theta)
    out, lmls, acc, tried = np.empty((iters, theta.size)), np.empty(iters), 0, 0
    for it in range(iters):
        for k in range(theta.size):
            if prior_sampler is not None:
                cand = theta.copy()
                cand[k] = prior_sampler(k, rng)
                new_trace, weight, _ = gen_fn.regenerate(trace, (T, *cand))
                tried += 1
                if math.log(rng.random()) < weight:
                    theta, trace, lp = cand, new_trace, log_prior(cand)
                    acc += 1
            cand = theta.copy()
            cand[k] = theta[k] + step_sd[k] * rng.standard_normal()
            lp_c = log_prior(cand)
            tried += 1
            if np.isfinite(lp_c):
                new_trace, weight, _, _ = gen_fn.update(trace, (T, *cand))
                if math.log(rng.random()) < weight + lp_c - lp:
                    theta, trace, lp = cand, new_trace, lp_c
                    acc += 1
        out[it], lmls[it] = theta, trace.get_score()
        if callback:
            callback(it, theta, trace)
    return out, lmls, acc / max(tried, 1)
