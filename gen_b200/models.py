"""Model catalogue: the host-side descriptors of the static-IR + Unfold models the device code
implements (gen_b200/csrc/models.cuh). Each class documents the Gen program it stands for and the
addresses of its choices, which are the reference's own conventions
(/root/reference/test/inference/particle_filter.jl:66-94: `:z_init`, `:x_init`,
`:chain => t => :z`, `:chain => t => :x`)."""
import numpy as np

from . import _lib


class DeviceProposal:
    """A catalogue proposal (stands for the `proposal::GenerativeFunction` argument of
    initialize_particle_filter / particle_filter_step! / importance_sampling)."""

    def __init__(self, model, name, params=()):
        self.model, self.name = model, name
        self.params = np.asarray(params, dtype=np.float64)
        self.proposal_id = _lib.PROPOSAL_CUSTOM

    def __repr__(self):
        return "DeviceProposal(%s, %s)" % (type(self.model).__name__, self.name)


class DeviceModel:
    """Base of catalogue models (a `GenerativeFunction` subtype in julia/GenB200.jl)."""
    family = 0
    state_names = ()          # names of the latent columns
    init_suffix = "_init"
    obs_name = "y"

    def params(self):
        raise NotImplementedError

    # --- addresses ---------------------------------------------------------------------------
    def obs_address(self, T):
        """Address of the observation that extends the trace to T time steps."""
        return self.obs_name + self.init_suffix if T == 1 else ("chain", T - 1, self.obs_name)

    def latent_address(self, T, name):
        return name + self.init_suffix if T == 1 else ("chain", T - 1, name)

    def extract_observations(self, T, observations):
        """Observation vector of time step T out of a ChoiceMap; like the reference, constraints at
        addresses the model does not visit are an error (src/dynamic/update.jl:191-193)."""
        addr = self.obs_address(T)
        extra = [k for k in observations.keys() if k != addr]
        if extra:
            raise _lib.GsmcError(_lib.E_BADARG, "constraints at addresses the model does not visit at this step: %r" % (extra,))
        if addr not in observations:
            # an empty choice map: nothing is constrained at this step, the reference samples the observation choice
            # (static_ir/generate.jl:36-42) and the weight does not change
            return None
        return np.array([float(observations[addr])], dtype=np.float64)

    def custom_proposal(self, *params):
        return DeviceProposal(self, "custom", params)


class HMM(DeviceModel):
    """Discrete HMM of test/inference/particle_filter.jl:52-78.

        @gen function kernel(t::Int, prev_z::Int, params::Nothing)
            z = @trace(categorical(transition_dists[:,prev_z]), :z)
            @trace(categorical(emission_dists[:,z]), :x)
            return z
        end
        chain = Unfold(kernel)
        @gen function model(num_steps::Int)
            z_init = @trace(categorical(prior), :z_init)
            @trace(categorical(emission_dists[:,z_init]), :x_init)
            @trace(chain(num_steps-1, z_init, nothing), :chain)
        end

    Julia layout: emission_dists[x, z] = P(x|z), transition_dists[z, z_prev] = P(z|z_prev).
    The custom proposal is the locally optimal one of :104-127.
    """
    family = _lib.MODEL_HMM
    state_names = ("z",)
    obs_name = "x"

    def __init__(self, prior, emission_dists, transition_dists):
        self.prior = np.asarray(prior, dtype=np.float64)
        self.emission = np.asarray(emission_dists, dtype=np.float64)
        self.transition = np.asarray(transition_dists, dtype=np.float64)
        K = self.prior.size
        if self.transition.shape != (K, K) or self.emission.shape[1] != K:
            raise ValueError("shapes: prior[K], emission_dists[V,K], transition_dists[K,K]")

    def params(self):
        K, V = self.prior.size, self.emission.shape[0]
        return np.concatenate([[K, V], self.prior, self.transition.T.reshape(-1), self.emission.T.reshape(-1)])


class LinearGaussianSSM(DeviceModel):
    """1-D linear-Gaussian state-space model; static kernel of test/modeling_library/unfold.jl:5-8
    with an observation choice.

        @gen (static) function kernel(t::Int, x_prev::Float64, a, b, q, c, r)
            x = @trace(normal(x_prev * a + b, q), :x)
            @trace(normal(c * x, r), :y)
            return x
        end
        @gen (static) function model(T::Int)
            x_init = @trace(normal(m0, s0), :x_init)
            @trace(normal(c * x_init, r), :y_init)
            @trace(Unfold(kernel)(T-1, x_init, a, b, q, c, r), :chain)
        end

    custom proposal: the locally optimal Gaussian q(x_t | x_{t-1}, y_t).
    """
    family = _lib.MODEL_LGSSM
    state_names = ("x",)

    def __init__(self, m0=0.0, s0=1.0, a=0.9, b=0.0, q=1.0, c=1.0, r=1.0):
        self.p = np.array([m0, s0, a, b, q, c, r], dtype=np.float64)

    def params(self):
        return self.p


class StochasticVolatility(DeviceModel):
    """h_init ~ normal(mu, sigma/sqrt(1-phi^2)); h ~ normal(mu + phi*(h_prev-mu), sigma);
    y ~ normal(0, exp(h/2)). Static kernel + Unfold as in examples/pmmh/model.jl:40-50."""
    family = _lib.MODEL_SV
    state_names = ("h",)

    def __init__(self, mu=-1.0, phi=0.97, sigma=0.2):
        self.p = np.array([mu, phi, sigma], dtype=np.float64)

    def params(self):
        return self.p


class BearingsOnly(DeviceModel):
    """2-D bearings-only tracking, state (x, vx, y, vy), constant-velocity dynamics driven by
    wx, wy ~ normal(0, sigma_w); bearing ~ normal(atan(y, x), sigma_theta). Custom proposal:
    independent Gaussians on (wx, wy) from a one-step EKF update."""
    family = _lib.MODEL_BEARINGS
    state_names = ("x", "vx", "y", "vy")
    obs_name = "bearing"

    def __init__(self, prior_mean=(0.0, 0.0, 12.4, -0.05), prior_std=(0.5, 0.005, 0.3, 0.01), sigma_w=0.001, sigma_theta=0.005):
        self.p = np.concatenate([prior_mean, prior_std, [sigma_w, sigma_theta]]).astype(np.float64)

    def params(self):
        return self.p


class LinearRegression(DeviceModel):
    """examples/regression/quickstart.jl:3-9 (importance sampling):

        @gen function my_model(xs)
            slope = @trace(normal(0, 2), :slope)
            intercept = @trace(normal(0, 10), :intercept)
            for (i, x) in enumerate(xs)
                @trace(normal(slope * x + intercept, 1), "y-$i")
            end
        end
    custom proposal: slope ~ normal(mu_s, sd_s), intercept ~ normal(mu_i, sd_i).
    """
    family = _lib.MODEL_REGRESSION
    state_names = ("slope", "intercept")

    def __init__(self, sd_slope=2.0, sd_intercept=10.0, sd_noise=1.0):
        self.sd = (float(sd_slope), float(sd_intercept), float(sd_noise))
        self.xs = None

    def bind(self, xs):
        m = LinearRegression(*self.sd)
        m.xs = np.asarray(xs, dtype=np.float64)
        return m

    def params(self):
        return np.concatenate([[self.xs.size], self.sd, self.xs])

    def latent_address(self, T, name):
        return name

    def extract_observations(self, T, observations):
        n = self.xs.size
        want = ["y-%d" % (i + 1) for i in range(n)]
        extra = [k for k in observations.keys() if k not in want]
        if extra:
            raise _lib.GsmcError(_lib.E_BADARG, "constraints at addresses the model does not visit: %r" % (extra,))
        missing = [k for k in want if k not in observations]
        if missing:
            raise _lib.GsmcError(_lib.E_BADARG, "observations must constrain %r" % (missing,))
        return np.array([float(observations[k]) for k in want], dtype=np.float64)


class NormalNormal(DeviceModel):
    """test/inference/importance_sampling.jl:3-12: x ~ normal(mu0, sd0); y ~ normal(x, sd_y);
    custom proposal x ~ normal(mu_q, sd_q)."""
    family = _lib.MODEL_NORMAL_NORMAL
    state_names = ("x",)

    def __init__(self, mu0=0.0, sd0=1.0, sd_y=1.0):
        self.p = np.array([mu0, sd0, sd_y], dtype=np.float64)

    def params(self):
        return self.p

    def obs_address(self, T):
        return "y"

    def latent_address(self, T, name):
        return name


class OutlierRegression(DeviceModel):
    """examples/regression/static_model.jl:3-23 (importance sampling; prior as proposal):

        @gen (static) function datum(x, inlier_std, outlier_std, slope, intercept)
            is_outlier = @trace(bernoulli(0.5), :z)
            std = ifelse(is_outlier, inlier_std, outlier_std)
            y = @trace(normal(x * slope + intercept, std), :y)
        end
        data = Map(datum)
        @gen (static) function model(xs)
            inlier_log_std = @trace(normal(0, 2), :log_inlier_std); outlier_log_std = @trace(normal(0, 2), :log_outlier_std)
            slope = @trace(normal(0, 2), :slope); intercept = @trace(normal(0, 2), :intercept)
            @trace(data(xs, fill(exp(inlier_log_std), n), ...), :data)
        end

    Observations constrain (:data, i, :y) for i = 1..n (n <= 256); the flags (:data, i, :z) are latent, stored bit-packed."""
    family = _lib.MODEL_OUTLIER_REGRESSION
    ZWORDS = 8
    state_names = ("log_inlier_std", "log_outlier_std", "slope", "intercept") + tuple("z_word_%d" % k for k in range(8))

    def __init__(self, prob_outlier=0.5, prior_sd=2.0):
        self.prob, self.sd = float(prob_outlier), float(prior_sd)
        self.xs = None

    def bind(self, xs):
        m = OutlierRegression(self.prob, self.sd)
        m.xs = np.asarray(xs, dtype=np.float64)
        return m

    def params(self):
        return np.concatenate([[self.xs.size, self.prob, self.sd], self.xs])

    def latent_address(self, T, name):
        return name

    def extract_observations(self, T, observations):
        n = self.xs.size
        want = [("data", i + 1, "y") for i in range(n)]
        extra = [k for k in observations.keys() if k not in want]
        if extra:
            raise _lib.GsmcError(_lib.E_BADARG, "constraints at addresses the model does not visit (or latent flags): %r" % (extra[:4],))
        missing = [k for k in want if k not in observations]
        if missing:
            raise _lib.GsmcError(_lib.E_BADARG, "observations must constrain %r" % (missing[:4],))
        return np.array([float(observations[k]) for k in want], dtype=np.float64)

    def choices(self, latent_row, obs):
        """Choice map of one trace from its latent row [4 reals + packed flag words] and the observations."""
        from .choicemap import ChoiceMap
        cm = ChoiceMap()
        for d, name in enumerate(self.state_names[:4]):
            cm[name] = float(latent_row[d])
        for i in range(self.xs.size):
            cm[("data", i + 1, "z")] = bool((int(latent_row[4 + (i >> 5)]) >> (i & 31)) & 1)
            cm[("data", i + 1, "y")] = float(obs[i])
        return cm


class UniformNormal(DeviceModel):
    """x ~ uniform(low, high); y ~ normal(x, sd_y) (uniform_continuous.jl:12-23 on the device path);
    custom proposal x ~ uniform(low_q, high_q)."""
    family = _lib.MODEL_UNIFORM_NORMAL
    state_names = ("x",)

    def __init__(self, low=0.0, high=1.0, sd_y=1.0):
        self.p = np.array([low, high, sd_y], dtype=np.float64)

    def params(self):
        return self.p

    def obs_address(self, T):
        return "y"

    def latent_address(self, T, name):
        return name
