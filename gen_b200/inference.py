"""Gen's src/inference particle-filter and importance-sampling API on top of libgensmc.so.

Names, argument order and meaning follow /root/reference/src/inference/particle_filter.jl and
importance.jl; Python cannot spell `!`, so `particle_filter_step!` is `particle_filter_step_b`
(b for bang) with `particle_filter_step_` as an alias, likewise `maybe_resample_b`.
The Julia spelling lives in julia/GenB200.jl.
"""
import ctypes as C
import math

import numpy as np

from . import _lib
from .choicemap import ChoiceMap, choicemap
from .models import DeviceModel, DeviceProposal, LinearRegression, OutlierRegression


class ParticleFilterState:
    """Device-resident ParticleFilterState{U} (particle_filter.jl:18-24): `traces` are
    structure-of-arrays columns behind an opaque handle, `log_weights`, `log_ml_est` and
    `parents` live on the device."""

    def __init__(self, model, num_particles, seed=0, dtype="f64", resample="multinomial", keep_history=True,
                 history_capacity=128, device=-1, stream=None, comm=None):
        lib = _lib.load()
        self.model = model
        self.num_particles = int(num_particles)
        cfg = _lib.Config()
        cfg.struct_size = C.sizeof(_lib.Config)
        cfg.model_id = model.family
        cfg.dtype = {"f64": _lib.F64, "f32": _lib.F32}[dtype]
        cfg.resample_scheme = {"multinomial": _lib.RESAMPLE_MULTINOMIAL, "residual": _lib.RESAMPLE_RESIDUAL}[resample]
        cfg.num_particles = self.num_particles
        cfg.seed = int(seed)
        cfg.device = int(device)
        cfg.keep_history = 1 if keep_history else 0
        cfg.history_capacity = int(history_capacity)
        cfg.stream = stream
        self._cfg = cfg
        p = np.ascontiguousarray(model.params(), dtype=np.float64)
        h = C.c_void_p()
        _lib.check(lib.gsmc_create(C.byref(cfg), _lib.dptr(p), p.size, C.byref(h)))
        self.handle = h
        self.lib = lib
        self.observations = []       # per time step, host copies (for get_traces)
        self.T = 0
        if comm is not None:
            comm.attach(self)
        n, first = C.c_uint64(), C.c_uint64()
        _lib.check(lib.gsmc_local_count(h, C.byref(n), C.byref(first)), h)
        self.num_local, self.first_global = n.value, first.value
        d = C.c_int()
        _lib.check(lib.gsmc_state_dim(h, C.byref(d)), h)
        self.D = d.value

    @classmethod
    def _adopt(cls, model, handle, num_particles):
        self = cls.__new__(cls)
        self.lib = _lib.load()
        self.model, self.handle, self.num_particles = model, handle, int(num_particles)
        self.observations, self.T = [], 1
        n, first = C.c_uint64(), C.c_uint64()
        _lib.check(self.lib.gsmc_local_count(handle, C.byref(n), C.byref(first)), handle)
        self.num_local, self.first_global = n.value, first.value
        d = C.c_int()
        _lib.check(self.lib.gsmc_state_dim(handle, C.byref(d)), handle)
        self.D = d.value
        return self

    def close(self):
        if getattr(self, "handle", None):
            self.lib.gsmc_destroy(self.handle)
            self.handle = None

    __del__ = close

    # --- raw device operations (thin wrappers over the C ABI) -----------------------------------
    def _refresh_counts(self):
        n, first = C.c_uint64(), C.c_uint64()
        _lib.check(self.lib.gsmc_local_count(self.handle, C.byref(n), C.byref(first)), self.handle)
        self.num_local, self.first_global = n.value, first.value

    def set_replay(self, normals=None, uniforms=None):
        z = None if normals is None else np.ascontiguousarray(normals, dtype=np.float64)
        u = None if uniforms is None else np.ascontiguousarray(uniforms, dtype=np.float64)
        _lib.check(self.lib.gsmc_set_replay(self.handle, _lib.dptr(z), 0 if z is None else z.size,
                                            _lib.dptr(u), 0 if u is None else u.size), self.handle)

    def _propagate(self, fn, obs, proposal):
        # per-step hot path of the host loop: short observation vectors go through a preallocated ctypes buffer
        # (no numpy array + pointer cast per call); the library copies what it needs before returning
        if obs is None:                      # unobserved step: the observation choice is sampled on the device
            if proposal is not None:
                raise _lib.GsmcError(_lib.E_BADARG, "an unobserved step takes the default proposal")
            rc = fn(self.handle, None, 0, _lib.PROPOSAL_DEFAULT, None, 0)
            if rc:
                _lib.check(rc, self.handle)
            self.T += 1
            self.observations.append(None)
            return
        n = len(obs)
        if n <= 16:
            buf = self.__dict__.get("_obs_buf")
            if buf is None:
                buf = self._obs_buf = (C.c_double * 16)()
            for i in range(n):
                buf[i] = obs[i]
            optr, kept = buf, np.array(buf[:n], dtype=np.float64)
        else:
            kept = np.ascontiguousarray(obs, dtype=np.float64).copy()
            optr = _lib.dptr(kept)
        if proposal is None:
            rc = fn(self.handle, optr, n, _lib.PROPOSAL_DEFAULT, None, 0)
        else:
            pp = np.ascontiguousarray(proposal.params, dtype=np.float64)
            rc = fn(self.handle, optr, n, proposal.proposal_id, _lib.dptr(pp), pp.size)
        if rc:
            _lib.check(rc, self.handle)
        self.T += 1
        self.observations.append(kept)

    def reset(self):
        _lib.check(self.lib.gsmc_reset(self.handle), self.handle)
        self.T = 0
        self.observations = []

    def init(self, obs, proposal=None):
        self._propagate(self.lib.gsmc_init, obs, proposal)

    def step(self, obs, proposal=None):
        self._propagate(self.lib.gsmc_step, obs, proposal)

    def run_steps(self, obs, ess_threshold, proposal=None):
        obs = np.ascontiguousarray(obs, dtype=np.float64)
        if obs.ndim == 1:
            obs = obs[:, None]
        pid, pp = _lib.PROPOSAL_DEFAULT, None
        if proposal is not None:
            pid, pp = proposal.proposal_id, np.ascontiguousarray(proposal.params, dtype=np.float64)
        _lib.check(self.lib.gsmc_run_steps(self.handle, _lib.dptr(obs), obs.shape[0], obs.shape[1], pid, _lib.dptr(pp),
                                           0 if pp is None else pp.size, float(ess_threshold)), self.handle)
        self.T += obs.shape[0]
        self.observations.extend(list(obs.copy()))

    def maybe_resample(self, ess_threshold):
        out = self.__dict__.get("_mr_out")
        if out is None:
            did, ess = C.c_int(), C.c_double()
            out = self._mr_out = (did, ess, C.byref(did), C.byref(ess))
        rc = self.lib.gsmc_maybe_resample(self.handle, float(ess_threshold), out[2], out[3])
        if rc:
            _lib.check(rc, self.handle)
        self.last_ess = out[1].value
        return bool(out[0].value)

    def log_ml_estimate(self):
        out = C.c_double()
        _lib.check(self.lib.gsmc_log_ml_estimate(self.handle, C.byref(out)), self.handle)
        return out.value

    def log_weights(self):
        out = np.empty(self.num_local, dtype=np.float64)
        _lib.check(self.lib.gsmc_get_log_weights(self.handle, _lib.dptr(out), out.size), self.handle)
        return out

    def state(self, t=0):
        """Latent columns [D, n_local] of time step t (1-based; 0 = current)."""
        out = np.empty((self.D, self.num_local), dtype=np.float64)
        _lib.check(self.lib.gsmc_get_state(self.handle, int(t), _lib.dptr(out), out.size), self.handle)
        return out

    def sampled_observation(self, t=0):
        """Observation choices the device sampled at UNOBSERVED step t (1-based; 0 = current), current particle order."""
        out = np.empty(self.num_local, dtype=np.float64)
        _lib.check(self.lib.gsmc_get_observation(self.handle, int(t), _lib.dptr(out), out.size), self.handle)
        return out

    def trajectories(self, idx, local=False):
        """Trajectories of particles given by GLOBAL index (local=True: indices into this rank's shard)."""
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        if local:
            idx = idx + self.first_global
        out = np.empty((idx.size, self.T, self.D), dtype=np.float64)
        _lib.check(self.lib.gsmc_get_trajectories(self.handle, _lib.iptr(idx), idx.size, _lib.dptr(out), out.size), self.handle)
        return out

    def ancestors(self):
        out = np.empty(self.num_local, dtype=np.int64)
        _lib.check(self.lib.gsmc_get_ancestors(self.handle, _lib.iptr(out), out.size), self.handle)
        return out

    def sample_unweighted(self, num_samples):
        out = np.empty(int(num_samples), dtype=np.int64)
        _lib.check(self.lib.gsmc_sample_unweighted(self.handle, int(num_samples), _lib.iptr(out)), self.handle)
        return out

    def save(self, path):
        """Checkpoint (gsmc_save): one file per handle / per rank."""
        _lib.check(self.lib.gsmc_save(self.handle, str(path).encode()), self.handle)

    def restore(self, path, observations=()):
        """Resume from a checkpoint written by a filter with the same configuration (gsmc_restore)."""
        _lib.check(self.lib.gsmc_restore(self.handle, str(path).encode()), self.handle)
        self.T = self.stats()["num_steps"]
        self.observations = [np.atleast_1d(np.asarray(o, dtype=np.float64)) for o in observations]

    def stats(self):
        s = _lib.Stats()
        _lib.check(self.lib.gsmc_get_stats(self.handle, C.byref(s)), self.handle)
        return s.as_dict()

    def set_profiling(self, on):
        _lib.check(self.lib.gsmc_set_profiling(self.handle, 1 if on else 0), self.handle)

    def synchronize(self):
        _lib.check(self.lib.gsmc_synchronize(self.handle), self.handle)

    def timer_start(self):
        _lib.check(self.lib.gsmc_timer_start(self.handle), self.handle)

    def timer_stop(self):
        ms = C.c_double()
        _lib.check(self.lib.gsmc_timer_stop(self.handle, C.byref(ms)), self.handle)
        return ms.value


class DeviceTrace:
    """One particle's trace, materialised lazily from the device columns: the latent trajectory
    (walking the ancestor columns) plus the observations, under the reference's addresses."""

    def __init__(self, state, index, trajectory):
        self._state, self.index, self._traj = state, index, trajectory

    def get_choices(self):
        st, m = self._state, self._state.model
        if isinstance(m, OutlierRegression):
            return m.choices(self._traj[0], st.observations[0])
        cm = ChoiceMap()
        for t in range(1, self._traj.shape[0] + 1):
            for d, name in enumerate(m.state_names):
                v = self._traj[t - 1, d]
                cm[m.latent_address(t, name)] = int(v) if m.family == _lib.MODEL_HMM else float(v)
            if t - 1 < len(st.observations):
                obs = st.observations[t - 1]
                if obs is None:              # unobserved step: the choice this particle's trace sampled
                    key = (st.T, t)
                    cache = st.__dict__.setdefault("_sampled_obs_cache", {})
                    if key not in cache:
                        cache.clear()
                        cache[key] = st.sampled_observation(t)
                    v = cache[key][self.index]
                    cm[m.obs_address(t)] = int(v) if m.family == _lib.MODEL_HMM else float(v)
                elif isinstance(m, LinearRegression):
                    for i, y in enumerate(obs):
                        cm["y-%d" % (i + 1)] = float(y)
                else:
                    cm[m.obs_address(t)] = int(obs[0]) if m.family == _lib.MODEL_HMM else float(obs[0])
        return cm

    def __getitem__(self, addr):
        return self.get_choices()[addr]

    def get_retval(self):
        return self._traj.copy()


class DeviceTraces:
    """`state.traces` (particle_filter.jl:31-34) as a lazy sequence."""

    def __init__(self, state):
        self._state = state

    def __len__(self):
        return self._state.num_local

    def __getitem__(self, i):
        if isinstance(i, slice):
            idx = np.arange(*i.indices(len(self)))
            tr = self._state.trajectories(idx, local=True)
            return [DeviceTrace(self._state, int(j), tr[k]) for k, j in enumerate(idx)]
        i = int(i)
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        return DeviceTrace(self._state, i, self._state.trajectories([i], local=True)[0])

    def __iter__(self):
        for lo in range(0, len(self), 4096):
            for tr in self[lo:min(lo + 4096, len(self))]:
                yield tr


def _check_proposal(model, proposal):
    if proposal is not None and (not isinstance(proposal, DeviceProposal) or proposal.model is not model):
        raise _lib.GsmcError(_lib.E_UNSUPPORTED, "proposal must be a DeviceProposal of the same catalogue model")


def _bind(model, model_args):
    if isinstance(model, (LinearRegression, OutlierRegression)):
        if len(model_args) != 1:
            raise _lib.GsmcError(_lib.E_BADARG, "model_args must be (xs,)")
        return model.bind(model_args[0])
    return model


# -------------------------------------------------------------------------------------------------
# particle_filter.jl
# -------------------------------------------------------------------------------------------------
def initialize_particle_filter(model, model_args, observations, *rest, **options):
    """initialize_particle_filter(model, model_args, observations, num_particles)
    initialize_particle_filter(model, model_args, observations, proposal, proposal_args, num_particles)
    (particle_filter.jl:79-108). model_args is (1,) for the state-space families. Keyword options
    (seed, dtype, resample, keep_history, history_capacity, device, stream, comm) configure the device state."""
    if len(rest) == 1:
        proposal, proposal_args, num_particles = None, (), rest[0]
    elif len(rest) == 3:
        proposal, proposal_args, num_particles = rest
    else:
        raise TypeError("initialize_particle_filter(model, model_args, observations[, proposal, proposal_args], num_particles)")
    if not isinstance(model, DeviceModel):
        raise _lib.GsmcError(_lib.E_UNSUPPORTED, "model must be a catalogue DeviceModel")
    _check_proposal(model, proposal)
    if tuple(model_args) != (1,):
        raise _lib.GsmcError(_lib.E_BADARG, "the filter starts with one time step: model_args must be (1,)")
    state = ParticleFilterState(model, num_particles, **options)
    state.init(model.extract_observations(1, observations), proposal)
    return state


def particle_filter_step_b(state, new_args, argdiffs, observations, proposal=None, proposal_args=()):
    """particle_filter_step!(state, new_args, argdiffs, observations[, proposal, proposal_args])
    (particle_filter.jl:139-180). new_args must be (T+1,): the traces are extended by one step."""
    _check_proposal(state.model, proposal)
    if tuple(new_args) != (state.T + 1,):
        raise _lib.GsmcError(_lib.E_BADARG, "new_args must be (%d,): a step extends the traces by one time step" % (state.T + 1))
    state.step(state.model.extract_observations(state.T + 1, observations), proposal)
    return None


def maybe_resample_b(state, ess_threshold=None, verbose=False):
    """maybe_resample!(state; ess_threshold=N/2, verbose=false) (particle_filter.jl:189-213)."""
    if ess_threshold is None:
        ess_threshold = state.num_particles / 2
    did = state.maybe_resample(ess_threshold)
    if verbose:
        print("effective sample size: %s, doing resample: %s" % (state.last_ess, "true" if did else "false"))
    return did


particle_filter_step_ = particle_filter_step_b
maybe_resample_ = maybe_resample_b


def log_ml_estimate(state):
    """particle_filter.jl:52-55."""
    return state.log_ml_estimate()


def get_log_weights(state):
    """particle_filter.jl:43-45 (unnormalised, log space)."""
    return state.log_weights()


def get_traces(state):
    """particle_filter.jl:31-34."""
    return DeviceTraces(state)


def sample_unweighted_traces(state, num_samples):
    """particle_filter.jl:62-70."""
    idx = state.sample_unweighted(num_samples)
    tr = state.trajectories(idx)
    return [DeviceTrace(state, int(j), tr[k]) for k, j in enumerate(idx)]


# -------------------------------------------------------------------------------------------------
# importance.jl
# -------------------------------------------------------------------------------------------------
def importance_sampling(model, model_args, observations, *rest, **options):
    """(traces, log_norm_weights, lml_est) = importance_sampling(model, model_args, observations, num_samples, verbose=false)
    (traces, log_norm_weights, lml_est) = importance_sampling(model, model_args, observations, proposal, proposal_args,
                                                              num_samples, verbose=false)          (importance.jl:20-52)"""
    if len(rest) >= 3 and isinstance(rest[0], DeviceProposal):
        proposal, proposal_args, num_samples = rest[0], rest[1], rest[2]
        verbose = rest[3] if len(rest) > 3 else False
    elif len(rest) >= 1:
        proposal, proposal_args, num_samples = None, (), rest[0]
        verbose = rest[1] if len(rest) > 1 else False
    else:
        raise TypeError("importance_sampling(model, model_args, observations[, proposal, proposal_args], num_samples, verbose=False)")
    if not isinstance(model, DeviceModel):
        raise _lib.GsmcError(_lib.E_UNSUPPORTED, "model must be a catalogue DeviceModel")
    _check_proposal(model, proposal)
    lib = _lib.load()
    bound = _bind(model, model_args)
    if proposal is not None:
        proposal = DeviceProposal(bound, proposal.name, proposal.params)
    if bound.family in _lib.IS_FAMILIES:
        cfg = _lib.Config()
        cfg.struct_size = C.sizeof(_lib.Config)
        cfg.model_id = bound.family
        cfg.dtype = {"f64": _lib.F64, "f32": _lib.F32}[options.get("dtype", "f64")]
        cfg.num_particles = int(num_samples)
        cfg.seed = int(options.get("seed", 0))
        cfg.device = int(options.get("device", -1))
        cfg.keep_history = 0
        cfg.stream = options.get("stream")
        p = np.ascontiguousarray(bound.params(), dtype=np.float64)
        obs = np.ascontiguousarray(bound.extract_observations(1, observations), dtype=np.float64)
        pid, pp = _lib.PROPOSAL_DEFAULT, None
        if proposal is not None:
            pid, pp = proposal.proposal_id, np.ascontiguousarray(proposal.params, dtype=np.float64)
        lml, h = C.c_double(), C.c_void_p()
        _lib.check(lib.gsmc_importance_sampling(C.byref(cfg), _lib.dptr(p), p.size, _lib.dptr(obs), obs.size, pid,
                                                _lib.dptr(pp), 0 if pp is None else pp.size, C.byref(lml), C.byref(h)))
        state = ParticleFilterState._adopt(bound, h, num_samples)
        state.observations = [obs.copy()]
        if verbose:
            print("sampled %d traces" % num_samples)
        return DeviceTraces(state), state.log_weights(), lml.value
    # state-space families: generate(model, (T,), observations) = init + T-1 extensions, no resampling
    (T,) = model_args
    state = ParticleFilterState(bound, num_samples, **options)
    seen = set()
    for t in range(1, T + 1):
        addr = bound.obs_address(t)
        if addr not in observations:
            raise _lib.GsmcError(_lib.E_BADARG, "observations must constrain %r" % (addr,))
        seen.add(addr)
        one = choicemap((addr, observations[addr]))
        if t == 1:
            state.init(bound.extract_observations(1, one), proposal)
        else:
            state.step(bound.extract_observations(t, one), proposal)
    extra = [k for k in observations.keys() if k not in seen]
    if extra:
        raise _lib.GsmcError(_lib.E_BADARG, "constraints at addresses the model does not visit: %r" % (extra,))
    lml = state.log_ml_estimate()                        # log_total - log(n): nothing was folded into log_ml_est
    lw = state.log_weights()
    log_total = lml + float(np.log(num_samples))
    return DeviceTraces(state), lw - log_total, lml


CHUNK_EVENT = 0xFFFFFFFF            # Philox event index of the chunk-merge draws of importance_resampling


def chunk_seed(seed, c):
    """Seed of chunk c of a chunked importance-sampling run (chunk 0 uses the seed itself)."""
    return (int(seed) + c * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF


def importance_resampling(model, model_args, observations, *rest, verbose=False, **options):
    """(trace, lml_est) = importance_resampling(model, model_args, observations, num_samples; verbose=false)
    (trace, lml_est) = importance_resampling(model, model_args, observations, proposal, proposal_args,
                                             num_samples; verbose=false)                 (importance.jl:70-108)
    `verbose` is keyword-only, as in the reference (importance.jl:72,89).

    Sampling importance resampling that returns ONE trace. The reference streams the samples one by one and keeps a
    reservoir of size one (`bernoulli(exp(log_weight - log_total_weight))`), so its memory does not grow with
    num_samples. Here the samples are generated `chunk_size` (option, default 2^24) at a time on the device; inside a
    chunk the kept trace is one categorical draw from the chunk's weights (the distribution the reference's reservoir
    has after the chunk), and chunks are merged with the reference's rule: the chunk's pick replaces the kept trace
    with probability exp(chunk_log_total - log_total_so_far). Device memory is bounded by the chunk size."""
    from . import philox
    if len(rest) == 3 and isinstance(rest[0], DeviceProposal):
        head, num_samples = (rest[0], rest[1]), rest[2]
    elif len(rest) == 1:
        head, num_samples = (), rest[0]
    else:
        raise TypeError("importance_resampling(model, model_args, observations[, proposal, proposal_args], num_samples; verbose=False)")
    num_samples = int(num_samples)
    if num_samples < 1:
        raise _lib.GsmcError(_lib.E_BADARG, "num_samples must be >= 1")
    options = dict(options)
    chunk = int(options.pop("chunk_size", 1 << 24))
    seed = int(options.pop("seed", 0))
    options.pop("keep_history", None)
    log_total, kept, done, c = -math.inf, None, 0, 0
    while done < num_samples:
        m = min(chunk, num_samples - done)
        # state-space families: the kept trace is a whole trajectory, so the chunk keeps its history
        extra = {} if model.family in _lib.IS_FAMILIES else {"keep_history": True, "history_capacity": int(model_args[0])}
        traces, _, lml_c = importance_sampling(model, model_args, observations, *head, m, seed=chunk_seed(seed, c), **extra, **options)
        state = traces._state
        lt_c = lml_c + math.log(m)
        cand = traces[int(state.sample_unweighted(1)[0])]
        new_total = lt_c if kept is None else float(np.logaddexp(log_total, lt_c))          # inference.jl:8-11
        if kept is None or philox.uniform(seed, 2 * c, CHUNK_EVENT, philox.STREAM_SAMPLE) < math.exp(lt_c - new_total):
            kept = cand
        log_total = new_total
        state.close()
        done += m
        c += 1
        if verbose:
            print("sample: %d of %d" % (done, num_samples))
    return kept, log_total - math.log(num_samples)
