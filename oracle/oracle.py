"""ctypes front end of the CPU oracle (oracle/gsmc_oracle.c). TEST INFRASTRUCTURE ONLY.

Nothing under gen_b200/ may import this module; it is used by tests/, by
`__graft_entry__.smoke()` and by bench.py's cpu_baseline / --impl reference legs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

HMM, LGSSM, SV, BEARINGS, REGRESSION, NORMAL_NORMAL, OUTLIER_REGRESSION, UNIFORM_NORMAL = 1, 2, 3, 4, 5, 6, 7, 8
PROPOSAL_DEFAULT, PROPOSAL_CUSTOM = 0, 1
MULTINOMIAL, RESIDUAL = 0, 1
STREAM_NORMAL, STREAM_UNIFORM, STREAM_RESAMPLE, STREAM_SAMPLE, STREAM_GAP, STREAM_OBS = 0, 1, 2, 3, 4, 5

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)
_up = C.POINTER(C.c_uint64)


def build(force=False):
    """Compile oracle/_build/liboracle{,_libm}.so with the committed Makefile."""
    out = os.path.join(_HERE, "_build", "liboracle.so")
    srcs = [os.path.join(_HERE, "gsmc_oracle.c"), os.path.join(_HERE, "gsmc_oracle.h"),
            os.path.join(_HERE, "..", "gen_b200", "csrc", "gsmc_math.h"),
            os.path.join(_HERE, "..", "gen_b200", "csrc", "gsmc_tables.h")]
    if force or not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE, "clean", "all"])
    return out


def _d(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a, t=_dp):
    return None if a is None else a.ctypes.data_as(t)


class Oracle:
    def __init__(self, libm=False):
        build()
        name = "liboracle_libm.so" if libm else "liboracle.so"
        L = self.L = C.CDLL(os.path.join(_HERE, "_build", name))
        L.orc_last_error.restype = C.c_char_p
        L.orc_exp.restype = L.orc_log.restype = C.c_double
        L.orc_exp.argtypes = L.orc_log.argtypes = [C.c_double]
        L.orc_div_inv.restype = L.orc_log_pos.restype = C.c_double
        L.orc_div_inv.argtypes = [C.c_double, C.c_double]
        L.orc_log_pos.argtypes = [C.c_double]
        L.orc_exp_nonpos.restype = C.c_double
        L.orc_exp_nonpos.argtypes = [C.c_double]
        L.orc_nlog_u32f.restype = C.c_float
        L.orc_nlog_u32f.argtypes = [C.c_uint32]
        L.orc_sincos_u32f.restype = None
        L.orc_sincos_u32f.argtypes = [C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.orc_log_unit.restype = C.c_double
        L.orc_log_unit.argtypes = [C.c_double]
        L.orc_muldiv_mismatches.restype = C.c_int64
        L.orc_muldiv_mismatches.argtypes = [C.c_uint64, C.c_int64]
        L.orc_div_inv_mismatches.restype = C.c_int64
        L.orc_div_inv_mismatches.argtypes = [C.c_uint64, C.c_int64]
        L.orc_atan2.restype = C.c_double
        L.orc_atan2.argtypes = [C.c_double, C.c_double]
        L.orc_sincospi.argtypes = [C.c_double, _dp, _dp]
        L.orc_logpdf_normal.restype = C.c_double
        L.orc_logpdf_normal.argtypes = [C.c_double] * 3
        L.orc_logpdf_categorical.restype = C.c_double
        L.orc_logpdf_categorical.argtypes = [C.c_int64, _dp, C.c_int64]
        L.orc_logpdf_uniform.restype = C.c_double
        L.orc_logpdf_uniform.argtypes = [C.c_double] * 3
        L.orc_logpdf_bernoulli.restype = C.c_double
        L.orc_logpdf_bernoulli.argtypes = [C.c_int, C.c_double]
        L.orc_logsumexp.restype = C.c_double
        L.orc_logsumexp.argtypes = [_dp, C.c_int64]
        L.orc_logsumexp2.restype = C.c_double
        L.orc_logsumexp2.argtypes = [C.c_double, C.c_double]
        L.orc_effective_sample_size.restype = C.c_double
        L.orc_effective_sample_size.argtypes = [_dp, C.c_int64]
        L.orc_philox4x32_10.argtypes = [C.POINTER(C.c_uint32)] * 3
        L.orc_fill_normals.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, _dp]
        L.orc_fill_uniforms.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, _dp]
        L.orc_fill_gaps.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, _up]
        L.orc_gap_head.restype = L.orc_gap_variate.restype = C.c_uint64
        L.orc_gap_head.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64]
        L.orc_gap_variate.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32]
        L.orc_weight_shift.restype = C.c_int
        L.orc_weight_shift.argtypes = [C.c_uint64]
        L.orc_quantise_weights.argtypes = [_dp, C.c_int64, C.c_uint64, _up, _dp]
        L.orc_search_iid.argtypes = [_up, C.c_int64, _dp, C.c_int64, _ip]
        L.orc_bracket_key.restype = C.c_uint32
        L.orc_bracket_key.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_bracket_scale.restype = C.c_uint64
        L.orc_bracket_scale.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_search_sorted.argtypes = [_up, C.c_int64, C.c_uint64, C.c_uint32, C.c_int64, _ip]
        L.orc_pf_create.restype = C.c_void_p
        L.orc_pf_create.argtypes = [C.c_int, _dp, C.c_int, C.c_int64, C.c_uint64, C.c_int, C.c_int]
        L.orc_pf_destroy.argtypes = [C.c_void_p]
        L.orc_pf_state_dim.argtypes = [C.c_void_p]
        L.orc_pf_num_normals.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_pf_num_uniforms.argtypes = [C.c_void_p, C.c_int, C.c_int]
        for f in (L.orc_pf_init, L.orc_pf_step):
            f.argtypes = [C.c_void_p, _dp, C.c_int, C.c_int, _dp, C.c_int, _dp, _dp]
        L.orc_pf_maybe_resample.argtypes = [C.c_void_p, C.c_double, C.c_int, _dp, C.POINTER(C.c_int), _dp, _dp]
        L.orc_pf_log_ml_estimate.restype = C.c_double
        L.orc_pf_log_ml_estimate.argtypes = [C.c_void_p]
        L.orc_pf_log_weights.restype = _dp
        L.orc_pf_log_weights.argtypes = [C.c_void_p]
        L.orc_pf_set_log_weights.argtypes = [C.c_void_p, _dp]
        L.orc_pf_parents.restype = _ip
        L.orc_pf_parents.argtypes = [C.c_void_p]
        L.orc_pf_state.restype = _dp
        L.orc_pf_state.argtypes = [C.c_void_p]
        L.orc_pf_history.argtypes = [C.c_void_p, C.c_int64, _dp]
        L.orc_pf_sampled_observation.argtypes = [C.c_void_p, C.c_int64, _dp]
        L.orc_pf_num_steps.restype = C.c_int64
        L.orc_pf_num_steps.argtypes = [C.c_void_p]
        L.orc_pf_sample_unweighted.argtypes = [C.c_void_p, C.c_int64, _dp, _ip]
        L.orc_importance_sampling.argtypes = [C.c_int, _dp, C.c_int, _dp, C.c_int, C.c_int, _dp, C.c_int,
                                              C.c_int64, C.c_uint64, _dp, _dp, _dp, _dp, C.c_int]

    def err(self):
        return self.L.orc_last_error().decode()

    # ---- primitives -------------------------------------------------------
    def philox(self, ctr, key):
        c = (C.c_uint32 * 4)(*ctr)
        k = (C.c_uint32 * 2)(*key)
        o = (C.c_uint32 * 4)()
        self.L.orc_philox4x32_10(c, k, o)
        return list(o)

    def normals(self, seed, t, first, count):
        out = np.empty(count, dtype=np.float64)
        self.L.orc_fill_normals(seed, t, first, count, _ptr(out))
        return out

    def uniforms(self, seed, t, stream, first, count):
        out = np.empty(count, dtype=np.float64)
        self.L.orc_fill_uniforms(seed, t, stream, first, count, _ptr(out))
        return out

    def gaps(self, seed, rho, m, first, count):
        """Fixed-point Gamma gaps (scale 2^20) of groups [first, first+count) of an event with m draws."""
        out = np.empty(count, dtype=np.uint64)
        self.L.orc_fill_gaps(seed, rho, m, first, count, _ptr(out, _up))
        return out

    def gap_variate(self, seed, rho, group, shape):
        return int(self.L.orc_gap_variate(seed, rho, group, shape))

    def gap_head(self, seed, rho, m):
        return int(self.L.orc_gap_head(seed, rho, m))

    def logsumexp(self, a):
        a = _d(a)
        return self.L.orc_logsumexp(_ptr(a), a.size)

    def effective_sample_size(self, lnw):
        a = _d(lnw)
        return self.L.orc_effective_sample_size(_ptr(a), a.size)

    def sincospi(self, t):
        s, c = C.c_double(), C.c_double()
        self.L.orc_sincospi(t, C.byref(s), C.byref(c))
        return s.value, c.value

    def logpdf_categorical(self, x, probs):
        p = _d(probs)
        return self.L.orc_logpdf_categorical(int(x), _ptr(p), p.size)

    def quantise_weights(self, lw, n_global=None):
        lw = _d(lw)
        q = np.empty(lw.size, dtype=np.uint64)
        m = C.c_double()
        self.L.orc_quantise_weights(_ptr(lw), lw.size, n_global or lw.size, _ptr(q, _up),
                                    C.cast(C.byref(m), _dp))
        return q, m.value

    def search_iid(self, cdf, u):
        cdf = np.ascontiguousarray(cdf, dtype=np.uint64)
        u = _d(u)
        anc = np.empty(u.size, dtype=np.int64)
        self.L.orc_search_iid(_ptr(cdf, _up), cdf.size, _ptr(u), u.size, _ptr(anc, _ip))
        return anc

    def search_sorted(self, cdf, seed, rho, m):
        """Ancestors of the m sorted-by-group draws of event rho against the inclusive integer CDF."""
        cdf = np.ascontiguousarray(cdf, dtype=np.uint64)
        anc = np.empty(m, dtype=np.int64)
        self.L.orc_search_sorted(_ptr(cdf, _up), cdf.size, seed, rho, m, _ptr(anc, _ip))
        return anc

    # ---- importance sampling ---------------------------------------------
    def importance_sampling(self, family, params, obs, num_samples, seed=0, proposal=PROPOSAL_DEFAULT,
                            prop_params=None, z_replay=None, num_threads=1):
        params, obs, pp, z = _d(params), _d(obs), _d(prop_params), _d(z_replay)
        D = {REGRESSION: 2, NORMAL_NORMAL: 1, OUTLIER_REGRESSION: 12, UNIFORM_NORMAL: 1}[family]
        lat = np.empty((D, num_samples), dtype=np.float64)
        lnw = np.empty(num_samples, dtype=np.float64)
        lml = C.c_double()
        rc = self.L.orc_importance_sampling(family, _ptr(params), params.size, _ptr(obs), obs.size, proposal,
                                            _ptr(pp), 0 if pp is None else pp.size, num_samples, seed,
                                            _ptr(z), _ptr(lat), _ptr(lnw), C.cast(C.byref(lml), _dp), num_threads)
        if rc:
            raise RuntimeError(self.err())
        return lat, lnw, lml.value

    def particle_filter(self, family, params, num_particles, seed=0, keep_history=False, num_threads=1):
        return OraclePF(self, family, params, num_particles, seed, keep_history, num_threads)


class OraclePF:
    """ParticleFilterState of the oracle (src/inference/particle_filter.jl:18-24)."""

    def __init__(self, orc, family, params, N, seed, keep_history, num_threads):
        self.o, self.L = orc, orc.L
        p = _d(params)
        self.h = self.L.orc_pf_create(family, _ptr(p), p.size, N, seed, int(keep_history), num_threads)
        if not self.h:
            raise RuntimeError(orc.err())
        self.N = N
        self.D = self.L.orc_pf_state_dim(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_pf_destroy(self.h)
            self.h = None

    def _prop(self, fn, obs, proposal, z, u):
        z, u = _d(z), _d(u)
        if obs is None:                       # unobserved step: the observation choice is sampled, weight += 0
            rc = fn(self.h, None, 0, proposal, None, 0, _ptr(z), _ptr(u))
        else:
            obs = _d(np.atleast_1d(obs))
            rc = fn(self.h, _ptr(obs), obs.size, proposal, None, 0, _ptr(z), _ptr(u))
        if rc:
            raise RuntimeError(self.o.err())

    def init(self, obs, proposal=PROPOSAL_DEFAULT, z_replay=None, u_replay=None):
        self._prop(self.L.orc_pf_init, obs, proposal, z_replay, u_replay)

    def step(self, obs, proposal=PROPOSAL_DEFAULT, z_replay=None, u_replay=None):
        self._prop(self.L.orc_pf_step, obs, proposal, z_replay, u_replay)

    def maybe_resample(self, ess_threshold=None, scheme=MULTINOMIAL, u_replay=None):
        if ess_threshold is None:
            ess_threshold = self.N / 2
        u = _d(u_replay)
        did, ess, lt = C.c_int(), C.c_double(), C.c_double()
        rc = self.L.orc_pf_maybe_resample(self.h, ess_threshold, scheme, _ptr(u), C.byref(did),
                                          C.cast(C.byref(ess), _dp), C.cast(C.byref(lt), _dp))
        if rc:
            raise RuntimeError(self.o.err())
        self.last_ess, self.last_log_total = ess.value, lt.value
        return bool(did.value)

    def log_ml_estimate(self):
        return self.L.orc_pf_log_ml_estimate(self.h)

    def log_weights(self):
        return np.ctypeslib.as_array(self.L.orc_pf_log_weights(self.h), (self.N,)).copy()

    def set_log_weights(self, lw):
        lw = _d(lw)
        assert lw.size == self.N
        self.L.orc_pf_set_log_weights(self.h, _ptr(lw))

    def parents(self):
        return np.ctypeslib.as_array(self.L.orc_pf_parents(self.h), (self.N,)).copy()

    def state(self):
        return np.ctypeslib.as_array(self.L.orc_pf_state(self.h), (self.D, self.N)).copy()

    def num_steps(self):
        return self.L.orc_pf_num_steps(self.h)

    def history(self, t):
        out = np.empty((self.D, self.N), dtype=np.float64)
        if self.L.orc_pf_history(self.h, t, _ptr(out)):
            raise RuntimeError(self.o.err())
        return out

    def sampled_observation(self, t):
        """Sampled observation choices of UNOBSERVED step t, in the particles' current order."""
        out = np.empty(self.N, dtype=np.float64)
        if self.L.orc_pf_sampled_observation(self.h, t, _ptr(out)):
            raise RuntimeError(self.o.err())
        return out

    def sample_unweighted(self, num_samples, u_replay=None):
        u = _d(u_replay)
        idx = np.empty(num_samples, dtype=np.int64)
        if self.L.orc_pf_sample_unweighted(self.h, num_samples, _ptr(u), _ptr(idx, _ip)):
            raise RuntimeError(self.o.err())
        return idx
