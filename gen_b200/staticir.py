"""Static IR -> CUDA code generation for state-space kernels beyond the catalogue.

The reference compiles a `@gen (static)` function into a DAG of nodes -- argument nodes, `JuliaNode`s (plain Julia
expressions), random-choice nodes with an address (src/static_ir/dag.jl:1-46, built by src/dsl/static.jl:248-281) -- and
generates per-node Julia code for `generate`/`update` inside generated functions (src/static_ir/generate.jl:68-116):
a constrained choice adds `logpdf` to the weight, an unconstrained one is `random(dist, args...)` (:24-43), in node
order (which is therefore also the order in which random draws are consumed).

Here the same structure is described in Python for a kernel under `Unfold` whose Julia nodes are arithmetic, and turned
into a CUDA model functor compiled into the library's `propagate_kernel` template:

    k = StaticKernel("lgssm", params=["m0", "s0", "a", "b", "q", "c", "r"], state=["x"], obs="y")
    x0 = k.init.trace("x", normal(Param("m0"), Param("s0")))          # x_init ~ normal(m0, s0)
    k.init.observe("y", normal(Param("c") * x0, Param("r")))          # y_init ~ normal(c * x, r)      (constrained)
    k.init.ret(x=x0)
    x = k.step.trace("x", normal(Prev("x") * Param("a") + Param("b"), Param("q")))
    k.step.observe("y", normal(Param("c") * x, Param("r")))
    k.step.ret(x=x)
    Model = k.compile()                                               # nvcc -> plugin .so -> gsmc_register_model_plugin
    state = initialize_particle_filter(Model(m0=0., s0=1., ...), (1,), choicemap(("y_init", 0.3)), 1 << 20)

Arithmetic is emitted literally (no reassociation, compiled with -fmad=false like the library), distributions call the
library's `random_normal` / `logpdf_normal` (normal.jl:56-60,96), so a kernel written here produces the same bits as
the hand-written catalogue functor and as the CPU oracle.
"""
import hashlib
import os
import subprocess

import numpy as np

from . import _lib
from .build import CSRC, nvcc_path
from .models import DeviceModel

PLUGIN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_plugins")


# ---------------------------------------------------------------------------------------------------------------------
# expression nodes (the arithmetic `JuliaNode`s)
# ---------------------------------------------------------------------------------------------------------------------
class Expr:
    def _bin(self, op, other, swap=False):
        other = other if isinstance(other, Expr) else Const(other)
        return BinOp(op, other, self) if swap else BinOp(op, self, other)

    def __add__(self, o): return self._bin("+", o)
    def __radd__(self, o): return self._bin("+", o, True)
    def __sub__(self, o): return self._bin("-", o)
    def __rsub__(self, o): return self._bin("-", o, True)
    def __mul__(self, o): return self._bin("*", o)
    def __rmul__(self, o): return self._bin("*", o, True)
    def __truediv__(self, o): return self._bin("/", o)
    def __rtruediv__(self, o): return self._bin("/", o, True)
    def __neg__(self): return Call("neg", self)

    def children(self):
        return ()

    def walk(self):
        yield self
        for c in self.children():
            yield from c.walk()


class Const(Expr):
    def __init__(self, v):
        self.v = float(v)


class Param(Expr):
    """A kernel argument that is the same for all particles (`params...` of the Unfold kernel)."""
    def __init__(self, name):
        self.name = name


class Prev(Expr):
    """A field of the previous state (the `prev_state` argument of the Unfold kernel)."""
    def __init__(self, name):
        self.name = name


class Choice(Expr):
    """The value of a random choice traced earlier in the same kernel."""
    def __init__(self, name):
        self.name = name


class New(Expr):
    """A field of the state this kernel returns (allowed in the arguments of the observed choice)."""
    def __init__(self, name):
        self.name = name


class BinOp(Expr):
    def __init__(self, op, a, b):
        self.op, self.a, self.b = op, a, b

    def children(self):
        return (self.a, self.b)


class Call(Expr):
    FUNCS = {"exp": "gm_exp", "log": "gm_log", "sqrt": "sqrt", "neg": "-"}

    def __init__(self, fn, a):
        if fn not in self.FUNCS:
            raise ValueError("unsupported function %r (arithmetic nodes: + - * / exp log sqrt)" % fn)
        self.fn, self.a = fn, a

    def children(self):
        return (self.a,)


def exp(x): return Call("exp", x if isinstance(x, Expr) else Const(x))
def log(x): return Call("log", x if isinstance(x, Expr) else Const(x))
def sqrt(x): return Call("sqrt", x if isinstance(x, Expr) else Const(x))


class normal:
    """normal(mu, std) (src/modeling_library/distributions/normal.jl)."""
    def __init__(self, mu, std):
        self.mu = mu if isinstance(mu, Expr) else Const(mu)
        self.std = std if isinstance(std, Expr) else Const(std)


# ---------------------------------------------------------------------------------------------------------------------
# one kernel body = the node list of a static IR, in evaluation order
# ---------------------------------------------------------------------------------------------------------------------
class _Body:
    def __init__(self, kernel, is_init):
        self.kernel, self.is_init = kernel, is_init
        self.nodes = []            # ("trace", name, dist) | ("observe", name, dist)
        self.returns = None

    def trace(self, name, dist):
        """`name ~ dist`: an unconstrained (latent) random choice."""
        if not isinstance(dist, normal):
            raise TypeError("generated kernels support `normal` choices")
        if any(n[1] == name for n in self.nodes):
            raise ValueError("address %r traced twice (src/dsl/static.jl rejects this too)" % name)
        self.nodes.append(("trace", name, dist))
        return Choice(name)

    def observe(self, name, dist):
        """`name ~ dist` for the address the particle filter constrains at every step."""
        if name != self.kernel.obs:
            raise ValueError("the observed address of this kernel is %r" % self.kernel.obs)
        if any(n[0] == "observe" for n in self.nodes):
            raise ValueError("one observed choice per kernel")
        self.nodes.append(("observe", name, dist))

    def ret(self, **fields):
        """The kernel's return value = the new state."""
        if sorted(fields) != sorted(self.kernel.state):
            raise ValueError("the state has the fields %r" % (self.kernel.state,))
        self.returns = {k: (v if isinstance(v, Expr) else Const(v)) for k, v in fields.items()}

    # --- code generation ---------------------------------------------------------------------------------------
    def emit_expr(self, e, mode):
        k = self.kernel
        if isinstance(e, Const):
            return "(%s)" % float(e.v).hex() if e.v == e.v and abs(e.v) != float("inf") else "(%r)" % e.v
        if isinstance(e, Param):
            return "p[%d]" % k.params.index(e.name)
        if isinstance(e, Prev):
            if self.is_init:
                raise ValueError("the init kernel has no previous state")
            if mode == "obs_sampler":
                raise _NoSampler()
            return "prev[%d]" % k.state.index(e.name)
        if isinstance(e, Choice):
            if mode == "obs_sampler":
                raise _NoSampler()
            return "c_%s" % e.name
        if isinstance(e, New):
            return "lat[%d]" % k.state.index(e.name) if mode == "obs_sampler" else "n_%s" % e.name
        if isinstance(e, BinOp):
            return "(%s %s %s)" % (self.emit_expr(e.a, mode), e.op, self.emit_expr(e.b, mode))
        if isinstance(e, Call):
            if e.fn == "neg":
                return "(-%s)" % self.emit_expr(e.a, mode)
            return "%s(%s)" % (Call.FUNCS[e.fn], self.emit_expr(e.a, mode))
        raise TypeError(e)

    def check(self):
        k = self.kernel
        if self.returns is None:
            raise ValueError("%s kernel: ret(...) missing" % ("init" if self.is_init else "step"))
        if not any(n[0] == "observe" for n in self.nodes):
            raise ValueError("the kernel must trace its observed address %r" % k.obs)
        known = set()
        for kind, name, dist in self.nodes:
            for e in list(dist.mu.walk()) + list(dist.std.walk()):
                if isinstance(e, Choice) and e.name not in known:
                    raise ValueError("choice %r is used before it is traced" % e.name)
                if isinstance(e, New) and kind != "observe":
                    raise ValueError("New(...) is only available to the observed choice")
                if isinstance(e, Param) and e.name not in k.params:
                    raise ValueError("unknown parameter %r" % e.name)
                if isinstance(e, (Prev, New)) and e.name not in k.state:
                    raise ValueError("unknown state field %r" % e.name)
            if kind == "trace":
                known.add(name)

    def emit(self):
        """C++ statements of particle<INIT>: latent choices consume z[0], z[1], ... in node order."""
        lines, nz = [], 0
        obs_seen = False
        ret_emitted = False

        def emit_ret():
            out = []
            for d, f in enumerate(self.kernel.state):
                out.append("      const double n_%s = %s;" % (f, self.emit_expr(self.returns[f], "particle")))
            return out
        for kind, name, dist in self.nodes:
            if kind == "trace":
                lines.append("      const double c_%s = random_normal(%s, %s, z[%d]);" % (name, self.emit_expr(dist.mu, "particle"), self.emit_expr(dist.std, "particle"), nz))
                nz += 1
            else:
                if _uses_new(dist) and not ret_emitted:
                    lines += emit_ret()
                    ret_emitted = True
                lines.append("      w += logpdf_normal(obs, %s, %s);" % (self.emit_expr(dist.mu, "particle"), self.emit_expr(dist.std, "particle")))
                obs_seen = True
        if not ret_emitted:
            lines += emit_ret()
        for d, f in enumerate(self.kernel.state):
            lines.append("      out[%d] = n_%s;" % (d, f))
        assert obs_seen
        return "\n".join(lines), nz

    def emit_obs_sampler(self):
        for kind, name, dist in self.nodes:
            if kind == "observe":
                try:
                    return "random_normal(%s, %s, z)" % (self.emit_expr(dist.mu, "obs_sampler"), self.emit_expr(dist.std, "obs_sampler"))
                except _NoSampler:
                    return None
        return None


class _NoSampler(Exception):
    pass


def _uses_new(dist):
    return any(isinstance(e, New) for e in list(dist.mu.walk()) + list(dist.std.walk()))


# ---------------------------------------------------------------------------------------------------------------------
# the kernel pair (init + step) and its compilation
# ---------------------------------------------------------------------------------------------------------------------
TEMPLATE = r'''// generated by gen_b200/staticir.py from the static IR of kernel "%(name)s" -- do not edit
#include "kernels.cuh"
#include "plugin.h"

struct GenModel {
  static constexpr int D = %(D)d;
  static constexpr int SMEM_DOUBLES = 0;
  __host__ __device__ static constexpr int nz(bool init, int) { return init ? %(nz_init)d : %(nz_step)d; }
  __host__ __device__ static constexpr int nu(bool, int) { return 0; }
  static constexpr bool OBS_DRAW_UNIFORM = false;
  template <bool INIT, int PROP>
  __device__ __forceinline__ static void prologue(const ModelArgs&, double*) {}
  template <bool INIT, int PROP>
  __device__ __forceinline__ static double particle(const ModelArgs& a, const double*, const double* prev, const double* z,
                                                    const double*, double* out) {
    const double* p = a.p;
    const double obs = a.obs[0];
    double w = 0.0;
    if (INIT) {
%(init_body)s
    } else {
%(step_body)s
    }
    return w;
  }
  __device__ __forceinline__ static double sample_obs(const ModelArgs& a, const double* lat, double z) {
    const double* p = a.p;
    (void)p; (void)lat;
    return %(obs_sampler)s;
  }
};

#define PLUGIN_API extern "C" __attribute__((visibility("default")))

PLUGIN_API int gsmc_plugin_describe(gsmc_plugin_info* o) {
  if (!o) return 1;
  o->abi = GSMC_PLUGIN_ABI; o->D = GenModel::D; o->n_params = %(n_params)d;
  o->nz_init = %(nz_init)d; o->nz_step = %(nz_step)d; o->has_obs_sampler = %(has_sampler)d;
  o->sizeof_prop_args = sizeof(PropArgs<double>); o->sizeof_model_args = sizeof(ModelArgs); o->sizeof_dev_scalars = sizeof(DevScalars);
  const char name[] = "%(name)s";
  for (size_t i = 0; i < sizeof name && i < sizeof o->name; ++i) o->name[i] = name[i];
  o->name[sizeof o->name - 1] = 0;
  return 0;
}

template <bool INIT>
static int launch(const PropArgs<double>& g0, const ModelArgs& a, int64_t n_pad, int sm_count, cudaStream_t stream, int pdl, int* n_blocks) {
  PropArgs<double> g = g0;
  static int occ_dev[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  int& occ = occ_dev[dev & 63];
  if (occ == 0 && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)propagate_kernel<GenModel, double, INIT, 0>, GSMC_BLOCK, 0) != cudaSuccess || occ < 1)) occ = 2;
  g.n_tiles = (int)(n_pad / PropTile<GenModel>::TILE);
  const int grid = g.n_tiles < sm_count * occ ? g.n_tiles : sm_count * occ;
  *n_blocks = grid;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(GSMC_BLOCK); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  return (int)cudaLaunchKernelEx(&cfg, propagate_kernel<GenModel, double, INIT, 0>, g, a);
}
PLUGIN_API int gsmc_plugin_propagate(const void* prop_args, const void* model_args, int init, int64_t n_pad, int sm_count, void* stream, int pdl, int* n_blocks) {
  const PropArgs<double>& g = *(const PropArgs<double>*)prop_args;
  const ModelArgs& a = *(const ModelArgs*)model_args;
  return init ? launch<true>(g, a, n_pad, sm_count, (cudaStream_t)stream, pdl, n_blocks) : launch<false>(g, a, n_pad, sm_count, (cudaStream_t)stream, pdl, n_blocks);
}
PLUGIN_API int gsmc_plugin_sample_obs(const void* model_args, const double* state, double* obs_col, int64_t n, int64_t stride, uint64_t first_global,
                                      uint64_t seed, uint32_t t, void* stream) {
  const ModelArgs& a = *(const ModelArgs*)model_args;
  sample_obs_kernel<GenModel, double><<<(int)((n + GSMC_BLOCK - 1) / GSMC_BLOCK), GSMC_BLOCK, 0, (cudaStream_t)stream>>>(a, state, obs_col, n, stride, first_global, seed, t);
  return (int)cudaGetLastError();
}
'''


class StaticKernel:
    """Init kernel + Unfold step kernel of a state-space model, as static IRs (see the module docstring)."""

    def __init__(self, name, params, state, obs="y"):
        if not name.replace("_", "").isalnum():
            raise ValueError("kernel name must be an identifier")
        self.name, self.params, self.state, self.obs = name, list(params), list(state), obs
        if len(self.params) > 32:
            raise ValueError("at most 32 parameters")
        if not 1 <= len(self.state) <= 16:
            raise ValueError("1..16 state fields")
        self.init, self.step = _Body(self, True), _Body(self, False)

    def source(self):
        self.init.check()
        self.step.check()
        init_body, nz_init = self.init.emit()
        step_body, nz_step = self.step.emit()
        sampler = self.step.emit_obs_sampler()
        sampler_init = self.init.emit_obs_sampler()
        has = sampler is not None and sampler == sampler_init        # one emission law for both kernels
        return TEMPLATE % {"name": self.name, "D": len(self.state), "nz_init": nz_init, "nz_step": nz_step, "n_params": len(self.params),
                           "init_body": init_body, "step_body": step_body, "obs_sampler": sampler if has else "0.0", "has_sampler": 1 if has else 0}

    def build(self, verbose=False):
        """CUDA source -> plugin .so (cached by content hash under gen_b200/_plugins/)."""
        src = self.source()
        headers = b"".join(open(os.path.join(CSRC, h), "rb").read() for h in sorted(os.listdir(CSRC)) if h.endswith((".h", ".cuh")))
        tag = hashlib.sha256(src.encode() + headers).hexdigest()[:16]
        os.makedirs(PLUGIN_DIR, exist_ok=True)
        so = os.path.join(PLUGIN_DIR, "%s_%s.so" % (self.name, tag))
        if os.path.exists(so):
            return so
        cu = os.path.join(PLUGIN_DIR, "%s_%s.cu" % (self.name, tag))
        with open(cu, "w") as fh:
            fh.write(src)
        cmd = [nvcc_path(), "-shared", "-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a", "-fmad=false",
               "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden", "-I", CSRC, "-o", so + ".tmp", cu]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise _lib.GsmcError(_lib.E_UNSUPPORTED, "nvcc failed on the generated model %s:\n%s" % (self.name, res.stderr[-4000:]))
        if verbose:
            print(res.stdout + res.stderr)
        os.replace(so + ".tmp", so)
        return so

    def compile(self):
        """Builds and registers the plugin; returns a DeviceModel class whose constructor takes the parameters by name."""
        so = self.build()
        lib = _lib.load()
        mid = _lib.C.c_int()
        _lib.check(lib.gsmc_register_model_plugin(so.encode(), _lib.C.byref(mid)))
        kernel = self

        class GeneratedSSM(DeviceModel):
            __doc__ = "State-space model generated from the static IR of kernel %r (parameters: %s)." % (kernel.name, ", ".join(kernel.params))
            family = mid.value
            state_names = tuple(kernel.state)
            obs_name = kernel.obs
            ir = kernel

            def __init__(self, *args, **kw):
                vals = dict(zip(kernel.params, args))
                vals.update(kw)
                missing = [p for p in kernel.params if p not in vals]
                extra = [p for p in vals if p not in kernel.params]
                if missing or extra or len(args) > len(kernel.params):
                    raise TypeError("parameters of %s: %s" % (kernel.name, ", ".join(kernel.params)))
                self.p = np.array([float(vals[p]) for p in kernel.params], dtype=np.float64)

            def params(self):
                return self.p

        GeneratedSSM.__name__ = "Generated_%s" % kernel.name
        return GeneratedSSM
