"""Static IR -> CUDA code generation (gen_b200/staticir.py; the reference's counterpart is
src/static_ir/dag.jl:1-46 + src/static_ir/generate.jl:68-116). The catalogue's LG-SSM and stochastic-volatility kernels
written in the IR must reproduce the hand-written functors and the CPU oracle bit for bit
(src/static_ir/generate.jl:24-43: evaluation order = order in which random draws are consumed)."""
import numpy as np
import pytest

import gen_b200 as g
from gen_b200 import staticir as ir
from oracle import closed_forms as cf
from oracle import oracle as O

LG = [0.0, 1.0, 0.9, 0.0, 1.0, 1.0, 1.0]
SVP = [-1.0, 0.97, 0.2]


def lgssm_kernel():
    k = ir.StaticKernel("lgssm_ir", params=["m0", "s0", "a", "b", "q", "c", "r"], state=["x"], obs="y")
    x0 = k.init.trace("x", ir.normal(ir.Param("m0"), ir.Param("s0")))                 # x_init ~ normal(m0, s0)
    k.init.observe("y", ir.normal(ir.Param("c") * ir.New("x"), ir.Param("r")))         # y_init ~ normal(c * x, r)
    k.init.ret(x=x0)
    x = k.step.trace("x", ir.normal(ir.Prev("x") * ir.Param("a") + ir.Param("b"), ir.Param("q")))
    k.step.observe("y", ir.normal(ir.Param("c") * ir.New("x"), ir.Param("r")))
    k.step.ret(x=x)
    return k


def sv_kernel():
    # examples/pmmh/model.jl:40-50 pattern: h_init ~ normal(mu, sigma / sqrt(1 - phi*phi)); h ~ normal(mu + phi*(h_prev - mu), sigma);
    # y ~ normal(0, exp(h / 2))
    k = ir.StaticKernel("sv_ir", params=["mu", "phi", "sigma"], state=["h"], obs="y")
    mu, phi, sigma = ir.Param("mu"), ir.Param("phi"), ir.Param("sigma")
    h0 = k.init.trace("h", ir.normal(mu, sigma / ir.sqrt(1.0 - phi * phi)))
    k.init.observe("y", ir.normal(0.0, ir.exp(ir.New("h") / 2.0)))
    k.init.ret(h=h0)
    h = k.step.trace("h", ir.normal(mu + phi * (ir.Prev("h") - mu), sigma))
    k.step.observe("y", ir.normal(0.0, ir.exp(ir.New("h") / 2.0)))
    k.step.ret(h=h)
    return k


def cv2d_kernel():
    """A model that is NOT in the catalogue: 2-D constant-velocity target, noisy observation of x + y.
    State (x, vx, y, vy) is a function of two latent choices (wx, wy), as in the bearings model."""
    k = ir.StaticKernel("cv2d_ir", params=["sw", "so"], state=["x", "vx", "y", "vy"], obs="z")
    sw, so = ir.Param("sw"), ir.Param("so")
    x0 = k.init.trace("x", ir.normal(0.0, 1.0))
    vx0 = k.init.trace("vx", ir.normal(0.0, 0.1))
    y0 = k.init.trace("y", ir.normal(0.0, 1.0))
    vy0 = k.init.trace("vy", ir.normal(0.0, 0.1))
    k.init.observe("z", ir.normal(ir.New("x") + ir.New("y"), so))
    k.init.ret(x=x0, vx=vx0, y=y0, vy=vy0)
    wx = k.step.trace("wx", ir.normal(0.0, sw))
    wy = k.step.trace("wy", ir.normal(0.0, sw))
    k.step.observe("z", ir.normal(ir.New("x") + ir.New("y"), so))
    k.step.ret(x=ir.Prev("x") + ir.Prev("vx") + 0.5 * wx, vx=ir.Prev("vx") + wx, y=ir.Prev("y") + ir.Prev("vy") + 0.5 * wy, vy=ir.Prev("vy") + wy)
    return k


def test_codegen_emits_nodes_in_evaluation_order():
    src = sv_kernel().source()
    assert "random_normal(p[0], (p[2] / sqrt(((0x1.0000000000000p+0) - (p[1] * p[1])))), z[0])" in src
    assert "random_normal((p[0] + (p[1] * (prev[0] - p[0]))), p[2], z[0])" in src
    assert "logpdf_normal(obs, (0x0.0p+0), gm_exp((n_h / (0x1.0000000000000p+1))))" in src
    src = cv2d_kernel().source()
    step = src[src.index("} else {"):]
    assert step.index("c_wx = random_normal") < step.index("c_wy = random_normal") < step.index("w += logpdf_normal")
    assert "z[0]" in src and "z[1]" in src and "nz(bool init, int) { return init ? 4 : 2; }" in src
    with pytest.raises(ValueError):                                  # choice used before it is traced
        k = ir.StaticKernel("bad", params=["a"], state=["x"])
        k.init.trace("x", ir.normal(ir.Choice("later"), 1.0))
        k.init.observe("y", ir.normal(0.0, 1.0))
        k.init.ret(x=ir.Choice("x"))
        k.init.check()
    with pytest.raises(ValueError):
        lgssm_kernel().init.trace("x", ir.normal(0.0, 1.0))          # address traced twice


def test_plugin_builds_and_registers_without_a_gpu():
    Model = lgssm_kernel().compile()
    assert Model.family >= 1000 and Model.state_names == ("x",)
    again = lgssm_kernel().compile()
    assert again.family == Model.family                               # same source -> same cached plugin, same id
    m = Model(*LG)
    assert np.array_equal(m.params(), np.array(LG))
    with pytest.raises(TypeError):
        Model(1.0, 2.0)


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["lgssm", "sv"])
def test_generated_kernels_reproduce_the_hand_written_functors(orc, which):
    if which == "lgssm":
        Model, hand, fam, params, ys = lgssm_kernel().compile(), g.LinearGaussianSSM(*LG), O.LGSSM, LG, cf.simulate_lgssm(20, LG, 3)
    else:
        Model, hand, fam, params, ys = sv_kernel().compile(), g.StochasticVolatility(*SVP), O.SV, SVP, cf.simulate_sv(20, SVP, 4)
    N, T = 20011, 14

    def bits(a):
        return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)
    gen = g.ParticleFilterState(Model(*params), N, seed=31, keep_history=True, history_capacity=T)
    ref = g.ParticleFilterState(hand, N, seed=31, keep_history=True, history_capacity=T)
    pf = orc.particle_filter(fam, params, N, seed=31, keep_history=True)
    for s in (gen, ref, pf):
        s.init([ys[0]])
    n_res = 0
    for t in range(1, T):
        d = [s.maybe_resample(N / 2) for s in (gen, ref, pf)]
        assert d[0] == d[1] == d[2]
        n_res += d[0]
        if d[0]:
            assert np.array_equal(gen.ancestors(), pf.parents())
        obs = None if t == 5 else [ys[t]]                             # one unobserved step: the generated observation sampler
        for s in (gen, ref, pf):
            s.step(obs)
        assert np.array_equal(bits(gen.log_weights()), bits(ref.log_weights())) and np.array_equal(bits(gen.log_weights()), bits(pf.log_weights())), t
        assert np.array_equal(bits(gen.state()), bits(ref.state())) and np.array_equal(bits(gen.state()), bits(pf.state())), t
    assert n_res >= 1
    assert np.array_equal(bits(gen.sampled_observation(6)), bits(pf.sampled_observation(6)))
    assert gen.log_ml_estimate() == ref.log_ml_estimate()
    assert np.array_equal(bits(gen.state(3)), bits(pf.history(3)))
    # the sync-free loop runs generated models too
    gen2 = g.ParticleFilterState(Model(*params), N, seed=31, keep_history=False)
    ref2 = g.ParticleFilterState(hand, N, seed=31, keep_history=False)
    for s in (gen2, ref2):
        s.init([ys[0]])
        s.run_steps(ys[1:T], N / 2)
    assert gen2.log_ml_estimate() == ref2.log_ml_estimate() and np.array_equal(bits(gen2.log_weights()), bits(ref2.log_weights()))
    for s in (gen, ref, gen2, ref2):
        s.close()


@pytest.mark.gpu
def test_generated_model_outside_the_catalogue():
    """A kernel the catalogue does not have (4 state fields computed from 2 latent choices): runs through the Gen API,
    is reproducible, and its log-ML estimate agrees with a second, larger run."""
    Model = cv2d_kernel().compile()
    model = Model(sw=0.05, so=0.3)
    rng = np.random.default_rng(0)
    zs = np.cumsum(0.1 + 0.05 * rng.standard_normal(15)) + 0.3 * rng.standard_normal(15)
    out = []
    for N in (1 << 16, 1 << 16, 1 << 19):
        st = g.initialize_particle_filter(model, (1,), g.choicemap(("z_init", float(zs[0]))), N, seed=2, keep_history=True, history_capacity=15)
        for T in range(2, 16):
            g.maybe_resample_b(st)
            g.particle_filter_step_b(st, (T,), (g.UnknownChange(),), g.choicemap((("chain", T - 1, "z"), float(zs[T - 1]))))
        out.append(g.log_ml_estimate(st))
        if N == 1 << 19:
            tr = g.get_traces(st)[0].get_choices()
            assert ("chain", 3, "vx") in tr and tr[("chain", 3, "z")] == zs[3] and "y_init" in tr
        st.close()
    assert out[0] == out[1] and np.isfinite(out[0])
    assert abs(out[0] - out[2]) < 0.1
