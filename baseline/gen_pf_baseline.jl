# gen_pf_baseline.jl -- the CPU figure SURVEY.md section 8(d)(ii) asks for: Gen.jl's OWN particle filter
# (initialize_particle_filter / maybe_resample! / particle_filter_step!, src/inference/particle_filter.jl) on the cfg-3
# model written the way the reference writes state-space models: a static-IR kernel under Unfold
# (test/modeling_library/unfold.jl:5-8 + an observation), driven by the loop of test/inference/particle_filter.jl:130-137.
#
# NOT MEASURED in the build environment (no Julia there; the bench's `cpu_baseline` is the C port in oracle/, which
# flatters Gen.jl: no per-particle trace allocation). Run wherever Julia >= 1.4 and Gen 0.4.1 exist:
#
#     julia --project=/path/to/Gen baseline/gen_pf_baseline.jl [log2_particles=16] [T=100]
#
# Prints one JSON line: particle-steps/s on ONE core (Gen's particle filter is single-threaded) and the log-ML estimate
# next to the exact Kalman value, for the same synthetic observations bench.py uses when given through a file.
using Gen
using Random

@gen (static) function lg_kernel(t::Int, x_prev::Float64, a::Float64, b::Float64, q::Float64, c::Float64, r::Float64)
    x = @trace(normal(x_prev * a + b, q), :x)
    @trace(normal(c * x, r), :y)
    return x
end
const lg_chain = Unfold(lg_kernel)

@gen (static) function lg_model(T::Int, m0::Float64, s0::Float64, a::Float64, b::Float64, q::Float64, c::Float64, r::Float64)
    x_init = @trace(normal(m0, s0), :x_init)
    @trace(normal(c * x_init, r), :y_init)
    @trace(lg_chain(T - 1, x_init, a, b, q, c, r), :chain)
end
Gen.load_generated_functions()

function kalman_log_ml(ys, m0, s0, a, b, q, c, r)
    m, P, ll = m0, s0^2, 0.0
    for (t, y) in enumerate(ys)
        if t > 1
            m, P = a * m + b, a^2 * P + q^2
        end
        S = c^2 * P + r^2
        ll += -0.5 * (y - c * m)^2 / S - 0.5 * log(2pi * S)
        K = c * P / S
        m, P = m + K * (y - c * m), (1 - K * c) * P
    end
    ll
end

function simulate(T, m0, s0, a, b, q, c, r)
    Random.seed!(0)
    x = m0 + s0 * randn()
    ys = Float64[]
    for t in 1:T
        t > 1 && (x = a * x + b + q * randn())
        push!(ys, c * x + r * randn())
    end
    ys
end

function run(num_particles::Int, ys::Vector{Float64}, p)
    args(T) = (T, p...)
    state = initialize_particle_filter(lg_model, args(1), choicemap((:y_init, ys[1])), num_particles)
    for T in 2:length(ys)
        maybe_resample!(state, ess_threshold = num_particles / 2)
        # only T changes: (UnknownChange, NoChange...) lets Unfold extend by one step instead of revisiting all
        argdiffs = (UnknownChange(), ntuple(_ -> NoChange(), length(p))...)
        particle_filter_step!(state, args(T), argdiffs, choicemap((:chain => (T - 1) => :y, ys[T])))
    end
    log_ml_estimate(state)
end

function main()
    log2n = length(ARGS) >= 1 ? parse(Int, ARGS[1]) : 16
    T = length(ARGS) >= 2 ? parse(Int, ARGS[2]) : 100
    p = (0.0, 1.0, 0.9, 0.0, 1.0, 1.0, 1.0)
    ys = simulate(T, p...)
    n = 1 << log2n
    run(min(n, 1024), ys[1:min(T, 5)], p)                       # compile
    t0 = time()
    lml = run(n, ys, p)
    dt = time() - t0
    println("{\"impl\": \"Gen.jl $(pkgversion(Gen))\", \"metric\": \"particle-steps/sec\", \"value\": $(n * T / dt), \"cores\": 1, ",
            "\"particles\": $n, \"time_steps\": $T, \"seconds\": $dt, \"log_ml\": $lml, \"log_ml_kalman\": $(kalman_log_ml(ys, p...))}")
end

main()
