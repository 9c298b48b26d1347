# GenB200.jl -- Julia host side of libgensmc.so.
#
# Adds methods to Gen's own generic functions (src/inference/particle_filter.jl:215-216,
# src/inference/importance.jl:110) for catalogue models, so that an inference program written
# against Gen's API runs unchanged on a B200 by swapping the model object:
#
#     model = GenB200.LinearGaussianSSM(0.0, 1.0, 0.9, 0.0, 1.0, 1.0, 1.0)
#     state = initialize_particle_filter(model, (1,), choicemap((:y_init, ys[1])), 2^24)
#     for T in 2:length(ys)
#         maybe_resample!(state)
#         particle_filter_step!(state, (T,), (UnknownChange(),), choicemap((:chain => T-1 => :y, ys[T])))
#     end
#     log_ml_estimate(state)
#
# NOT EXECUTED in the build environment (no Julia there); the same C ABI is exercised by the Python
# mirror gen_b200/inference.py, which the tests drive.
module GenB200

using Gen
import Gen: initialize_particle_filter, particle_filter_step!, maybe_resample!, log_ml_estimate,
            get_log_weights, get_traces, sample_unweighted_traces, importance_sampling

const LIB = get(ENV, "GENSMC_LIB", joinpath(@__DIR__, "..", "gen_b200", "libgensmc.so"))

# ---- include/gen_b200.h -------------------------------------------------------------------------
struct Config
    struct_size::UInt32
    model_id::Int32
    dtype::Int32
    resample_scheme::Int32
    num_particles::UInt64
    seed::UInt64
    device::Int32
    keep_history::Int32
    history_capacity::Int64
    stream::Ptr{Cvoid}
end

const MODEL_HMM, MODEL_LGSSM, MODEL_SV, MODEL_BEARINGS, MODEL_REGRESSION, MODEL_NORMAL_NORMAL = 1, 2, 3, 4, 5, 6

function check(rc::Cint, h = C_NULL)
    rc == 0 && return
    msg = unsafe_string(ccall((:gsmc_last_error, LIB), Cstring, (Ptr{Cvoid},), h))
    error("libgensmc error $rc: $msg")      # the reference signals failures with error(...)
end

# ---- catalogue models: GenerativeFunction subtypes used only for dispatch -----------------------
abstract type DeviceSSM <: GenerativeFunction{Any,Trace} end

struct LinearGaussianSSM <: DeviceSSM
    m0::Float64; s0::Float64; a::Float64; b::Float64; q::Float64; c::Float64; r::Float64
end
params(m::LinearGaussianSSM) = Float64[m.m0, m.s0, m.a, m.b, m.q, m.c, m.r]
model_id(::LinearGaussianSSM) = MODEL_LGSSM
state_names(::LinearGaussianSSM) = (:x,)
obs_name(::LinearGaussianSSM) = :y

struct HMM <: DeviceSSM
    prior::Vector{Float64}
    emission_dists::Matrix{Float64}     # [x, z], as in test/inference/particle_filter.jl:54-58
    transition_dists::Matrix{Float64}   # [z, z_prev]
end
params(m::HMM) = vcat(Float64[length(m.prior), size(m.emission_dists, 1)], m.prior,
                      vec(m.transition_dists), vec(m.emission_dists))   # column-major = rows by z_prev / by z
model_id(::HMM) = MODEL_HMM
state_names(::HMM) = (:z,)
obs_name(::HMM) = :x

struct StochasticVolatility <: DeviceSSM
    mu::Float64; phi::Float64; sigma::Float64
end
params(m::StochasticVolatility) = Float64[m.mu, m.phi, m.sigma]
model_id(::StochasticVolatility) = MODEL_SV
state_names(::StochasticVolatility) = (:h,)
obs_name(::StochasticVolatility) = :y

"A catalogue proposal (the `proposal::GenerativeFunction` argument)."
struct DeviceProposal <: GenerativeFunction{Any,Trace}
    params::Vector{Float64}
end

obs_address(m::DeviceSSM, T::Int) = T == 1 ? Symbol(obs_name(m), :_init) : (:chain => (T - 1) => obs_name(m))

function observation_vector(m::DeviceSSM, T::Int, observations::ChoiceMap)
    addr = obs_address(m, T)
    has_value(observations, addr) || error("observations must constrain $addr")
    # like src/dynamic/update.jl:191-193: constraints the model does not visit are an error
    n = length(collect(get_values_shallow(observations))) + sum(Int[1 for _ in get_submaps_shallow(observations)])
    n == 1 || error("constraints at addresses the model does not visit at this step")
    Float64[observations[addr]]
end

# ---- the device-resident ParticleFilterState (particle_filter.jl:18-24) --------------------------
mutable struct DeviceParticleFilterState{M<:DeviceSSM}
    handle::Ptr{Cvoid}
    model::M
    num_particles::Int
    T::Int
    observations::Vector{Vector{Float64}}
end

function create(model::DeviceSSM, num_particles::Int; seed = 0, dtype = 0, resample = 0, keep_history = true,
                history_capacity = 128, device = -1)
    cfg = Ref(Config(sizeof(Config), model_id(model), dtype, resample, num_particles, seed, device,
                     keep_history ? 1 : 0, history_capacity, C_NULL))
    p = params(model)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:gsmc_create, LIB), Cint, (Ref{Config}, Ptr{Float64}, Csize_t, Ref{Ptr{Cvoid}}), cfg, p, length(p), h))
    state = DeviceParticleFilterState(h[], model, num_particles, 0, Vector{Float64}[])
    finalizer(s -> ccall((:gsmc_destroy, LIB), Cvoid, (Ptr{Cvoid},), s.handle), state)
    state
end

proposal_args(p::Nothing) = (0, Float64[])
proposal_args(p::DeviceProposal) = (1, p.params)

function propagate!(state, fn::Symbol, obs::Vector{Float64}, proposal)
    (pid, pp) = proposal_args(proposal)
    rc = fn == :gsmc_init ?
        ccall((:gsmc_init, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Csize_t, Cint, Ptr{Float64}, Csize_t), state.handle, obs, length(obs), pid, pp, length(pp)) :
        ccall((:gsmc_step, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Csize_t, Cint, Ptr{Float64}, Csize_t), state.handle, obs, length(obs), pid, pp, length(pp))
    check(rc, state.handle)
    state.T += 1
    push!(state.observations, obs)
    nothing
end

# particle_filter.jl:99-108
function initialize_particle_filter(model::DeviceSSM, model_args::Tuple, observations::ChoiceMap, num_particles::Int; kwargs...)
    model_args == (1,) || error("the filter starts with one time step: model_args must be (1,)")
    state = create(model, num_particles; kwargs...)
    propagate!(state, :gsmc_init, observation_vector(model, 1, observations), nothing)
    state
end

# particle_filter.jl:79-91
function initialize_particle_filter(model::DeviceSSM, model_args::Tuple, observations::ChoiceMap,
                                    proposal::DeviceProposal, proposal_args::Tuple, num_particles::Int; kwargs...)
    model_args == (1,) || error("the filter starts with one time step: model_args must be (1,)")
    state = create(model, num_particles; kwargs...)
    propagate!(state, :gsmc_init, observation_vector(model, 1, observations), proposal)
    state
end

# particle_filter.jl:162-180
function particle_filter_step!(state::DeviceParticleFilterState, new_args::Tuple, argdiffs::Tuple, observations::ChoiceMap)
    new_args == (state.T + 1,) || error("new_args must be ($(state.T + 1),): a step extends the traces by one time step")
    propagate!(state, :gsmc_step, observation_vector(state.model, state.T + 1, observations), nothing)
end

# particle_filter.jl:139-154 (SimpleExtendingTraceTranslator weight rule, trace_translators.jl:783-802)
function particle_filter_step!(state::DeviceParticleFilterState, new_args::Tuple, argdiffs::Tuple, observations::ChoiceMap,
                               proposal::DeviceProposal, proposal_args::Tuple)
    new_args == (state.T + 1,) || error("new_args must be ($(state.T + 1),): a step extends the traces by one time step")
    propagate!(state, :gsmc_step, observation_vector(state.model, state.T + 1, observations), proposal)
end

# particle_filter.jl:189-213
function maybe_resample!(state::DeviceParticleFilterState; ess_threshold::Real = state.num_particles / 2, verbose = false)
    did = Ref{Cint}(0); ess = Ref{Float64}(0.0)
    check(ccall((:gsmc_maybe_resample, LIB), Cint, (Ptr{Cvoid}, Float64, Ref{Cint}, Ref{Float64}),
                state.handle, Float64(ess_threshold), did, ess), state.handle)
    do_resample = did[] != 0
    verbose && println("effective sample size: $(ess[]), doing resample: $do_resample")
    do_resample
end

# particle_filter.jl:52-55
function log_ml_estimate(state::DeviceParticleFilterState)
    out = Ref{Float64}(0.0)
    check(ccall((:gsmc_log_ml_estimate, LIB), Cint, (Ptr{Cvoid}, Ref{Float64}), state.handle, out), state.handle)
    out[]
end

# particle_filter.jl:43-45
function get_log_weights(state::DeviceParticleFilterState)
    lw = Vector{Float64}(undef, state.num_particles)
    check(ccall((:gsmc_get_log_weights, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Csize_t), state.handle, lw, length(lw)), state.handle)
    lw
end

"One particle's choices, materialised from the device columns under the reference's addresses."
function trace_choices(state::DeviceParticleFilterState, i::Int)
    D = length(state_names(state.model))
    idx = Int64[i - 1]                                  # 0-based in the C ABI
    out = Vector{Float64}(undef, state.T * D)
    check(ccall((:gsmc_get_trajectories, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Csize_t, Ptr{Float64}, Csize_t),
                state.handle, idx, 1, out, length(out)), state.handle)
    cm = choicemap()
    for t in 1:state.T, (d, name) in enumerate(state_names(state.model))
        v = out[(t - 1) * D + d]
        addr = t == 1 ? Symbol(name, :_init) : (:chain => (t - 1) => name)
        cm[addr] = state.model isa HMM ? Int(v) : v
        y = state.observations[t][1]
        cm[obs_address(state.model, t)] = state.model isa HMM ? Int(y) : y
    end
    cm
end

struct DeviceTraces{S}
    state::S
end
Base.length(t::DeviceTraces) = t.state.num_particles
Base.getindex(t::DeviceTraces, i::Int) = trace_choices(t.state, i)
Base.iterate(t::DeviceTraces, i = 1) = i > length(t) ? nothing : (t[i], i + 1)

# particle_filter.jl:31-34
get_traces(state::DeviceParticleFilterState) = DeviceTraces(state)

# particle_filter.jl:62-70
function sample_unweighted_traces(state::DeviceParticleFilterState, num_samples::Int)
    idx = Vector{Int64}(undef, num_samples)
    check(ccall((:gsmc_sample_unweighted, LIB), Cint, (Ptr{Cvoid}, UInt64, Ptr{Int64}), state.handle, num_samples, idx), state.handle)
    [trace_choices(state, Int(j) + 1) for j in idx]
end

# importance.jl:20-33 for state-space models: generate(model, (T,), observations) = init + T-1 extensions
function importance_sampling(model::DeviceSSM, model_args::Tuple, observations::ChoiceMap, num_samples::Int, verbose = false; kwargs...)
    (T,) = model_args
    state = create(model, num_samples; kwargs...)
    for t in 1:T
        addr = obs_address(model, t)
        propagate!(state, t == 1 ? :gsmc_init : :gsmc_step, Float64[observations[addr]], nothing)
    end
    lml = log_ml_estimate(state)                       # nothing folded: log_total - log(n)
    (get_traces(state), get_log_weights(state) .- (lml + log(num_samples)), lml)
end

# importance.jl:70-87: sampling importance resampling that returns ONE trace. The reference keeps a reservoir of size
# one while it streams the samples; here the samples are generated `chunk_size` at a time on the device, the kept trace
# of a chunk is one categorical draw from the chunk's weights, and chunks are merged with the reference's rule
# (`bernoulli(exp(log_weight - log_total_weight))` with the chunk's total in the place of one sample's weight).
chunk_seed(seed::UInt64, c::Int) = seed + UInt64(c) * 0x9E3779B97F4A7C15
function importance_resampling(model::DeviceSSM, model_args::Tuple, observations::ChoiceMap, num_samples::Int, verbose = false;
                               seed::UInt64 = UInt64(0), chunk_size::Int = 1 << 24, kwargs...)
    log_total, kept, done, c = -Inf, nothing, 0, 0
    while done < num_samples
        m = min(chunk_size, num_samples - done)
        (traces, _, lml_c) = importance_sampling(model, model_args, observations, m; seed = chunk_seed(seed, c), keep_history = true, kwargs...)
        lt_c = lml_c + log(m)
        cand = sample_unweighted_traces(traces.state, 1)[1]
        new_total = kept === nothing ? lt_c : logsumexp(log_total, lt_c)        # inference.jl:8-11
        # the merge draw is host-side: Philox (seed, element 2c, event 0xffffffff, stream 3), see gen_b200/philox.py
        if kept === nothing || philox_uniform(seed, 2 * c, 0xffffffff, 3) < exp(lt_c - new_total)
            kept = cand
        end
        log_total = new_total
        finalize(traces.state)                            # runs gsmc_destroy now: device memory stays bounded by one chunk
        done += m
        c += 1
        verbose && println("sample: $done of $num_samples")
    end
    (kept, log_total - log(num_samples))
end

"Philox4x32-10 with the draw layout of gsmc_rng.cuh: element e of the uniform array of (seed, t, stream)."
function philox_uniform(seed::UInt64, e::Int, t::Integer, stream::Integer)
    call = UInt64(e >> 1)
    c0, c1, c2, c3 = UInt32(call & 0xffffffff), UInt32(call >> 32), UInt32(t), UInt32(stream)
    k0, k1 = UInt32(seed & 0xffffffff), UInt32(seed >> 32)
    for _ in 1:10
        p0, p1 = UInt64(0xD2511F53) * c0, UInt64(0xCD9E8D57) * c2
        c0, c1, c2, c3 = UInt32(p1 >> 32) ⊻ c1 ⊻ k0, UInt32(p1 & 0xffffffff), UInt32(p0 >> 32) ⊻ c3 ⊻ k1, UInt32(p0 & 0xffffffff)
        k0 += 0x9E3779B9; k1 += 0xBB67AE85
    end
    w = isodd(e) ? (UInt64(c2) | (UInt64(c3) << 32)) : (UInt64(c0) | (UInt64(c1) << 32))
    Float64(w >> 11) * 2.0^-53
end

export LinearGaussianSSM, HMM, StochasticVolatility, DeviceProposal, DeviceParticleFilterState, importance_resampling

end # module
