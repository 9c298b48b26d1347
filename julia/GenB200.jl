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
# mirror gen_b200/inference.py and by a plain-C caller (tests/c_abi_smoke.c), which the tests drive.
# `GenB200.validate()` is the probe to run first wherever Julia + Gen + a B200 exist.
module GenB200

using Gen
import Gen: initialize_particle_filter, particle_filter_step!, maybe_resample!, log_ml_estimate,
            get_log_weights, get_traces, sample_unweighted_traces, importance_sampling, importance_resampling

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
const MODEL_OUTLIER_REGRESSION, MODEL_UNIFORM_NORMAL = 7, 8

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

"2-D bearings-only tracking, state (x, vx, y, vy); custom proposal = one-step EKF update (BASELINE.json configs[5])."
struct BearingsOnly <: DeviceSSM
    prior_mean::NTuple{4,Float64}; prior_std::NTuple{4,Float64}; sigma_w::Float64; sigma_theta::Float64
end
BearingsOnly() = BearingsOnly((0.0, 0.0, 12.4, -0.05), (0.5, 0.005, 0.3, 0.01), 0.001, 0.005)
params(m::BearingsOnly) = Float64[m.prior_mean..., m.prior_std..., m.sigma_w, m.sigma_theta]
model_id(::BearingsOnly) = MODEL_BEARINGS
state_names(::BearingsOnly) = (:x, :vx, :y, :vy)
obs_name(::BearingsOnly) = :bearing

# ---- importance-sampling families (one "time step", no resampling) --------------------------------
abstract type DeviceISModel <: GenerativeFunction{Any,Trace} end

"examples/regression/quickstart.jl:3-9; model_args = (xs,); observations constrain \"y-\$i\"."
struct LinearRegression <: DeviceISModel
    sd_slope::Float64; sd_intercept::Float64; sd_noise::Float64
end
LinearRegression() = LinearRegression(2.0, 10.0, 1.0)
model_id(::LinearRegression) = MODEL_REGRESSION
is_params(m::LinearRegression, model_args) = (xs = Float64.(model_args[1]); vcat(Float64[length(xs), m.sd_slope, m.sd_intercept, m.sd_noise], xs))
is_observations(m::LinearRegression, model_args, obs::ChoiceMap) = Float64[obs["y-$i"] for i in 1:length(model_args[1])]
latent_names(::LinearRegression) = (:slope, :intercept)

"test/inference/importance_sampling.jl:3-12: x ~ normal(mu0, sd0); y ~ normal(x, sd_y)."
struct NormalNormal <: DeviceISModel
    mu0::Float64; sd0::Float64; sd_y::Float64
end
model_id(::NormalNormal) = MODEL_NORMAL_NORMAL
is_params(m::NormalNormal, model_args) = Float64[m.mu0, m.sd0, m.sd_y]
is_observations(m::NormalNormal, model_args, obs::ChoiceMap) = Float64[obs[:y]]
latent_names(::NormalNormal) = (:x,)

"examples/regression/static_model.jl:3-23 (bernoulli outlier flags, Map of the static `datum`); model_args = (xs,), n <= 256."
struct OutlierRegression <: DeviceISModel
    prob_outlier::Float64; prior_sd::Float64
end
OutlierRegression() = OutlierRegression(0.5, 2.0)
model_id(::OutlierRegression) = MODEL_OUTLIER_REGRESSION
is_params(m::OutlierRegression, model_args) = (xs = Float64.(model_args[1]); vcat(Float64[length(xs), m.prob_outlier, m.prior_sd], xs))
is_observations(m::OutlierRegression, model_args, obs::ChoiceMap) = Float64[obs[:data => i => :y] for i in 1:length(model_args[1])]
latent_names(::OutlierRegression) = (:log_inlier_std, :log_outlier_std, :slope, :intercept)

"x ~ uniform(low, high); y ~ normal(x, sd_y) (uniform_continuous.jl:12-23 on the device)."
struct UniformNormal <: DeviceISModel
    low::Float64; high::Float64; sd_y::Float64
end
model_id(::UniformNormal) = MODEL_UNIFORM_NORMAL
is_params(m::UniformNormal, model_args) = Float64[m.low, m.high, m.sd_y]
is_observations(m::UniformNormal, model_args, obs::ChoiceMap) = Float64[obs[:y]]
latent_names(::UniformNormal) = (:x,)

"A catalogue proposal (the `proposal::GenerativeFunction` argument)."
struct DeviceProposal <: GenerativeFunction{Any,Trace}
    params::Vector{Float64}
end

obs_address(m::DeviceSSM, T::Int) = T == 1 ? Symbol(obs_name(m), :_init) : (:chain => (T - 1) => obs_name(m))

function observation_vector(m::DeviceSSM, T::Int, observations::ChoiceMap)
    addr = obs_address(m, T)
    # an empty choice map: nothing is constrained at this step -- the reference samples the observation choice
    # (src/static_ir/generate.jl:36-42) and the weight does not change; the C ABI takes (NULL, 0) for that
    isempty(observations) && return Float64[]
    has_value(observations, addr) || error("constraints at addresses the model does not visit at this step")
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
                history_capacity = 128, device = -1, comm = nothing)
    cfg = Ref(Config(sizeof(Config), model_id(model), dtype, resample, num_particles, seed, device,
                     keep_history ? 1 : 0, history_capacity, C_NULL))
    p = params(model)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:gsmc_create, LIB), Cint, (Ref{Config}, Ptr{Float64}, Csize_t, Ref{Ptr{Cvoid}}), cfg, p, length(p), h))
    state = DeviceParticleFilterState(h[], model, num_particles, 0, Vector{Float64}[])
    finalizer(s -> ccall((:gsmc_destroy, LIB), Cvoid, (Ptr{Cvoid},), s.handle), state)
    comm === nothing || check(ccall((:gsmc_comm_attach, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), state.handle, comm.handle), state.handle)
    state
end

proposal_args(p::Nothing) = (0, Float64[])
proposal_args(p::DeviceProposal) = (1, p.params)

function propagate!(state, fn::Symbol, obs::Vector{Float64}, proposal)
    (pid, pp) = proposal_args(proposal)
    optr = isempty(obs) ? Ptr{Float64}(C_NULL) : pointer(obs)          # (NULL, 0): unobserved step
    rc = GC.@preserve obs (fn == :gsmc_init ?
        ccall((:gsmc_init, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Csize_t, Cint, Ptr{Float64}, Csize_t), state.handle, optr, length(obs), pid, pp, length(pp)) :
        ccall((:gsmc_step, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Csize_t, Cint, Ptr{Float64}, Csize_t), state.handle, optr, length(obs), pid, pp, length(pp)))
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
        y = isempty(state.observations[t]) ? sampled_observation(state, t)[i] : state.observations[t][1]
        cm[obs_address(state.model, t)] = state.model isa HMM ? Int(y) : y
    end
    cm
end

"Observation choices the device sampled at UNOBSERVED step t (current particle order)."
function sampled_observation(state::DeviceParticleFilterState, t::Int)
    out = Vector{Float64}(undef, state.num_particles)
    check(ccall((:gsmc_get_observation, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Csize_t), state.handle, t, out, length(out)), state.handle)
    out
end

"The whole loop `for t: maybe_resample!; particle_filter_step!` enqueued without a host round trip per step (gsmc_run_steps)."
function run_steps!(state::DeviceParticleFilterState, ys::Vector{Float64}; ess_threshold::Real = state.num_particles / 2, proposal = nothing)
    (pid, pp) = proposal_args(proposal)
    check(ccall((:gsmc_run_steps, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Csize_t, Csize_t, Cint, Ptr{Float64}, Csize_t, Float64),
                state.handle, ys, length(ys), 1, pid, pp, length(pp), Float64(ess_threshold)), state.handle)
    state.T += length(ys)
    append!(state.observations, [Float64[y] for y in ys])
    nothing
end

"Checkpoint / resume (gsmc_save / gsmc_restore): one file per handle (per rank of a sharded filter)."
save_checkpoint(state::DeviceParticleFilterState, path::String) =
    check(ccall((:gsmc_save, LIB), Cint, (Ptr{Cvoid}, Cstring), state.handle, path), state.handle)
function restore_checkpoint!(state::DeviceParticleFilterState, path::String, observations::Vector{Vector{Float64}})
    check(ccall((:gsmc_restore, LIB), Cint, (Ptr{Cvoid}, Cstring), state.handle, path), state.handle)
    state.T = length(observations); state.observations = observations
    nothing
end

# ---- multi-GPU: one Julia process per GPU (e.g. under MPI.jl); rank 0 makes the id and broadcasts its 128 bytes ----
mutable struct Communicator
    handle::Ptr{Cvoid}
    rank::Int
    nranks::Int
end
function comm_unique_id()
    id = Vector{UInt8}(undef, 128)
    check(ccall((:gsmc_comm_unique_id, LIB), Cint, (Ptr{UInt8}, Csize_t), id, 128))
    id
end
function Communicator(unique_id::Vector{UInt8}, rank::Int, nranks::Int; device::Int = -1)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:gsmc_comm_create, LIB), Cint, (Ptr{UInt8}, Csize_t, Cint, Cint, Cint, Ref{Ptr{Cvoid}}), unique_id, 128, rank, nranks, device, h))
    c = Communicator(h[], rank, nranks)
    finalizer(x -> ccall((:gsmc_comm_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.handle), c)
    c
end
"After attach the state owns particles [rank*N/R, (rank+1)*N/R); pass `comm = c` to initialize_particle_filter."
attach!(state::DeviceParticleFilterState, c::Communicator) =
    check(ccall((:gsmc_comm_attach, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), state.handle, c.handle), state.handle)

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

# importance.jl:20-33 and :35-52 for the importance-sampling families: gsmc_importance_sampling returns a handle whose
# log weights are already normalised and whose columns are the sampled latents
struct ISTraces{M<:DeviceISModel}
    handle::Ptr{Cvoid}
    model::M
    model_args::Tuple
    observations::ChoiceMap
    num_samples::Int
end
Base.length(t::ISTraces) = t.num_samples
function Base.getindex(t::ISTraces, i::Int)
    names = latent_names(t.model)
    dim = Ref{Cint}(0)
    check(ccall((:gsmc_state_dim, LIB), Cint, (Ptr{Cvoid}, Ref{Cint}), t.handle, dim), t.handle)
    row = Vector{Float64}(undef, Int(dim[]))
    idx = Int64[i - 1]
    check(ccall((:gsmc_get_trajectories, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Csize_t, Ptr{Float64}, Csize_t), t.handle, idx, 1, row, length(row)), t.handle)
    cm = choicemap()
    for (d, name) in enumerate(names)
        cm[name] = row[d]
    end
    if t.model isa OutlierRegression                     # flags are bit-packed, 32 per column, after the four reals
        for j in 1:length(t.model_args[1])
            cm[:data => j => :z] = ((UInt64(row[4 + ((j - 1) >> 5) + 1]) >> ((j - 1) & 31)) & 1) == 1
        end
    end
    merge(cm, t.observations)
end
function is_call(model::DeviceISModel, model_args::Tuple, observations::ChoiceMap, proposal, num_samples::Int; seed = 0, dtype = 0, device = -1)
    cfg = Ref(Config(sizeof(Config), model_id(model), dtype, 0, num_samples, seed, device, 0, 0, C_NULL))
    p = is_params(model, model_args)
    obs = is_observations(model, model_args, observations)
    (pid, pp) = proposal_args(proposal)
    lml = Ref{Float64}(0.0); h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:gsmc_importance_sampling, LIB), Cint,
                (Ref{Config}, Ptr{Float64}, Csize_t, Ptr{Float64}, Csize_t, Cint, Ptr{Float64}, Csize_t, Ref{Float64}, Ref{Ptr{Cvoid}}),
                cfg, p, length(p), obs, length(obs), pid, pp, length(pp), lml, h))
    lw = Vector{Float64}(undef, num_samples)
    check(ccall((:gsmc_get_log_weights, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Csize_t), h[], lw, length(lw)), h[])
    (ISTraces(h[], model, model_args, observations, num_samples), lw, lml[])
end
# importance.jl:20-33
importance_sampling(model::DeviceISModel, model_args::Tuple, observations::ChoiceMap, num_samples::Int, verbose = false; kwargs...) =
    is_call(model, model_args, observations, nothing, num_samples; kwargs...)
# importance.jl:35-52 (custom proposal: weight = model weight - proposal score)
importance_sampling(model::DeviceISModel, model_args::Tuple, observations::ChoiceMap, proposal::DeviceProposal, proposal_args::Tuple,
                    num_samples::Int, verbose = false; kwargs...) =
    is_call(model, model_args, observations, proposal, num_samples; kwargs...)

# importance.jl:70-87: sampling importance resampling that returns ONE trace. The reference keeps a reservoir of size
# one while it streams the samples; here the samples are generated `chunk_size` at a time on the device, the kept trace
# of a chunk is one categorical draw from the chunk's weights, and chunks are merged with the reference's rule
# (`bernoulli(exp(log_weight - log_total_weight))` with the chunk's total in the place of one sample's weight).
chunk_seed(seed::UInt64, c::Int) = seed + UInt64(c) * 0x9E3779B97F4A7C15
function importance_resampling(model::DeviceSSM, model_args::Tuple, observations::ChoiceMap, num_samples::Int;
                               verbose = false, seed::UInt64 = UInt64(0), chunk_size::Int = 1 << 24, kwargs...)   # verbose is a keyword (importance.jl:72)
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

"""
    validate(; num_particles = 10_000)

The reference's own particle-filter test (test/inference/particle_filter.jl:52-81,130-142) run twice: once with Gen's
dynamic-DSL model on the CPU, once with the device model; both estimates must be within the test's tolerance (0.01) of
the exact forward-algorithm value -4.87645083351704. The probe SURVEY.md section 8(b) asks for; needs Gen and a GPU.
"""
function validate(; num_particles::Int = 10_000)
    prior = [0.2, 0.3, 0.5]
    emission_dists = [0.1 0.2 0.7; 0.2 0.7 0.1; 0.7 0.2 0.1]'
    transition_dists = [0.4 0.4 0.2; 0.2 0.3 0.5; 0.9 0.05 0.05]'
    obs_x = [1, 1, 2, 3]
    expected = -4.87645083351704
    model = HMM(prior, Matrix(emission_dists), Matrix(transition_dists))
    state = initialize_particle_filter(model, (1,), choicemap((:x_init, obs_x[1])), num_particles)
    for T in 2:length(obs_x)
        maybe_resample!(state; ess_threshold = num_particles)
        particle_filter_step!(state, (T,), (UnknownChange(),), choicemap((:chain => (T - 1) => :x, obs_x[T])))
    end
    lml = log_ml_estimate(state)
    isapprox(lml, expected; atol = 0.02) || error("device estimate $lml is not within 0.02 of $expected")
    (device = lml, exact = expected)
end

export LinearGaussianSSM, HMM, StochasticVolatility, BearingsOnly, LinearRegression, NormalNormal, OutlierRegression, UniformNormal,
       DeviceProposal, DeviceParticleFilterState, Communicator, comm_unique_id, attach!, run_steps!, sampled_observation,
       save_checkpoint, restore_checkpoint!, validate

end # module
