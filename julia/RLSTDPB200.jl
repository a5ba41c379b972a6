# RLSTDPB200.jl — thin `ccall` layer over libsnn_b200.so that keeps the reference's L1 API
# (Julia/Izhikevic_dastdp/Old_net/new_net.jl: Network, step!, input!; field access net.da, net.v,
#  net.s[pre,post]; Julia/Modules/Networks/net_arch.jl: step_LIF! for a population).
#
# NOT EXECUTED IN THE BUILD IMAGE (Julia is not installed there): the Python ctypes mirror
# rl_stdp_b200/_lib.py + network.py binds the identical symbols and is what the tests drive.
# Usage inside the reference repo:
#     include("/path/to/julia/RLSTDPB200.jl"); using .RLSTDPB200
#     n = RLSTDPB200.Network(1000, 0.1, 800; connection = conn)      # instead of Old_net/new_net.jl
#     spiked = step!(n, noise_row, train)                              # noise_row = 13 .* (rand(N) .- 0.5)
module RLSTDPB200

import Random

export Network, NetworkCol, step!, run_window!, add_dopamine!, reset_v!, weights, eligibility, get_synapses,
       set_param!, get_param, checkpoint, restore!,
       ActorPopulation, set_weights!, play_episodes,
       LayeredNetwork, step_LIF!, step_LIF_critic!, learn!, learn_critic!

const LIB = get(ENV, "SNN_B200_LIB", joinpath(@__DIR__, "..", "rl_stdp_b200", "libsnn_b200.so"))

# mirrors `struct snn_config` of include/snn_b200.h (field order and types must match)
mutable struct SnnConfig
    struct_size::Int32; dtype::Int32; mode::Int32; device::Int32
    n_neurons::Int64; n_per_instance::Int64; n_exc::Int64
    max_delay::Int32; threshold_ge::Int32
    vthresh::Float64; v_reset::Float64
    a_exc::Float64; a_inh::Float64; d_exc::Float64; d_inh::Float64
    trace_set::Float64; trace_decay::Float64; apost::Float64; apre::Float64
    da_decay::Float64; learn_base::Float64; eta::Float64
    wmin::Float64; wmax::Float64; sd_decay::Float64
    sd_decay_mode::Int32; rate_mode::Int32; noise_mode::Int32; plastic_ei::Int32
    noise_amp::Float64; noise_seed::UInt64
    theta::Float64; ttheta::Float64; ho_from::Int64
    ho_inh::Int32; variant_g::Int32
    reserved::NTuple{8,Int32}
    SnnConfig() = new()
end

struct SnnDaEvent
    step::Int32; instance::Int32; amount::Float64
end
struct SnnForceEvent
    step::Int32; neuron::Int32
end
mutable struct SnnStepStats
    n_spikes::Int64; n_syn_events::Int64; n_ltp_events::Int64
    device_ms::Float64; learn_ms::Float64; neuron_ms::Float64; synapse_ms::Float64
    learn_launches::Int32; neuron_launches::Int32; synapse_launches::Int32; reserved::Int32
    SnnStepStats() = new(0, 0, 0, 0.0, 0.0, 0.0, 0.0, 0, 0, 0, 0)
end

const FIELD_V, FIELD_U, FIELD_TRACE, FIELD_HO, FIELD_W, FIELD_SD, FIELD_DA, FIELD_I = Int32.(0:7)

function check(rc::Integer)
    rc == 0 && return
    msg = unsafe_string(ccall((:snn_last_error, LIB), Cstring, ()))
    error("libsnn_b200 error $rc: $msg")
end

"""
    Network(n_neurons, connectivity, n_exc; connection, dtype=:f64, mode=:exact)

Drop-in for `Network(n_neurons::Int64, connectivity::Float64, n_exc::Int64)`
(Old_net/new_net.jl:80-98).  `connection[pre,post]` defaults to `rand(N,N) .< connectivity`.
"""
mutable struct Network
    handle::Ptr{Cvoid}
    n_neurons::Int64
    n_exc::Int64
    connection::Matrix{Int64}
    rowptr::Vector{Int64}     # CSR of `connection` (0-based post ids), caller order of W / SD
    post::Vector{Int32}
    exc::Vector{Int64}
    inh::Vector{Int64}
end

function Network(n_neurons::Int64, connectivity::Float64, n_exc::Int64;
                 connection::AbstractMatrix = (rand(n_neurons, n_neurons) .< connectivity),
                 dtype::Symbol = :f64, mode::Symbol = :exact, device::Integer = 0)
    conn = Matrix{Int64}(connection)
    s = zeros(n_neurons, n_neurons)                        # new_net.jl:87-91
    s[1:n_exc, :] .= 1.0
    s[n_exc+1:end, :] .= -1.0
    s .*= conn
    cfg = SnnConfig()
    check(ccall((:snn_config_default, LIB), Int32, (Ref{SnnConfig},), cfg))
    cfg.dtype = dtype == :f64 ? 1 : 0
    cfg.mode = mode == :exact ? 1 : 0
    cfg.device = device
    cfg.n_neurons = n_neurons; cfg.n_per_instance = n_neurons; cfg.n_exc = n_exc
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:snn_create, LIB), Int32, (Ref{SnnConfig}, Ref{Ptr{Cvoid}}), cfg, h))
    # snn_set_synapses_dense takes exactly what variant A holds (column-major s, connection)
    check(ccall((:snn_set_synapses_dense, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Int64}), h[], s, conn))
    rowptr = zeros(Int64, n_neurons + 1); post = Int32[]
    for i in 1:n_neurons
        for j in 1:n_neurons
            conn[i, j] != 0 && push!(post, Int32(j - 1))
        end
        rowptr[i+1] = length(post)
    end
    net = Network(h[], n_neurons, n_exc, conn, rowptr, post, collect(1:n_exc), collect(n_exc+1:n_neurons))
    finalizer(n -> (n.handle != C_NULL && ccall((:snn_destroy, LIB), Int32, (Ptr{Cvoid},), n.handle); n.handle = C_NULL), net)
    return net
end

function get_array(net::Network, field::Int32, n::Integer)
    out = zeros(Float64, n)
    check(ccall((:snn_get_array, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}, Int64), net.handle, field, out, n))
    out
end
"""`(net.s[n1, syn], net.sd[n1, syn])` (newnet_DASTDP.jl:57-60) for a few synapses: `idx` are 0-based positions in the CSR
order the net was built with; a gather of `length(idx)` elements instead of two E-element copies."""
function get_synapses(net::Network, idx::Vector{Int64})
    w = zeros(Float64, length(idx)); sd = zeros(Float64, length(idx))
    check(ccall((:snn_get_synapses, LIB), Int32, (Ptr{Cvoid}, Ptr{Int64}, Int32, Ptr{Float64}, Ptr{Float64}),
                net.handle, idx, Int32(length(idx)), w, sd))
    (w, sd)
end
set_array!(net::Network, field::Int32, a::Vector{Float64}) =
    check(ccall((:snn_set_array, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}, Int64), net.handle, field, a, length(a)))

# `net.da`, `net.v`, `net.u`, `net.STDP`, `net.I` read through; `net.da = x` / `net.da += x` write through
function Base.getproperty(net::Network, f::Symbol)
    f === :da && return get_array(net, FIELD_DA, 1)[1]
    f === :v && return get_array(net, FIELD_V, getfield(net, :n_neurons))
    f === :u && return get_array(net, FIELD_U, getfield(net, :n_neurons))
    f === :STDP && return get_array(net, FIELD_TRACE, getfield(net, :n_neurons))
    f === :I && return get_array(net, FIELD_I, getfield(net, :n_neurons))
    f === :s && return weights(net)
    f === :sd && return eligibility(net)
    return getfield(net, f)
end
function Base.setproperty!(net::Network, f::Symbol, x)
    f === :da && return set_array!(net, FIELD_DA, [Float64(x)])
    f === :v && return set_array!(net, FIELD_V, Vector{Float64}(x))
    return setfield!(net, f, x)
end

"dense s[pre,post] (new_net.jl:15); write back with `set_weights!(net, s)` — `n.s[n1,n2] = 0.0` (newnet_DASTDP.jl:19)"
function dense_from_csr(net::Network, field::Int32)
    vals = get_array(net, field, length(getfield(net, :post)))
    N = getfield(net, :n_neurons); m = zeros(N, N)
    rp = getfield(net, :rowptr); po = getfield(net, :post)
    for i in 1:N, e in rp[i]+1:rp[i+1]
        m[i, po[e]+1] = vals[e]
    end
    m
end
weights(net::Network) = dense_from_csr(net, FIELD_W)
eligibility(net::Network) = dense_from_csr(net, FIELD_SD)
function set_weights!(net::Network, s::AbstractMatrix)
    rp = getfield(net, :rowptr); po = getfield(net, :post)
    vals = zeros(Float64, length(po))
    for i in 1:getfield(net, :n_neurons), e in rp[i]+1:rp[i+1]
        vals[e] = s[i, po[e]+1]
    end
    set_array!(net, FIELD_W, vals)
end

"""
    run_window!(net, noise, train; da_events, force) -> Vector{Vector{Int64}}

`noise` is N x T (column t = the recorded `13 .* (rand(N) .- 0.5)` of step t), `train` a Bool vector.
Runs T calls of step! on the device without a host round trip and returns the 1-based spike ids
of every step (ascending, like `findall`).
"""
function run_window!(net::Network, noise::Matrix{Float64}, train::AbstractVector{Bool};
                     da_events::Vector{SnnDaEvent} = SnnDaEvent[], force::Vector{SnnForceEvent} = SnnForceEvent[])
    T = length(train); N = getfield(net, :n_neurons)
    @assert size(noise) == (N, T)           # column-major N x T == row-major [T][N] of the C ABI
    flags = UInt8.(train)
    cap = max(1024, N * T)
    out = Vector{Int32}(undef, cap); offs = zeros(Int64, T + 1)
    st = SnnStepStats()
    GC.@preserve noise flags da_events force out offs begin
        check(ccall((:snn_step, LIB), Int32,
                    (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{UInt8}, Ptr{SnnForceEvent}, Int32, Ptr{SnnDaEvent}, Int32,
                     Ptr{Int32}, Ptr{Int64}, Int64, Ref{SnnStepStats}),
                    net.handle, T, noise, flags, force, length(force), da_events, length(da_events), out, offs, cap, st))
    end
    [Int64.(out[offs[k]+1:offs[k+1]]) .+ 1 for k in 1:T]
end

"step!(net, train) / step!(net, input_spikes, train)  (new_net.jl:146-161); `noise` replaces the internal rand"
function step!(net::Network, noise::Vector{Float64}, train::Bool = false)
    run_window!(net, reshape(noise, :, 1), [train])[1]
end
function step!(net::Network, input_spikes::Array{Int64}, noise::Vector{Float64}, train::Bool = false)
    force = [SnnForceEvent(0, Int32(j - 1)) for j in input_spikes]   # input! new_net.jl:104-106
    run_window!(net, reshape(noise, :, 1), [train]; force = force)[1]
end

add_dopamine!(net::Network, x::Real; instance::Integer = 0) =
    check(ccall((:snn_add_dopamine, LIB), Int32, (Ptr{Cvoid}, Int32, Float64), net.handle, instance, x))
reset_v!(net::Network) = check(ccall((:snn_reset_v, LIB), Int32, (Ptr{Cvoid},), net.handle))

# ---- named parameters (ABI 3): the model's scalars by their field name (`net.param["apre"] = ...` of the reference's
#      Dict-driven nets), scheduling / diagnostic switches ("persistent", "graph", "timeline", ...) — snn_b200.h
set_param!(net::Network, name::AbstractString, value::Real) =
    check(ccall((:snn_set_param, LIB), Int32, (Ptr{Cvoid}, Cstring, Float64), net.handle, name, value))
function get_param(net::Network, name::AbstractString)
    v = Ref{Float64}(0.0)
    check(ccall((:snn_get_param, LIB), Int32, (Ptr{Cvoid}, Cstring, Ref{Float64}), net.handle, name, v))
    v[]
end

# ---- checkpoint / resume with the axonal delays in flight (the reference only saves genomes, save_param.jl:4-47)
function checkpoint(net::Network)
    n = Ref{Int64}(0)
    check(ccall((:snn_checkpoint_bytes, LIB), Int32, (Ptr{Cvoid}, Ref{Int64}), net.handle, n))
    blob = zeros(UInt8, n[])
    check(ccall((:snn_checkpoint, LIB), Int32, (Ptr{Cvoid}, Ptr{UInt8}, Int64), net.handle, blob, n[]))
    blob
end
restore!(net::Network, blob::Vector{UInt8}) =
    check(ccall((:snn_restore, LIB), Int32, (Ptr{Cvoid}, Ptr{UInt8}, Int64), net.handle, blob, length(blob)))

"""
    NetworkCol(n_neurons, connectivity, n_exc; connection)

Variant B, `Network(n_neurons, connectivity, n_exc)` of Old_net/new_net_col.jl:86-103: `connection`, `s`, `sd` are
held TRANSPOSED ([post, pre], "read by columns", :13-15).  `connection` is what `connect()` draws ([pre, post], :28);
the constructor transposes it (:94) and hands the matrices to the library exactly as that struct holds them.
Everything else (`step!`, `run_window!`, `net.da`, ...) is the `Network` API; `weights(net)` returns s[pre, post].
"""
function NetworkCol(n_neurons::Int64, connectivity::Float64, n_exc::Int64;
                    connection::AbstractMatrix = (rand(n_neurons, n_neurons) .< connectivity),
                    dtype::Symbol = :f64, mode::Symbol = :exact, device::Integer = 0)
    conn = Matrix{Int64}(connection)
    conn_t = Matrix{Int64}(transpose(conn))               # new_net_col.jl:94
    s = zeros(n_neurons, n_neurons)                        # :95-99
    s[:, 1:n_exc] .= 1.0
    s[:, n_exc+1:end] .= -1.0
    s .*= conn_t
    cfg = SnnConfig()
    check(ccall((:snn_config_default, LIB), Int32, (Ref{SnnConfig},), cfg))
    cfg.dtype = dtype == :f64 ? 1 : 0
    cfg.mode = mode == :exact ? 1 : 0
    cfg.device = device
    cfg.n_neurons = n_neurons; cfg.n_per_instance = n_neurons; cfg.n_exc = n_exc
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:snn_create, LIB), Int32, (Ref{SnnConfig}, Ref{Ptr{Cvoid}}), cfg, h))
    check(ccall((:snn_set_synapses_dense_col, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Int64}), h[], s, conn_t))
    rowptr = zeros(Int64, n_neurons + 1); post = Int32[]
    for i in 1:n_neurons
        for j in 1:n_neurons
            conn[i, j] != 0 && push!(post, Int32(j - 1))
        end
        rowptr[i+1] = length(post)
    end
    net = Network(h[], n_neurons, n_exc, conn, rowptr, post, collect(1:n_exc), collect(n_exc+1:n_neurons))
    finalizer(n -> (n.handle != C_NULL && ccall((:snn_destroy, LIB), Int32, (Ptr{Cvoid},), n.handle); n.handle = C_NULL), net)
    return net
end

# ---- one network over several GPUs (one Julia process per GPU; include/snn_b200.h snn_partition_*)
const PEER_BLOB_BYTES = 192

"Integrate only neurons lo:hi (1-based, inclusive; multiples of 128 except the last bound) on this handle."
set_partition!(net::Network, lo::Integer, hi::Integer) =
    check(ccall((:snn_set_partition, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Ptr{Cvoid}, Ptr{Cvoid}),
                net.handle, lo - 1, hi, C_NULL, C_NULL))

"This rank's exchange-ring descriptor (a CUDA IPC handle); gather one per rank, e.g. with MPI.Allgather."
function partition_export(net::Network)
    blob = zeros(UInt8, PEER_BLOB_BYTES)
    check(ccall((:snn_partition_export, LIB), Int32, (Ptr{Cvoid}, Ptr{UInt8}), net.handle, blob))
    return blob
end

"`blobs` = the descriptors of all ranks concatenated in rank order; call a barrier afterwards, then step."
partition_connect!(net::Network, rank::Integer, world::Integer, blobs::Vector{UInt8}) =
    check(ccall((:snn_partition_connect, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{UInt8}),
                net.handle, rank, world, blobs))

# ---- LIF actor populations (Modules/Networks/net_arch.jl variant F; Actors/cartpole_fitness.jl,
#      Actors/mountcar_fitness.jl): one handle = n_agents networks of the same `arch`

# mirrors `struct snn_actor_config` of include/snn_b200.h
mutable struct SnnActorConfig
    struct_size::Int32; device::Int32; n_agents::Int32; n_layers::Int32
    arch::NTuple{8,Int32}
    vthresh::Float64; c::Float64
    wei::Float64; wie::Float64; wmax::Float64
    theta::Float64; ttheta::Float64
    imin::Float64
    apost::Float64; apre::Float64
    tde::Float64; ttrace::Float64
    taulif::Float64
    eta::Float64; winh::Float64
    da_inc::Float64
    reserved::NTuple{8,Int32}
    SnnActorConfig() = new()
end

const ENV_CARTPOLE, ENV_MOUNTAINCAR = Int32(0), Int32(1)

mutable struct ActorPopulation
    handle::Ptr{Cvoid}
    arch::Vector{Int64}
    n_agents::Int64
end

"`ActorPopulation(arch, params, n)` stands for `[Network(arch, params) for _ in 1:n]` (net_arch.jl:98-128)."
function ActorPopulation(arch::Vector{Int64}, params::Dict, n_agents::Integer; device::Integer = 0)
    cfg = SnnActorConfig()
    check(ccall((:snn_actor_config_default, LIB), Int32, (Ref{SnnActorConfig},), cfg))
    cfg.device = device; cfg.n_agents = n_agents; cfg.n_layers = length(arch)
    cfg.arch = ntuple(i -> i <= length(arch) ? Int32(arch[i]) : Int32(0), 8)
    for (k, f) in (("vthresh", :vthresh), ("c", :c), ("wei", :wei), ("wie", :wie), ("wmax", :wmax),
                   ("theta", :theta), ("ttheta", :ttheta), ("imin", :imin), ("apost", :apost), ("apre", :apre),
                   ("tde", :tde), ("ttrace", :ttrace), ("taulif", :taulif), ("eta", :eta), ("da_inc", :da_inc))
        haskey(params, k) && setfield!(cfg, f, Float64(params[k]))
    end
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:snn_actor_create, LIB), Int32, (Ref{SnnActorConfig}, Ref{Ptr{Cvoid}}), cfg, h))
    pop = ActorPopulation(h[], arch, n_agents)
    finalizer(p -> (p.handle != C_NULL && ccall((:snn_actor_destroy, LIB), Int32, (Ptr{Cvoid},), p.handle); p.handle = C_NULL), pop)
    return pop
end

"`e[:, :, a]` = the `net.s.e[layer]` of agent a (arch[layer] x arch[layer+1], as set_weight! fills it)."
function set_weights!(pop::ActorPopulation, layer::Integer, e::Array{Float64,3})
    @assert size(e) == (pop.arch[layer], pop.arch[layer+1], pop.n_agents)   # column-major per agent == the C layout
    check(ccall((:snn_actor_set_array, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Int64),
                pop.handle, FIELD_W, layer - 1, e, length(e)))
end

"""
    play_episodes(pop, matalpha, state0; env, tie_picks, n_steps = 200) -> (rewards, n_decisions, n_rand)

`play_episode` (cartpole_fitness.jl:51-114 / mountcar_fitness.jl:48-107) for every agent at once — the body of
the `fitness(i::Individual)` callback for a whole Cambrian population.  `matalpha[:, :, a]` is arch[1] x 4
(CartPole) or arch[1] x 2 (MountainCar); `state0[:, a]` the `env.reset()` observation; `tie_picks[:, a]` the
`rand(rng_action, ...)` draws as 0-based picks (default: what `MersenneTwister(0)` gives).
"""
function play_episodes(pop::ActorPopulation, matalpha::Array{Float64,3}, state0::Matrix{Float64};
                       env::Int32 = ENV_CARTPOLE, n_steps::Integer = 200, win_size::Integer = 40,
                       tie_picks::Matrix{Int32} = default_tie_picks(env, n_steps, pop.n_agents))
    A = pop.n_agents
    s0 = zeros(4, A); s0[1:size(state0, 1), :] .= state0
    rewards = zeros(A); nd = zeros(Int32, A); nr = zeros(Int32, A); so = zeros(4, A); ms = Ref{Float64}(0.0)
    GC.@preserve matalpha s0 tie_picks begin
        check(ccall((:snn_actor_episodes_env, LIB), Int32,
                    (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Int32, Int32, Int32, Int32,
                     Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ref{Float64}),
                    pop.handle, env, matalpha, s0, tie_picks, size(tie_picks, 1), n_steps, win_size, 0,
                    rewards, nd, nr, so, ms))
    end
    return rewards, nd, nr
end

# ---- the layered LIF network as the reference's scripts hold it: ONE `Network(arch, param)` (net_arch.jl:98-128)
#      with step_LIF! / step_LIF_critic! / learn! / learn_critic! of net_arch.jl:213-272 — a population of one.

"`LayeredNetwork(arch, param)` stands for `Network(arch::Array{Int64}, param::Dict)` (net_arch.jl:98-128)."
mutable struct LayeredNetwork
    pop::ActorPopulation
    arch::Vector{Int64}
    param::Dict
    n_exc::Int64
    n_neurons::Int64
end
function LayeredNetwork(arch::Vector{Int64}, param::Dict = Dict(); device::Integer = 0)
    pop = ActorPopulation(arch, param, 1; device = device)
    n_exc = sum(arch)
    LayeredNetwork(pop, arch, param, n_exc, n_exc + sum(arch[2:end-1]))
end

# one call of step_LIF!(net, input_spikes, train): `input_spikes` lists every input neuron once, as int_spikes
# (PoissonStateSpike.jl:62-68) builds it — the form every script of the reference passes
function lif_step(net::LayeredNetwork, input_spikes::Array{Tuple{Int64,Float64}}, train::Bool, critic::Bool)
    a1 = net.arch[1]
    @assert sort([t[1] for t in input_spikes]) == collect(1:a1) "input_spikes must list the neurons 1:arch[1] once each"
    inten = zeros(Float64, a1)
    for (i, x) in input_spikes
        inten[i] = x
    end
    mask = zeros(Int32, 2)            # bit j of the 64-bit mask = neuron j+1 spiked
    ms = Ref{Float64}(0.0)
    GC.@preserve inten mask begin
        check(ccall((:snn_actor_window, LIB), Int32,
                    (Ptr{Cvoid}, Int32, Ptr{Float64}, Int32, Int32, Int32, Ptr{Int32}, Ptr{Int32}, Ref{Float64}),
                    net.pop.handle, 1, inten, train ? 1 : 0, critic ? 1 : 0, 0, C_NULL, mask, ms))
    end
    bits = reinterpret(UInt32, mask[1])
    [j + 1 for j in 0:min(31, net.n_neurons - 1) if (bits >> j) & 0x1 == 0x1]
end

"step_LIF!(net, input_spikes, train)  (net_arch.jl:256-263) -> spiked (1-based ids)"
step_LIF!(net::LayeredNetwork, input_spikes::Array{Tuple{Int64,Float64}}, train::Bool = false) =
    lif_step(net, input_spikes, train, false)
"step_LIF_critic!(net, input_spikes, train)  (net_arch.jl:265-272)"
step_LIF_critic!(net::LayeredNetwork, input_spikes::Array{Tuple{Int64,Float64}}, train::Bool = false) =
    lif_step(net, input_spikes, train, true)
"learn!(net)  (net_arch.jl:213-220): e[l] += da * de[l], clamped to [0, wmax]"
learn!(net::LayeredNetwork) = check(ccall((:snn_actor_learn, LIB), Int32, (Ptr{Cvoid}, Int32), net.pop.handle, 0))
"learn_critic!(net)  (net_arch.jl:222-235)"
learn_critic!(net::LayeredNetwork) = check(ccall((:snn_actor_learn, LIB), Int32, (Ptr{Cvoid}, Int32), net.pop.handle, 1))

# `net.da`, `net.v`, `net.ho`, `net.trace_e` read through; `net.da = x` / `net.da += x`, `net.v .= c` write through
function actor_array(net::LayeredNetwork, field::Int32, n::Integer)
    out = zeros(Float64, n)
    check(ccall((:snn_actor_get_array, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Int64),
                getfield(net, :pop).handle, field, 0, out, n))
    out
end
function Base.getproperty(net::LayeredNetwork, f::Symbol)
    f === :da && return actor_array(net, FIELD_DA, 1)[1]
    f === :v && return actor_array(net, FIELD_V, getfield(net, :n_neurons))
    f === :ho && return actor_array(net, FIELD_HO, getfield(net, :n_neurons))
    f === :trace_e && return actor_array(net, FIELD_TRACE, getfield(net, :n_exc))
    return getfield(net, f)
end
function Base.setproperty!(net::LayeredNetwork, f::Symbol, x)
    if f === :da || f === :v
        a = f === :da ? [Float64(x)] : Vector{Float64}(x)
        return check(ccall((:snn_actor_set_array, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Int64),
                           getfield(net, :pop).handle, f === :da ? FIELD_DA : FIELD_V, 0, a, length(a)))
    end
    return setfield!(net, f, x)
end

function default_tie_picks(env::Int32, n_steps::Integer, n_agents::Integer)
    k = env == ENV_CARTPOLE ? n_steps : 3 * n_steps
    rng = Random.MersenneTwister(0)
    picks = Int32[rand(rng, 0:1) for _ in 1:k]            # cartpole_fitness.jl:55,89; a fresh generator per episode
    return repeat(picks, 1, n_agents)
end

end # module
