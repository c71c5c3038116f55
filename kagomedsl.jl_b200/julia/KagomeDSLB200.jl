# KagomeDSLB200.jl -- thin `ccall` layer over libkdsl.so that keeps KagomeDSL.jl's Carlo interface.
#
# WRITTEN BLIND: Julia is not available in the build image, so this file has never been executed.
# It is pure marshalling (no arithmetic): every hot-path operation is one C-ABI call declared in
# include/kdsl.h.  Host-side setup (DoubleKagome / Hamiltonian / init_conf_qr!) is taken from the
# reference package itself, which stays a dependency for its cold path.
module KagomeDSLB200

using Carlo
using HDF5
using Random
using LinearAlgebra
import KagomeDSL                     # reference package: lattice, Hamiltonian, QR start configuration

const libkdsl = get(ENV, "KDSL_LIB", joinpath(@__DIR__, "..", "csrc", "libkdsl.so"))

struct KdslError <: Exception
    code::Cint
    msg::String
end

function check(rc::Cint)
    rc >= 0 && return rc
    msg = unsafe_string(ccall((:kdsl_last_error, libkdsl), Cstring, ()))
    rc == -3 && throw(LinearAlgebra.SingularException(0))          # KDSL_ERR_SINGULAR
    rc == -1 && throw(ArgumentError(msg))                          # KDSL_ERR_INVALID_ARGUMENT
    throw(KdslError(rc, msg))
end

"""
    MC(params) -- same keys as KagomeDSL.MC (src/MonteCarlo.jl:165-185) plus
    :n_walkers (default 1), :device (default 0), :sweeps_per_call (default 1).
"""
mutable struct MC <: AbstractMC
    Ham::KagomeDSL.Hamiltonian
    handle::Ptr{Cvoid}
    ns::Int
    n_walkers::Int
    sweeps_per_call::Int
    acc_seen::Vector{Int64}          # per-walker accepted-move counts already reported as :acc
end

function MC(params::AbstractDict)
    ref = KagomeDSL.MC(params)                                      # builds lattice + Hamiltonian (cold path)
    Ham = ref.Ham
    ns = length(ref.kappa_up)
    nw = get(params, :n_walkers, 1)
    dev = get(params, :device, 0)
    bonds = Matrix{Int32}(undef, 2, length(Ham.nn))
    for (b, (i, j)) in enumerate(Ham.nn); bonds[1, b] = i; bonds[2, b] = j; end
    h = Ref{Ptr{Cvoid}}(C_NULL)
    if maximum(abs.(imag.(Ham.U_up))) == 0 && maximum(abs.(imag.(Ham.U_down))) == 0
        Uu = Matrix{Float64}(real.(Ham.U_up)); Ud = Matrix{Float64}(real.(Ham.U_down))
        check(ccall((:kdsl_create, libkdsl), Cint,
                    (Ref{Ptr{Cvoid}}, Cint, Cint, Cint, Cint, Cint, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Cint),
                    h, dev, ns, Ham.N_up, Ham.N_down, length(Ham.nn), bonds, Uu, Ud, nw))
    else    # Peierls flux B != 0: Matrix{ComplexF64} is exactly the interleaved (re, im) layout of kdsl_create_c128
        Uu = Matrix{ComplexF64}(Ham.U_up); Ud = Matrix{ComplexF64}(Ham.U_down)
        check(ccall((:kdsl_create_c128, libkdsl), Cint,
                    (Ref{Ptr{Cvoid}}, Cint, Cint, Cint, Cint, Cint, Ptr{Int32}, Ptr{ComplexF64}, Ptr{ComplexF64}, Cint),
                    h, dev, ns, Ham.N_up, Ham.N_down, length(Ham.nn), bonds, Uu, Ud, nw))
    end
    mc = MC(Ham, h[], ns, nw, get(params, :sweeps_per_call, 1), zeros(Int64, nw))
    finalizer(m -> ccall((:kdsl_destroy, libkdsl), Cint, (Ptr{Cvoid},), m.handle), mc)
    return mc
end

function load_configuration!(mc::MC, kappa_up::AbstractMatrix{<:Integer}, kappa_down::AbstractMatrix{<:Integer})
    ku = Matrix{Int32}(kappa_up); kd = Matrix{Int32}(kappa_down)     # ns x n_walkers (column = walker)
    check(ccall((:kdsl_set_config, libkdsl), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}), mc.handle, ku, kd))
    nsing = Ref{Cint}(0)
    check(ccall((:kdsl_refresh, libkdsl), Cint, (Ptr{Cvoid}, Ref{Cint}), mc.handle, nsing))
end

# Carlo.init! (src/MonteCarlo.jl:435-441): QR start state for every walker, walker RNG streams split off ctx.rng
function Carlo.init!(mc::MC, ctx::MCContext, params::AbstractDict)
    ref = KagomeDSL.MC(params)
    KagomeDSL.init_conf_qr!(ref, mc.ns, params[:N_up])
    states = rand(ctx.rng, UInt64, 4, mc.n_walkers)
    check(ccall((:kdsl_set_rng, libkdsl), Cint, (Ptr{Cvoid}, Ptr{UInt64}), mc.handle, states))
    try
        load_configuration!(mc, repeat(ref.kappa_up, 1, mc.n_walkers), repeat(ref.kappa_down, 1, mc.n_walkers))
    catch e
        e isa LinearAlgebra.SingularException &&
            error("QR-based configuration is singular. The Hamiltonian may be rank-deficient.")
        rethrow(e)
    end
    check(ccall((:kdsl_reset_accumulators, libkdsl), Cint, (Ptr{Cvoid},), mc.handle))
    fill!(mc.acc_seen, 0)
    return nothing
end

"sums over this handle's walkers (KDSL_ACC_* order) and the per-walker accepted-move counts"
function accumulators(mc::MC)
    out = zeros(Float64, 8)
    acc_w = zeros(Int64, mc.n_walkers)
    check(ccall((:kdsl_accumulators, libkdsl), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}),
                mc.handle, out, acc_w, C_NULL))
    return out, acc_w
end

# Carlo.sweep! (src/MonteCarlo.jl:538-607) for every walker
function Carlo.sweep!(mc::MC, ctx::MCContext)
    k = mc.sweeps_per_call
    check(ccall((:kdsl_set_sweeps, libkdsl), Cint, (Ptr{Cvoid}, Int64), mc.handle, ctx.sweeps * k))
    check(ccall((:kdsl_sweep, libkdsl), Cint, (Ptr{Cvoid}, Int64, Int64), mc.handle, k, -1))
    a, acc_w = accumulators(mc)
    a[8] > 0 && throw(LinearAlgebra.SingularException(0))           # KDSL_ACC_N_SINGULAR
    # :acc (src/MonteCarlo.jl:548-589): the reference records 0.0 / 1.0 for its one walker; a batch records the vector of
    # the walkers' acceptance fractions over this call (a Carlo vector observable: one chain per component)
    d = (acc_w .- mc.acc_seen) ./ k
    mc.acc_seen .= acc_w
    measure!(ctx, :acc, mc.n_walkers == 1 ? d[1] : d)
    return nothing
end

# Carlo.measure! (src/MonteCarlo.jl:628-634): vector observable, one O_L per walker
function Carlo.measure!(mc::MC, ctx::MCContext)
    n_occ = min(mc.Ham.N_up, mc.Ham.N_down)
    if (ctx.sweeps * mc.sweeps_per_call) % n_occ == 0
        ol = zeros(Float64, mc.n_walkers)
        check(ccall((:kdsl_measure, libkdsl), Cint, (Ptr{Cvoid}, Ptr{Float64}), mc.handle, ol))
        measure!(ctx, :OL, mc.n_walkers == 1 ? ol[1] : ol)
    end
    return nothing
end

# src/MonteCarlo.jl:675-687
function Carlo.register_evaluables(::Type{MC}, eval::Evaluator, params::AbstractDict)
    ns = params[:n1] * params[:n2] * 3
    evaluate!(eval, :energy, (:OL,)) do OL
        return OL / ns
    end
    return nothing
end

# src/MonteCarlo.jl:715-719 (+ the device-owned RNG streams)
function Carlo.write_checkpoint(mc::MC, out::HDF5.Group)
    ku = Matrix{Int32}(undef, mc.ns, mc.n_walkers); kd = similar(ku)
    check(ccall((:kdsl_get_config, libkdsl), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}), mc.handle, ku, kd))
    st = Matrix{UInt64}(undef, 4, mc.n_walkers)
    check(ccall((:kdsl_get_rng, libkdsl), Cint, (Ptr{Cvoid}, Ptr{UInt64}), mc.handle, st))
    out["kappa_up"] = mc.n_walkers == 1 ? Vector{Int}(ku[:, 1]) : Matrix{Int}(ku)
    out["kappa_down"] = mc.n_walkers == 1 ? Vector{Int}(kd[:, 1]) : Matrix{Int}(kd)
    out["rng_state"] = st
    return nothing
end

# src/MonteCarlo.jl:751-755; W is recomputed (the reference leaves it zero until the next refresh)
function Carlo.read_checkpoint!(mc::MC, in::HDF5.Group)
    ku = read(in, "kappa_up"); kd = read(in, "kappa_down")
    ku = ku isa AbstractVector ? reshape(ku, :, 1) : ku
    kd = kd isa AbstractVector ? reshape(kd, :, 1) : kd
    if haskey(in, "rng_state")
        st = Matrix{UInt64}(read(in, "rng_state"))
        check(ccall((:kdsl_set_rng, libkdsl), Cint, (Ptr{Cvoid}, Ptr{UInt64}), mc.handle, st))
    end
    load_configuration!(mc, ku, kd)
    return nothing
end

# ---- multi-GPU in one Julia process: one MC (handle) per device, NCCL sum of the accumulators inside libkdsl ----
"ncclCommInitAll over the handles' devices (kdsl_comm_init_all)"
function comm_init_all(mcs::Vector{MC})
    hs = Ptr{Cvoid}[m.handle for m in mcs]
    check(ccall((:kdsl_comm_init_all, libkdsl), Cint, (Cint, Ptr{Ptr{Cvoid}}), length(hs), hs))
end

"global sums (KDSL_ACC_* order) over all handles: one grouped ncclAllReduce of 8 doubles"
function accumulators_allreduce(mcs::Vector{MC})
    hs = Ptr{Cvoid}[m.handle for m in mcs]
    out = zeros(Float64, 8)
    check(ccall((:kdsl_group_accumulators_allreduce, libkdsl), Cint, (Cint, Ptr{Ptr{Cvoid}}, Ptr{Float64}), length(hs), hs, out))
    return out
end

# one process per GPU (Carlo's MPI ranks): rank 0 calls comm_unique_id(), ships the 128 bytes (MPI.Bcast!), all call comm_init_rank!
function comm_unique_id()
    id = zeros(UInt8, 128)
    check(ccall((:kdsl_comm_unique_id, libkdsl), Cint, (Ptr{UInt8},), id))
    return id
end
comm_init_rank!(mc::MC, n_ranks::Integer, rank::Integer, id::Vector{UInt8}) =
    check(ccall((:kdsl_comm_init_rank, libkdsl), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), mc.handle, n_ranks, rank, id))
function accumulators_allreduce(mc::MC)
    out = zeros(Float64, 8)
    check(ccall((:kdsl_accumulators_allreduce, libkdsl), Cint, (Ptr{Cvoid}, Ptr{Float64}), mc.handle, out))
    return out
end

# ---- extra observables (structure factor at the wave vectors qs [2 x nq], Z_mu-reweighted |psi|^2 averages) ----
function set_observables!(mc::MC, qs::AbstractMatrix{<:Real}, coords::AbstractMatrix{<:Real})   # coords: 2 x ns
    ph = transpose(coords) * qs                                     # ns x nq, column q contiguous = [nq][ns] in C
    c = Matrix{Float64}(cos.(ph)); s = Matrix{Float64}(sin.(ph))
    check(ccall((:kdsl_set_observables, libkdsl), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Float64}), mc.handle, size(qs, 2), c, s))
end
function observables(mc::MC, nq::Integer; allreduce::Bool = false)
    out = zeros(Float64, 4 + 2nq)
    check(ccall((:kdsl_get_observables, libkdsl), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint), mc.handle, out, allreduce))
    return out
end

export MC
end # module
