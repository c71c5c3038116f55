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
    acc_seen::Float64
    ws_seen::Float64
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
    mc = MC(Ham, h[], ns, nw, get(params, :sweeps_per_call, 1), 0.0, 0.0)
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
    return nothing
end

function accumulators(mc::MC)
    out = zeros(Float64, 8)
    check(ccall((:kdsl_accumulators, libkdsl), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}),
                mc.handle, out, C_NULL, C_NULL))
    return out
end

# Carlo.sweep! (src/MonteCarlo.jl:538-607) for every walker
function Carlo.sweep!(mc::MC, ctx::MCContext)
    k = mc.sweeps_per_call
    check(ccall((:kdsl_set_sweeps, libkdsl), Cint, (Ptr{Cvoid}, Int64), mc.handle, ctx.sweeps * k))
    check(ccall((:kdsl_sweep, libkdsl), Cint, (Ptr{Cvoid}, Int64, Int64), mc.handle, k, -1))
    a = accumulators(mc)
    a[8] > 0 && throw(LinearAlgebra.SingularException(0))           # KDSL_ACC_N_SINGULAR
    dacc = a[2] - mc.acc_seen; dws = a[1] - mc.ws_seen
    mc.acc_seen = a[2]; mc.ws_seen = a[1]
    measure!(ctx, :acc, dws > 0 ? dacc / dws : 0.0)
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

export MC
end # module
