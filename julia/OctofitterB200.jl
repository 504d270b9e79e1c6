# OctofitterB200.jl — Julia glue over libocto_b200.so (include/octo_b200.h), Octofitter v8.3.0.
#
# STATUS: written against the C ABI and the reference sources (file:line citations below); NOT executed in the build
# environment (no julia binary in the image or on the GPU box).  What keeps it honest here: tests/test_julia_glue_cpu.py
# parses the struct definitions of this file and checks their layout against the C header, and
# tests/c_abi_julia_replay.c replays this file's call sequence through the ABI on the GPU.
#
# HOW IT DROPS IN.  The path is put behind the reference's own plugin API for likelihoods — a subtype of
# `Octofitter.AbstractObs` with an `ln_like(obs, ctx)` method (src/variables.jl:87-134, docs/src/custom-likelihood.md):
#
#   * every observation the kernel supports (PlanetRelAstromObs incl. the ObsPriorAstromONeil2019 wrapper,
#     StarAbsoluteRVObs / MarginalizedStarAbsoluteRVObs / PlanetRelativeRVObs without GP and with a trend_function that is
#     zero or linear in the observation variables (probed),
#     HGCAInstantaneousObs; planets on Visual{KepOrbit} or ThieleInnesOrbit) is replaced by a `BlankLikelihood` that keeps
#     its priors and derived variables — exactly what `prior_only_model` does (src/cross-validation.jl:60-99) — so D, the
#     parameter order and arr2nt are unchanged;
#   * ONE system-level `B200Likelihood` is appended.  Its `ln_like` reads the kernel inputs out of `ctx.θ_system` by
#     path and calls the library: `octo_logp` for Float64 parameters; for ForwardDiff duals (the reference differentiates
#     ℓπcallback with chunk = D duals, src/logdensitymodel.jl:43-45,169-177) ONE `octo_logp_grad` call and a dual whose
#     partials are Σ_k g_k · partials(input_k) — the chain rule through priors, bijectors and arbitrary `Derived`
#     expressions stays ForwardDiff's, the epoch loop and its gradient are the GPU's.
#
# `B200Model(system)` then is `Octofitter.LogDensityModel(b200_system(system))`: a GENUINE LogDensityModel.  `octofit`,
# `octofit_pigeons`, `Pigeons.initialization / sample_iid! / default_reference / default_explorer`
# (ext/OctofitterPigeonsExt/OctofitterPigeonsExt.jl:10-72), `prior_only_model`, the initialisers and every other
# function typed `::LogDensityModel` (src/sampling.jl:117,149,319) take it unchanged.  The tempering reference that
# `default_reference` builds blanks the B200Likelihood like any other data term (`_isprior` is false).
#
# BATCHED / DEVICE-RESIDENT use is a separate object, `DevicePosterior(model)`: when every variable of the model belongs
# to the standard families (Normal, Uniform, LogUniform, Sine, truncated Normal priors; UniformCircular; constants;
# tp = θ_at_epoch_to_tperi(…)) the whole log posterior — invlink, logpdf_with_trans, UnitLengthPrior, derived inputs,
# likelihood — is evaluated on the device for a D x N matrix of unconstrained vectors in one launch
# (`octo_set_parameterization` + `octo_logpost_grad`), and the chain-batched explorers run there too (`octo_hmc_run`,
# `octo_pt_hmc_run`, `octo_pt_hmc_run_dist`).  `DevicePosterior` returns `nothing` for any other model.
module OctofitterB200

using Octofitter, PlanetOrbits, ForwardDiff, LogDensityProblems, Distributions, Random
import Octofitter: Planet, System, BlankLikelihood, Priors, Derived, normalizename, likelihoodname

const LIB = get(ENV, "OCTO_B200_LIB", joinpath(@__DIR__, "..", "octofitter.jl_b200", "lib", "libocto_b200.so"))
const MAXP = 4
const OCTO_ABI_VERSION = 4

# ------------------------------------------------------------------------------------------------------------------
# ABI structs (field for field as include/octo_b200.h; layout checked by tests/test_julia_glue_cpu.py)
# ------------------------------------------------------------------------------------------------------------------
struct OctoConstants
    kepler_year_days::Cdouble
    year2day::Cdouble
    rad2as::Cdouble
    pc2au::Cdouble
    au2m::Cdouble
    sec2year::Cdouble
    mjup2msol::Cdouble
end
# SURVEY.md R1: the PlanetOrbits source is not part of the Octofitter tree — never hard-code, inject its live values
OctoConstants() = OctoConstants(PlanetOrbits.kepler_year_to_julian_day_conversion_factor, PlanetOrbits.year2day_julian,
    PlanetOrbits.rad2as, PlanetOrbits.pc2au, PlanetOrbits.au2m, PlanetOrbits.sec2year_julian, PlanetOrbits.mjup2msol_IAU)

struct OctoObsBlock
    kind::Int32
    planet::Int32
    n_epochs::Int32
    has_cor::Int32
    epoch::Ptr{Cdouble}
    y1::Ptr{Cdouble}
    y2::Ptr{Cdouble}
    s1::Ptr{Cdouble}
    s2::Ptr{Cdouble}
    cor::Ptr{Cdouble}
    idx_jitter::Int32
    idx_platescale::Int32
    idx_northangle::Int32
    idx_offset::Int32
    obs_prior::Int32
    idx_pmra::Int32
    idx_pmdec::Int32
    reserved::Int32
    aux::Ptr{Cdouble}
    n_trend::Int32
    idx_trend::NTuple{3,Int32}
    trend_basis::Ptr{Cdouble}
    trend_const::Ptr{Cdouble}
end

struct OctoLayout
    n_planets::Int32
    n_in::Int32
    idx_plx::NTuple{MAXP,Int32}
    idx_a::NTuple{MAXP,Int32}
    idx_e::NTuple{MAXP,Int32}
    idx_i::NTuple{MAXP,Int32}
    idx_w::NTuple{MAXP,Int32}
    idx_W::NTuple{MAXP,Int32}
    idx_tp::NTuple{MAXP,Int32}
    idx_M::NTuple{MAXP,Int32}
    idx_mass::NTuple{MAXP,Int32}
    basis::NTuple{MAXP,Int32}
    idx_A::NTuple{MAXP,Int32}
    idx_B::NTuple{MAXP,Int32}
    idx_F::NTuple{MAXP,Int32}
    idx_G::NTuple{MAXP,Int32}
end

struct OctoPrior
    family::Int32
    reserved::Int32
    p::NTuple{4,Cdouble}
end

struct OctoInputDef
    op::Int32
    a::NTuple{8,Int32}
    value::Cdouble
end

const PRIOR_NORMAL, PRIOR_UNIFORM, PRIOR_LOGUNIFORM, PRIOR_SINE, PRIOR_TRUNCNORMAL = Int32(0), Int32(1), Int32(2), Int32(3), Int32(4)
const IN_PARAM, IN_CONST, IN_CIRC, IN_TPERI, IN_TPERI_TI = Int32(0), Int32(1), Int32(2), Int32(3), Int32(4)
const KIND_HGCA = Int32(5)

octo_error() = unsafe_string(ccall((:octo_last_error, LIB), Cstring, ()))
check(rc) = rc == 0 ? nothing : error("libocto_b200: $(octo_error())")

function __init__()
    v = ccall((:octo_abi_version, LIB), Cint, ())
    v == OCTO_ABI_VERSION || error("libocto_b200 ABI version $v, this glue was written for $OCTO_ABI_VERSION")
end

# the context is owned by a mutable holder so that a finalizer can destroy it
mutable struct Context
    ptr::Ptr{Cvoid}
    keep::Vector{Any}            # table columns, alive until octo_create copied them (and for good measure after)
    function Context(ptr, keep)
        c = new(ptr, keep)
        finalizer(x -> (x.ptr == C_NULL || ccall((:octo_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.ptr); x.ptr = C_NULL), c)
        return c
    end
end

# ------------------------------------------------------------------------------------------------------------------
# What can be offloaded
# ------------------------------------------------------------------------------------------------------------------
"""
    probe_trend(obs, θ_obs) -> (names, basis, constant) | nothing

`trend_function(θ_obs, epoch)` of an RV observation (rv-absolute.jl:69,143; rv-absolute-margin.jl:52,111;
rv-relative.jl:64,131) is an arbitrary closure; the default one cannot be recognised by identity.  It is PROBED: the
constant part b₀(t) = f(0, t), one basis b_v(t) = f(e_v, t) - b₀(t) per observation variable, then a linearity check at
random points.  Linear in at most three variables (none of them offset / jitter) — the docs' `θ_obs.trend_slope *
(epoch - 57000)`, polynomials, fixed-period sinusoids — is offloaded as per-epoch basis values; the zero trend gives
`(Symbol[], zeros(0, n), nothing)`; anything else (`nothing`) keeps the observation in Julia.
"""
function probe_trend(obs, θ_obs)
    ep = collect(Float64, obs.table.epoch); n = length(ep)
    hasproperty(obs, :trend_function) || return (Symbol[], zeros(0, n), nothing)
    f = obs.trend_function
    ks = keys(θ_obs)
    at(vals) = NamedTuple{ks}(Tuple(vals))
    try
        z = zeros(length(ks))
        b0 = [Float64(f(at(z), t)) for t in ep]
        names = Symbol[]; rows = Vector{Float64}[]
        for (j, k) in enumerate(ks)
            e = copy(z); e[j] = 1.0
            b = [Float64(f(at(e), t)) for t in ep] .- b0
            any(!iszero, b) && (push!(names, k); push!(rows, b))
        end
        (length(names) <= 3 && !(:offset in names) && !(:jitter in names)) || return nothing
        rng = Random.Xoshiro(0x0c70)
        for _ in 1:4
            θr = 10 .* randn(rng, length(ks))
            lin = copy(b0)
            for (nm, b) in zip(names, rows)
                lin .+= θr[findfirst(==(nm), ks)] .* b
            end
            got = [Float64(f(at(θr), t)) for t in ep]
            all(isapprox.(got, lin; rtol=1e-11, atol=1e-11 * max(1.0, maximum(abs, lin; init=0.0)))) || return nothing
        end
        B = isempty(rows) ? zeros(0, n) : permutedims(reduce(hcat, rows))          # n_trend x n
        return (names, B, any(!iszero, b0) ? b0 : nothing)
    catch
        return nothing
    end
end

"Orbit basis code of a planet: 0 Visual{KepOrbit}, 1 ThieleInnesOrbit, -1 anything else (AbsoluteVisual, Cartesian,
FixedPosition, RadialVelocityOrbit, ... stay in Julia: the kernel implements the Visual{KepOrbit} constructor only)."
function basis_of(pl::Planet)
    OT = Octofitter._planet_orbit_type(pl)                              # src/variables.jl:1534
    OT === Visual{KepOrbit} && return Int32(0)
    (OT isa Type && OT <: Visual{<:KepOrbit}) && return Int32(0)
    (OT === ThieleInnesOrbit || (OT isa Type && OT <: ThieleInnesOrbit)) && return Int32(1)
    return Int32(-1)
end

"Kernel table kind of an observation, or -1 when it has to stay in Julia."
function kind_of(obs, θ_obs)
    T = nameof(typeof(obs))
    if T === :ObsPriorAstromONeil2019                                   # wraps a table (prior-observable.jl:57-67)
        return nameof(typeof(obs.wrapped_like)) === :PlanetRelAstromObs ? kind_of(obs.wrapped_like, θ_obs) : Int32(-1)
    elseif T === :PlanetRelAstromObs
        return hasproperty(obs.table, :pa) && hasproperty(obs.table, :sep) ? Int32(1) : Int32(0)
    elseif T === :StarAbsoluteRVObs
        return (isnothing(obs.gaussian_process) && !isnothing(probe_trend(obs, θ_obs))) ? Int32(2) : Int32(-1)
    elseif T === :MarginalizedStarAbsoluteRVObs
        return !isnothing(probe_trend(obs, θ_obs)) ? Int32(3) : Int32(-1)
    elseif T === :PlanetRelativeRVObs
        return (isnothing(obs.gaussian_process) && !isnothing(probe_trend(obs, θ_obs))) ? Int32(4) : Int32(-1)
    elseif T === :HGCAInstantaneousObs
        return KIND_HGCA
    end
    return Int32(-1)
end

# ------------------------------------------------------------------------------------------------------------------
# The plugin observation
# ------------------------------------------------------------------------------------------------------------------
"""
    B200Likelihood

System-level observation standing for every offloaded table of a system.  `paths[k]` is the access path of kernel
input k inside `ctx.θ_system` (e.g. `(:planets, :b, :a)`, `(:observations, :rv, :jitter)`), or a `Float64` constant.
"""
struct B200Likelihood{TPaths<:Tuple} <: Octofitter.AbstractObs
    ctx::Context
    paths::TPaths
    priors::Priors
    derived::Derived
    name::String
end
Octofitter.likelihoodname(o::B200Likelihood) = o.name
Octofitter._isprior(::B200Likelihood) = false
Octofitter.requires_solutions_for_zero_mass(::B200Likelihood) = false   # reads no PlanetOrbits solutions (system.jl:18)
Octofitter.likeobj_from_epoch_subset(o::B200Likelihood, obs_inds) = o
Octofitter.TypedTables.Table(::B200Likelihood) = nothing
Octofitter.generate_from_params(o::B200Likelihood, ctx::Octofitter.SystemObservationContext; add_noise=false) = o
Base.show(io::IO, ::MIME"text/plain", o::B200Likelihood) =
    print(io, "B200Likelihood: $(length(o.paths)) kernel inputs, $(total_epochs(o.ctx)) epochs on the GPU")

total_epochs(c::Context) = ccall((:octo_total_epochs, LIB), Int64, (Ptr{Cvoid},), c.ptr)

@inline fetch_input(θ, p::Tuple) = foldl(getproperty, p; init=θ)
@inline fetch_input(θ, p::Float64) = p

function Octofitter.ln_like(o::B200Likelihood, ctx::Octofitter.SystemObservationContext)
    x = promote(map(p -> fetch_input(ctx.θ_system, p), o.paths)...)
    return evaluate(o, x)
end

# Float64 parameters: value only (K1v)
function evaluate(o::B200Likelihood, x::NTuple{K,Float64}) where {K}
    xv = collect(x); ll = Ref{Cdouble}(0)
    GC.@preserve xv check(ccall((:octo_logp, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Int64, Int64, Ref{Cdouble}),
                                o.ctx.ptr, xv, 1, 1, ll))
    return ll[]
end
# ForwardDiff duals: one value + gradient call, chain rule through the duals' partials
function evaluate(o::B200Likelihood, x::NTuple{K,ForwardDiff.Dual{Tag,Float64,N}}) where {K,Tag,N}
    xv = Vector{Cdouble}(undef, K)
    @inbounds for k in 1:K
        xv[k] = ForwardDiff.value(x[k])
    end
    ll = Ref{Cdouble}(0); g = Vector{Cdouble}(undef, K)
    GC.@preserve xv g check(ccall((:octo_logp_grad, LIB), Cint,
                                  (Ptr{Cvoid}, Ptr{Cdouble}, Int64, Int64, Ref{Cdouble}, Ptr{Cdouble}),
                                  o.ctx.ptr, xv, 1, 1, ll, g))
    parts = zero(ForwardDiff.Partials{N,Float64})
    @inbounds for k in 1:K
        parts += g[k] * ForwardDiff.partials(x[k])          # g is zero for an invalid chain (ll = -Inf)
    end
    return ForwardDiff.Dual{Tag}(ll[], parts)
end
evaluate(o::B200Likelihood, x::Tuple) =
    error("B200Likelihood: parameters of type $(eltype(x)) are not supported (Float64 or first-order ForwardDiff duals)")

# ------------------------------------------------------------------------------------------------------------------
# Building the offloaded system
# ------------------------------------------------------------------------------------------------------------------
struct Offload
    ctx::Context
    paths::Vector{Any}                    # access path (Tuple of Symbols) or Float64 constant per kernel input
    offloaded::Vector{Any}                # the observation objects that were turned into tables
end

function build_context(system::System, θ0; device::Integer=0)
    paths = Any[]
    col(p) = (i = findfirst(==(p), paths); isnothing(i) ? (push!(paths, p); Int32(length(paths) - 1)) : Int32(i - 1))
    P = length(system.planets)
    (1 <= P <= MAXP) || error("libocto_b200 supports 1..$MAXP planets")
    idx = Dict(k => fill(Int32(-1), MAXP) for k in (:plx, :a, :e, :i, :ω, :Ω, :tp, :M, :mass, :A, :B, :F, :G))
    basis = fill(Int32(0), MAXP)
    planet_ok = falses(P)
    for (ip, pl) in enumerate(system.planets)
        b = basis_of(pl)
        b < 0 && continue
        θp = getproperty(θ0.planets, pl.name)
        # orbit kwargs are merge(θ_system, θ_planet): the planet's own variable wins (system.jl:117)
        where_is(k) = hasproperty(θp, k) ? (:planets, pl.name, k) : (hasproperty(θ0, k) ? (k,) : nothing)
        # tp comes last: when it is derived by θ_at_epoch_to_tperi the device-side parameterisation (DevicePosterior)
        # needs its arguments — the other elements and the position-angle variable θ — as EARLIER kernel inputs
        need = b == 0 ? (:plx, :M, :a, :e, :i, :ω, :Ω, :tp) : (:plx, :M, :e, :A, :B, :F, :G, :tp)
        locs = map(where_is, need)
        any(isnothing, locs) && continue
        for (k, loc) in zip(need, locs)
            if k === :tp
                t = match_tperi(get(pl.derived.variables, :tp, nothing))
                (!isnothing(t) && hasproperty(θp, t.θ)) && col((:planets, pl.name, t.θ))     # unused by the kernel itself
            end
            idx[k][ip] = col(loc)
        end
        hasproperty(θp, :mass) && (idx[:mass][ip] = col((:planets, pl.name, :mass)))
        basis[ip] = b; planet_ok[ip] = true
    end
    keep = Any[]; blocks = OctoObsBlock[]; offloaded = Any[]
    f64(x) = (v = collect(Float64, x); push!(keep, v); v)
    ptr(x) = isnothing(x) ? Ptr{Cdouble}(0) : pointer(x)
    function add_block(obs, ip, path)
        θobs = (hasproperty(foldl(getproperty, path[1:end-1]; init=θ0), path[end])) ? foldl(getproperty, path; init=θ0) : (;)
        k = kind_of(obs, θobs); k < 0 && return
        astrom = k <= 1
        # a table is offloaded only if every planet it touches is (system-level tables touch all of them; planet tables
        # touch theirs and, through the reflex terms, every companion with a mass: relative-astrometry.jl:117-133)
        all(planet_ok) || return
        any(b -> b == 1, basis[1:P]) && !astrom && k != KIND_HGCA && return      # RV of/with a Thiele-Innes planet: not offloaded
        v(sym) = hasproperty(θobs, sym) ? col((path..., sym)) : Int32(-1)
        if k == KIND_HGCA
            (hasproperty(θ0, :pmra) && hasproperty(θ0, :pmdec)) || return
            all(i -> idx[:mass][i] >= 0, 1:P) || return
            tbl = obs.table                                              # rows built by the reference ctor (hgca.jl:94-110)
            code = map(zip(tbl.meas, tbl.inst)) do (m, inst)
                Float64((inst === :gaia ? 2 : 0) + (m === :dec ? 1 : 0))
            end
            # catalogue proper motions, and the error model exactly as the ctor built it (the `factor` keyword is folded
            # into the covariance of dist_hip / dist_hg / dist_gaia and not kept anywhere else: hgca.jl:119-139)
            h = obs.hgca
            aux = f64(vcat(map(("hip", "hg", "gaia")) do t
                Σ = Matrix(getproperty(h, Symbol("dist_", t)).Σ)
                s1, s2 = sqrt(Σ[1, 1]), sqrt(Σ[2, 2])
                [getproperty(h, Symbol("pmra_", t))[1], getproperty(h, Symbol("pmdec_", t))[1], s1, s2, Σ[1, 2] / (s1 * s2)]
            end...))
            ep = f64(tbl.epoch); y1 = f64(code)
            push!(blocks, OctoObsBlock(k, Int32(-1), length(ep), 0, ptr(ep), ptr(y1), ptr(nothing), ptr(nothing), ptr(nothing),
                                       ptr(nothing), -1, -1, -1, -1, 0, col((:pmra,)), col((:pmdec,)), 0, ptr(aux),
                                       0, (Int32(-1), Int32(-1), Int32(-1)), Ptr{Cdouble}(0), Ptr{Cdouble}(0)))
            push!(offloaded, obs); return
        end
        tbl = nameof(typeof(obs)) === :ObsPriorAstromONeil2019 ? obs.wrapped_like.table : obs.table
        ep = f64(tbl.epoch)
        if k == 0;     y1, y2, s1, s2 = f64(tbl.ra), f64(tbl.dec), f64(tbl.σ_ra), f64(tbl.σ_dec)
        elseif k == 1; y1, y2, s1, s2 = f64(tbl.pa), f64(tbl.sep), f64(tbl.σ_pa), f64(tbl.σ_sep)
        else;          y1, s1 = f64(tbl.rv), f64(tbl.σ_rv); y2 = s2 = nothing
        end
        cor = (astrom && hasproperty(tbl, :cor)) ? f64(tbl.cor) : nothing
        (k == 3 && !hasproperty(θobs, :jitter)) && return                # the reference reads θ_obs.jitter unconditionally
        (k in (2, 3)) && !all(i -> idx[:mass][i] >= 0, 1:P) && return
        # RV: the trend closure as basis values (probe_trend; kind_of already made sure it is linear)
        n_tr = Int32(0); idx_tr = (Int32(-1), Int32(-1), Int32(-1)); tb = tc = nothing
        if !astrom
            tnames, TB, t0 = probe_trend(obs, θobs)
            n_tr = Int32(length(tnames))
            idx_tr = ntuple(i -> i <= n_tr ? v(tnames[i]) : Int32(-1), 3)
            n_tr > 0 && (tb = f64(vec(permutedims(TB))))               # row-major [n_trend x n_epochs]
            isnothing(t0) || (tc = f64(t0))
        end
        push!(blocks, OctoObsBlock(k, Int32(ip - 1), length(ep), isnothing(cor) ? 0 : 1, ptr(ep), ptr(y1), ptr(y2), ptr(s1),
                                   ptr(s2), ptr(cor), v(:jitter), astrom ? v(:platescale) : Int32(-1),
                                   astrom ? v(:northangle) : Int32(-1), astrom ? Int32(-1) : v(:offset),
                                   Int32(nameof(typeof(obs)) === :ObsPriorAstromONeil2019), -1, -1, 0, Ptr{Cdouble}(0),
                                   n_tr, idx_tr, ptr(tb), ptr(tc)))
        push!(offloaded, obs)
    end
    # the reference's summation order: planet observations first, then system observations (system.jl:223-236)
    for (ip, pl) in enumerate(system.planets), obs in pl.observations
        add_block(obs, ip, (:planets, pl.name, :observations, Symbol(normalizename(likelihoodname(obs)))))
    end
    for obs in system.observations
        add_block(obs, 0, (:observations, Symbol(normalizename(likelihoodname(obs)))))
    end
    isempty(blocks) && return nothing
    # planets that no offloaded table touches still need valid columns: they were only registered if offloadable
    tup(k) = ntuple(i -> idx[k][i], MAXP)
    layout = OctoLayout(P, length(paths), tup(:plx), tup(:a), tup(:e), tup(:i), tup(:ω), tup(:Ω), tup(:tp), tup(:M), tup(:mass),
                        ntuple(i -> basis[i], MAXP), tup(:A), tup(:B), tup(:F), tup(:G))
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep blocks begin
        check(ccall((:octo_create, LIB), Cint,
                    (Ref{OctoConstants}, Ref{OctoLayout}, Ptr{OctoObsBlock}, Int32, Int32, Ref{Ptr{Cvoid}}),
                    OctoConstants(), layout, blocks, length(blocks), device, out))
    end
    return Offload(Context(out[], keep), paths, offloaded)
end

"""
    b200_system(system; device=0) -> System

The same system with every offloadable observation blanked (variables kept) and one `B200Likelihood` appended.
Returns `system` itself when nothing can be offloaded.
"""
function b200_system(system::System; device::Integer=0, θ0=nothing)
    if isnothing(θ0)
        arr2nt = Octofitter.make_arr2nt(system)
        θ0 = arr2nt(Octofitter.make_prior_sampler(system)(Random.default_rng()))
    end
    off = build_context(system, θ0; device)
    isnothing(off) && return system
    isoff(obs) = any(o -> o === obs, off.offloaded)
    blank(obs) = isoff(obs) ? BlankLikelihood((obs.priors, obs.derived), likelihoodname(obs)) : obs
    planets = map(system.planets) do pl
        Planet(name=pl.name, basis=Octofitter.orbittype(pl), variables=(pl.priors, pl.derived),
               observations=map(blank, collect(pl.observations)))
    end
    gpu = B200Likelihood(off.ctx, Tuple(off.paths), Priors(), Derived(), "b200")
    return System(name=system.name, variables=(system.priors, system.derived), companions=planets,
                  observations=[map(blank, collect(system.observations))..., gpu])
end

"""
    B200Model(system::System; device=0, kwargs...) -> Octofitter.LogDensityModel
    B200Model(model::Octofitter.LogDensityModel; device=0, kwargs...)

A genuine `Octofitter.LogDensityModel` whose epoch loop runs on the GPU.  Starting points of an existing model are
carried over.  `kwargs` go to the `LogDensityModel` constructor (`verbosity`, `autodiff`, ...).
"""
function B200Model(system::System; device::Integer=0, kwargs...)
    return Octofitter.LogDensityModel(b200_system(system; device); kwargs...)
end
function B200Model(model::Octofitter.LogDensityModel; device::Integer=0, kwargs...)
    m = B200Model(model.system; device, kwargs...)
    m.starting_points = model.starting_points
    return m
end

"The `B200Likelihood` of a model built by `B200Model`, or `nothing`."
function b200_likelihood(model::Octofitter.LogDensityModel)
    i = findfirst(o -> o isa B200Likelihood, collect(model.system.observations))
    return isnothing(i) ? nothing : model.system.observations[i]
end

# ------------------------------------------------------------------------------------------------------------------
# Asynchronous evaluation of natural-space inputs (octo_logp_grad_begin / octo_wait): the caller overlaps its own
# host work — e.g. the priors of the next batch — with the GPU.  X is N x n_in (column-major == the ABI layout).
# ------------------------------------------------------------------------------------------------------------------
struct Ticket
    ptr::Ptr{Cvoid}
    keep::Tuple
end
function logp_grad_begin(o::B200Likelihood, X::Matrix{Float64}, ll::Vector{Float64}, G::Matrix{Float64})
    N = size(X, 1)
    t = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:octo_logp_grad_begin, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Cdouble}, Int64, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ref{Ptr{Cvoid}}), o.ctx.ptr, X, N, N, ll, G, t))
    return Ticket(t[], (X, ll, G))
end
ready(t::Ticket) = ccall((:octo_ready, LIB), Cint, (Ptr{Cvoid},), t.ptr) != 0
wait_for(t::Ticket) = (check(ccall((:octo_wait, LIB), Cint, (Ptr{Cvoid},), t.ptr)); (t.keep[2], t.keep[3]))

# ------------------------------------------------------------------------------------------------------------------
# DevicePosterior: the standard parameterisation on the device (SURVEY.md §8f N1)
# ------------------------------------------------------------------------------------------------------------------
"OctoPrior of a Distributions.jl prior, or nothing (Appendix B of SURVEY.md: the six families of the reference's docs)."
function octo_prior(d)
    if d isa Normal
        return OctoPrior(PRIOR_NORMAL, 0, (d.μ, d.σ, 0.0, 0.0))
    elseif d isa Uniform
        return OctoPrior(PRIOR_UNIFORM, 0, (d.a, d.b, 0.0, 0.0))
    elseif d isa LogUniform
        return OctoPrior(PRIOR_LOGUNIFORM, 0, (d.a, d.b, 0.0, 0.0))
    elseif nameof(typeof(d)) === :Sine                                  # src/distributions.jl:15-40
        return OctoPrior(PRIOR_SINE, 0, (0.0, 0.0, 0.0, 0.0))
    elseif d isa Truncated && d.untruncated isa Normal
        lo = isnothing(d.lower) ? -Inf : Float64(d.lower); hi = isnothing(d.upper) ? Inf : Float64(d.upper)
        return OctoPrior(PRIOR_TRUNCNORMAL, 0, (d.untruncated.μ, d.untruncated.σ, lo, hi))
    end
    return nothing
end

# `atan(y, x) / 2π * domain` as UniformCircular's expandparam writes it (src/variables.jl:279-283)
function match_circular(ex)
    (ex isa Expr && ex.head === :call && ex.args[1] === :* && length(ex.args) == 3) || return nothing
    num, dom = ex.args[2], ex.args[3]
    dom isa Real || return nothing
    (num isa Expr && num.head === :call && num.args[1] === :/ && length(num.args) == 3) || return nothing
    at, den = num.args[2], num.args[3]
    (den == :(2π)) || return nothing
    (at isa Expr && at.head === :call && at.args[1] === :atan && length(at.args) == 3) || return nothing
    (at.args[2] isa Symbol && at.args[3] isa Symbol) || return nothing
    return (y=at.args[2], x=at.args[3], domain=Float64(dom))
end
# `θ_at_epoch_to_tperi(θ, t_ref; M=system.M, e, a, i, ω, Ω)` (src/parameterizations.jl:6-69; docs/src/rel-astrom.md)
function match_tperi(ex)
    (ex isa Expr && ex.head === :call && ex.args[1] === :θ_at_epoch_to_tperi) || return nothing
    pos = filter(a -> !(a isa Expr && a.head === :parameters), ex.args[2:end])
    par = filter(a -> a isa Expr && a.head === :parameters, ex.args[2:end])
    (length(pos) == 2 && pos[1] isa Symbol && pos[2] isa Real && length(par) == 1) || return nothing
    kw = Dict{Symbol,Any}()
    for a in par[1].args
        if a isa Symbol
            kw[a] = a
        elseif a isa Expr && a.head === :kw
            kw[a.args[1]] = a.args[2]
        else
            return nothing
        end
    end
    return (θ=pos[1], t_ref=Float64(pos[2]), kw=kw)
end

"""
    DevicePosterior(model) -> DevicePosterior | nothing

`model` must come from `B200Model`.  Succeeds when the model consists of offloaded observations only (besides the
UnitLengthPrior terms of UniformCircular variables, which the device adds itself) and every variable is a prior of the
six standard families, a UniformCircular angle, a number, an alias of a system variable (`plx = system.plx`) or
`θ_at_epoch_to_tperi(θ, t_ref; M=…, e, a, i, ω, Ω)`.
"""
struct DevicePosterior
    model::Any
    like::B200Likelihood
    D::Int
end
function DevicePosterior(model::Octofitter.LogDensityModel)
    like = b200_likelihood(model); isnothing(like) && return nothing
    sys = model.system
    # nothing but blanked tables, UnitLengthPriors and the plugin may remain
    remaining_ok(obs) = obs isa BlankLikelihood || obs isa B200Likelihood || nameof(typeof(obs)) === :UnitLengthPrior
    all(remaining_ok, sys.observations) && all(pl -> all(remaining_ok, pl.observations), sys.planets) || return nothing
    # θ_t order: system priors, system-observation priors, then per planet its priors and its observations' priors
    # (src/variables.jl:691-730)
    priors = OctoPrior[]; where_prior = Dict{Any,Int}()
    function add_priors(prefix, pri::Priors)
        for (k, d) in pri.priors
            op = octo_prior(d); isnothing(op) && return false
            push!(priors, op); where_prior[(prefix..., k)] = length(priors) - 1
        end
        return true
    end
    add_priors((), sys.priors) || return nothing
    for o in sys.observations
        add_priors((:observations, Symbol(normalizename(likelihoodname(o)))), o.priors) || return nothing
    end
    for pl in sys.planets
        add_priors((:planets, pl.name), pl.priors) || return nothing
        for o in pl.observations
            add_priors((:planets, pl.name, :observations, Symbol(normalizename(likelihoodname(o)))), o.priors) || return nothing
        end
    end
    D = length(priors); D == model.D || return nothing
    # one definition per kernel input, in input order; derived expressions are looked up in the scope the path names
    derived_of(path) = length(path) == 1 ? sys.derived :
        path[1] === :observations ? only(filter(o -> Symbol(normalizename(likelihoodname(o))) === path[2], collect(sys.observations))).derived :
        length(path) == 3 ? only(filter(p -> p.name === path[2], collect(sys.planets))).derived :
        only(filter(o -> Symbol(normalizename(likelihoodname(o))) === path[4],
                    collect(only(filter(p -> p.name === path[2], collect(sys.planets))).observations))).derived
    defs = Vector{OctoInputDef}(undef, length(like.paths))
    zero8 = ntuple(_ -> Int32(0), 8)
    set8(v...) = ntuple(i -> i <= length(v) ? Int32(v[i]) : Int32(0), 8)
    input_of = Dict{Any,Int}(p => k - 1 for (k, p) in enumerate(like.paths) if p isa Tuple)
    pending_tperi = Tuple{Int,Any,Any}[]
    for (k, p) in enumerate(like.paths)
        if p isa Float64
            defs[k] = OctoInputDef(IN_CONST, zero8, p); continue
        end
        if haskey(where_prior, p)
            defs[k] = OctoInputDef(IN_PARAM, set8(where_prior[p]), 0.0); continue
        end
        scope, name = p[1:end-1], p[end]
        ex = get(derived_of(p).variables, name, nothing)
        isnothing(ex) && return nothing
        if ex isa Real
            defs[k] = OctoInputDef(IN_CONST, zero8, Float64(ex)); continue
        end
        # alias of a system variable: `plx = system.plx`
        if ex isa Expr && ex.head === :. && ex.args[1] === :system && ex.args[2] isa QuoteNode && haskey(where_prior, (ex.args[2].value,))
            defs[k] = OctoInputDef(IN_PARAM, set8(where_prior[(ex.args[2].value,)]), 0.0); continue
        end
        c = match_circular(ex)
        if !isnothing(c) && haskey(where_prior, (scope..., c.x)) && haskey(where_prior, (scope..., c.y))
            defs[k] = OctoInputDef(IN_CIRC, set8(where_prior[(scope..., c.x)], where_prior[(scope..., c.y)]), c.domain); continue
        end
        t = match_tperi(ex)
        isnothing(t) && return nothing
        push!(pending_tperi, (k, scope, t))
    end
    # θ_at_epoch_to_tperi arguments are EARLIER kernel inputs: θ itself is not an orbit element, so it may have to be
    # appended — which the plugin context cannot do after the fact.  Require it to be there already (it is whenever a
    # table has been offloaded for the planet and θ is one of the planet's variables the extractor registered) ...
    for (k, scope, t) in pending_tperi
        look(s) = s isa Symbol ? get(input_of, (scope..., s), get(input_of, (s,), nothing)) :
                  (s isa Expr && s.head === :. && s.args[1] === :system) ? get(input_of, (s.args[2].value,), nothing) : nothing
        ti = basis_of(only(filter(pl -> pl.name === scope[2], collect(sys.planets)))) == 1
        need = ti ? (:M, :e, :plx, :A, :B, :F, :G) : (:M, :e, :a, :i, :ω, :Ω)
        args = Any[look(t.θ)]
        for s in need
            push!(args, look(get(t.kw, s, s)))
        end
        # ... otherwise this model needs the plugin path
        (any(isnothing, args) || any(a -> a >= k - 1, args)) && return nothing
        defs[k] = OctoInputDef(ti ? IN_TPERI_TI : IN_TPERI, set8(args...), t.t_ref)
    end
    check(ccall((:octo_set_parameterization, LIB), Cint, (Ptr{Cvoid}, Ptr{OctoPrior}, Int32, Ptr{OctoInputDef}),
                like.ctx.ptr, priors, D, defs))
    return DevicePosterior(model, like, D)
end

"log posterior and gradient of a D x N matrix of unconstrained vectors, one launch (octo_logpost_grad)"
function LogDensityProblems.logdensity_and_gradient(p::DevicePosterior, Θ::AbstractMatrix)
    N = size(Θ, 2)
    T = Matrix{Cdouble}(permutedims(Θ))                       # N x D column-major: chain index fastest (the ABI layout)
    lp = Vector{Cdouble}(undef, N); G = Matrix{Cdouble}(undef, N, p.D)
    check(ccall((:octo_logpost_grad, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Int64, Int64, Ptr{Cdouble}, Ptr{Cdouble}),
                p.like.ctx.ptr, T, N, N, lp, G))
    return lp, permutedims(G)
end
function LogDensityProblems.logdensity(p::DevicePosterior, Θ::AbstractMatrix)
    N = size(Θ, 2)
    T = Matrix{Cdouble}(permutedims(Θ)); lp = Vector{Cdouble}(undef, N)
    check(ccall((:octo_logpost_grad, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Int64, Int64, Ptr{Cdouble}, Ptr{Cdouble}),
                p.like.ctx.ptr, T, N, N, lp, Ptr{Cdouble}(0)))
    return lp
end
LogDensityProblems.logdensity_and_gradient(p::DevicePosterior, θ::AbstractVector) =
    ((lp, G) = LogDensityProblems.logdensity_and_gradient(p, reshape(θ, :, 1)); (lp[1], vec(G)))
LogDensityProblems.logdensity(p::DevicePosterior, θ::AbstractVector) = LogDensityProblems.logdensity(p, reshape(θ, :, 1))[1]
LogDensityProblems.dimension(p::DevicePosterior) = p.D
LogDensityProblems.capabilities(::Type{DevicePosterior}) = LogDensityProblems.LogDensityOrder{1}()

"""
    hmc_run(p, Θ0; n_iter, n_leapfrog, step_size, inv_mass=nothing, seed=0) -> (Θ_final, lp_final, accept_rate, samples)

The device-resident, chain-batched static-trajectory HMC explorer (`octo_hmc_run`): all N columns of Θ0 move in
lockstep inside one launch of the trajectory-resident kernel.  `samples` is D x N x n_iter.
"""
function hmc_run(p::DevicePosterior, Θ0::AbstractMatrix; n_iter::Integer, n_leapfrog::Integer, step_size::Real,
                 inv_mass::Union{Nothing,Vector{Float64}}=nothing, seed::Integer=0, keep_samples::Bool=true)
    N = size(Θ0, 2); D = p.D
    T0 = Matrix{Cdouble}(permutedims(Θ0)); Tf = similar(T0); lpf = Vector{Cdouble}(undef, N); acc = similar(lpf)
    S = keep_samples ? Array{Cdouble}(undef, N, D, n_iter) : nothing
    L = keep_samples ? Matrix{Cdouble}(undef, N, n_iter) : nothing
    check(ccall((:octo_hmc_run, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Cdouble}, Int64, Int64, Int32, Int32, Cdouble, Ptr{Cdouble}, UInt64, Ptr{Cdouble}, Ptr{Cdouble},
                 Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                p.like.ctx.ptr, T0, N, N, n_iter, n_leapfrog, step_size, isnothing(inv_mass) ? Ptr{Cdouble}(0) : pointer(inv_mass),
                UInt64(seed), isnothing(S) ? Ptr{Cdouble}(0) : pointer(S), isnothing(L) ? Ptr{Cdouble}(0) : pointer(L), Tf, lpf, acc))
    return permutedims(Tf), lpf, acc, isnothing(S) ? nothing : permutedims(S, (2, 1, 3))
end

"""
    pt_init(p; id, rank, world, n_local, seed) / pt_unique_id()

Communicator for the sharded ladder (one process per GPU; the 128-byte id of rank 0 travels over whatever the launcher
offers, e.g. MPI.bcast).  world = 1 needs no id.
"""
pt_unique_id() = (id = zeros(UInt8, 128); check(ccall((:octo_pt_unique_id, LIB), Cint, (Ptr{UInt8},), id)); id)
pt_init(p::DevicePosterior; id=nothing, rank::Integer=0, world::Integer=1, n_local::Integer, seed::Integer=0) =
    check(ccall((:octo_pt_init, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Int32, Int32, Int32, UInt64),
                p.like.ctx.ptr, isnothing(id) ? Ptr{UInt8}(0) : pointer(id), rank, world, n_local, UInt64(seed)))

"""
    pt_hmc_run(p, Θ0_local, ladder; n_rounds, n_iter=1, n_leapfrog, step_size, inv_mass=nothing, seed=0, sharded=false)

Device-resident parallel tempering between the prior-only reference (weight 0) and the target (weight 1) — the path
Pigeons tempers along (ext/OctofitterPigeonsExt:61-67).  `sharded = true` after `pt_init`: this process holds the
columns of its rank, `ladder` is the whole ladder; one ncclAllGather of (ℓ_ref, ℓ_target) per round
(`octo_pt_hmc_run_dist`).  Returns a NamedTuple (θ, lp, ll, β, rung, swap_counts, cold_trace, accept).
"""
function pt_hmc_run(p::DevicePosterior, Θ0::AbstractMatrix, ladder::Vector{Float64}; n_rounds::Integer, n_iter::Integer=1,
                    n_leapfrog::Integer, step_size::Real, inv_mass::Union{Nothing,Vector{Float64}}=nothing, seed::Integer=0,
                    sharded::Bool=false)
    N = size(Θ0, 2); D = p.D; R = length(ladder)
    T0 = Matrix{Cdouble}(permutedims(Θ0)); Tf = similar(T0)
    lp = Vector{Cdouble}(undef, N); ll = similar(lp); β = similar(lp); acc = similar(lp)
    rung = Vector{Int32}(undef, N); swaps = Vector{Cdouble}(undef, R - 1); cold = Matrix{Cdouble}(undef, D, n_rounds)
    f = sharded ? :octo_pt_hmc_run_dist : :octo_pt_hmc_run
    sharded || R == N || error("ladder needs one weight per chain")
    check(ccall((f, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Cdouble}, Int64, Int64, Ptr{Cdouble}, Int32, Int32, Int32, Cdouble, Ptr{Cdouble}, UInt64,
                 Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Int32}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                p.like.ctx.ptr, T0, N, N, ladder, n_rounds, n_iter, n_leapfrog, step_size,
                isnothing(inv_mass) ? Ptr{Cdouble}(0) : pointer(inv_mass), UInt64(seed), Tf, lp, ll, β, rung, swaps, cold, acc))
    return (θ=permutedims(Tf), lp=lp, ll=ll, β=β, rung=rung, swap_counts=swaps, cold_trace=cold, accept=acc)
end

export B200Model, b200_system, B200Likelihood, DevicePosterior, hmc_run, pt_hmc_run, pt_init, pt_unique_id

end # module
