# OctofitterB200.jl — thin Julia glue over libocto_b200.so (include/octo_b200.h).
#
# STATUS: written against the C ABI and the reference sources at /root/reference (Octofitter v8.3.0);
# NOT executed in the build environment (no julia binary in the image or on the GPU box).  The Python
# mirror octofitter.jl_b200/model.py makes the same calls in the same order and IS exercised by the tests.
#
# What it does (SURVEY.md §8b "split"):
#   * every observation the kernel supports (PlanetRelAstromObs, StarAbsoluteRVObs / Marginalized /
#     PlanetRelativeRVObs without GP and with the default trend_function) becomes an OctoObsBlock;
#   * a "rest" model is built the way `prior_only_model` does (src/cross-validation.jl:60-99), blanking only
#     the offloaded observations, so D, the parameter order and arr2nt are unchanged: it supplies ln_prior,
#     the non-epoch likelihood terms (UnitLengthPrior, UserLikelihood, ...) and their ForwardDiff gradient;
#   * kernel inputs are read out of model.arr2nt(model.invlink(θ_t)) by name; their Jacobian w.r.t. θ_t
#     comes from one ForwardDiff pass over that extractor (no epoch loop on the host);
#   * logp = ℓπ_rest + ll_kernel,  ∇ = ∇ℓπ_rest + Jᵀ g_in.
module OctofitterB200

using Octofitter, PlanetOrbits, ForwardDiff, LogDensityProblems
import Octofitter: Planet, System, LogDensityModel, BlankLikelihood, normalizename, likelihoodname

const LIB = get(ENV, "OCTO_B200_LIB", joinpath(@__DIR__, "..", "octofitter.jl_b200", "lib", "libocto_b200.so"))
const MAXP = 4

struct OctoConstants
    kepler_year_days::Cdouble; year2day::Cdouble; rad2as::Cdouble; pc2au::Cdouble
    au2m::Cdouble; sec2year::Cdouble; mjup2msol::Cdouble
end
# R1 of SURVEY.md: never hard-code — inject PlanetOrbits' live values
OctoConstants() = OctoConstants(PlanetOrbits.kepler_year_to_julian_day_conversion_factor, PlanetOrbits.year2day_julian,
    PlanetOrbits.rad2as, PlanetOrbits.pc2au, PlanetOrbits.au2m, PlanetOrbits.sec2year_julian, PlanetOrbits.mjup2msol_IAU)

struct OctoObsBlock
    kind::Int32; planet::Int32; n_epochs::Int32; has_cor::Int32
    epoch::Ptr{Cdouble}; y1::Ptr{Cdouble}; y2::Ptr{Cdouble}; s1::Ptr{Cdouble}; s2::Ptr{Cdouble}; cor::Ptr{Cdouble}
    idx_jitter::Int32; idx_platescale::Int32; idx_northangle::Int32; idx_offset::Int32
    obs_prior::Int32                       # 1: astrometry table wrapped in ObsPriorAstromONeil2019
    idx_pmra::Int32; idx_pmdec::Int32; reserved::Int32   # kind 5 (HGCAInstantaneousObs): system proper-motion columns
    aux::Ptr{Cdouble}                      # kind 5: the 15 catalogue numbers (see include/octo_b200.h)
end

struct OctoLayout
    n_planets::Int32; n_in::Int32
    idx_plx::NTuple{MAXP,Int32}; idx_a::NTuple{MAXP,Int32}; idx_e::NTuple{MAXP,Int32}; idx_i::NTuple{MAXP,Int32}
    idx_w::NTuple{MAXP,Int32}; idx_W::NTuple{MAXP,Int32}; idx_tp::NTuple{MAXP,Int32}; idx_M::NTuple{MAXP,Int32}
    idx_mass::NTuple{MAXP,Int32}
    basis::NTuple{MAXP,Int32}              # 0 Visual{KepOrbit}, 1 ThieleInnesOrbit (then idx_A.. are used instead of idx_a, idx_i, idx_w, idx_W)
    idx_A::NTuple{MAXP,Int32}; idx_B::NTuple{MAXP,Int32}; idx_F::NTuple{MAXP,Int32}; idx_G::NTuple{MAXP,Int32}
end

octo_error() = unsafe_string(ccall((:octo_last_error, LIB), Cstring, ()))
check(rc) = rc == 0 ? nothing : error("libocto_b200: $(octo_error())")

"Can this observation be offloaded?  (a9-a12 of SURVEY.md; GP and custom trends stay in Julia.)"
function kind_of(obs)
    T = nameof(typeof(obs))
    # HGCAInstantaneousObs (kind 5) is supported by the library; this glue does not offload it yet (it would build the
    # rows from obs.table.epoch / .meas / .inst and aux from obs.hgca, see octofitter.jl_b200/model.py)
    # the observable-based prior wraps an astrometry table (src/likelihoods/prior-observable.jl:57-67)
    T === :ObsPriorAstromONeil2019 && nameof(typeof(obs.wrapped_like)) === :PlanetRelAstromObs && return kind_of(obs.wrapped_like)
    if T === :PlanetRelAstromObs
        return hasproperty(obs.table, :pa) && hasproperty(obs.table, :sep) ? Int32(1) : Int32(0)
    elseif T === :StarAbsoluteRVObs && isnothing(obs.gaussian_process)
        return Int32(2)
    elseif T === :MarginalizedStarAbsoluteRVObs
        return Int32(3)
    elseif T === :PlanetRelativeRVObs && isnothing(obs.gaussian_process)
        return Int32(4)
    end
    return Int32(-1)
end

mutable struct B200Model{M,R}
    model::M                 # the untouched Octofitter.LogDensityModel (link/invlink/arr2nt/starting_points…)
    rest::R                  # same parameters, offloaded observations blanked
    ctx::Ptr{Cvoid}
    n_in::Int
    extract::Function        # θ_nt -> Vector of kernel inputs (natural space), fixed order
    keep::Vector{Any}        # table columns kept alive until octo_create copied them
    D::Int
end

"""
    B200Model(model::Octofitter.LogDensityModel; device=0)

Drop-in for the sampler-facing surface (`LogDensityProblems.logdensity[_and_gradient]`, `dimension`,
`capabilities`, callable for Pigeons) with the epoch loop on the GPU.
"""
function B200Model(model::LogDensityModel; device::Integer=0)
    system = model.system
    names = Tuple{Vararg{Symbol}}[]              # access path of every kernel input inside θ_nt
    col(path) = (i = findfirst(==(path), names); isnothing(i) ? (push!(names, path); length(names) - 1) : i - 1)
    θ0 = model.arr2nt(model.invlink(first(model.starting_points === nothing ? [zeros(model.D)] : model.starting_points)))
    P = length(system.planets)
    P <= MAXP || error("at most $MAXP planets")
    idx = Dict(k => fill(Int32(-1), MAXP) for k in (:plx, :a, :e, :i, :ω, :Ω, :tp, :M, :mass))
    for (ip, pl) in enumerate(system.planets)
        θp = getproperty(θ0.planets, pl.name)
        for k in (:plx, :a, :e, :i, :ω, :Ω, :tp, :M)                 # merge(θ_system, θ_planet): planet wins
            idx[k][ip] = hasproperty(θp, k) ? col((:planets, pl.name, k)) : col((k,))
        end
        hasproperty(θp, :mass) && (idx[:mass][ip] = col((:planets, pl.name, :mass)))
    end
    blocks = OctoObsBlock[]; keep = Any[]; offloaded = Set{Any}()
    function add_block(obs, ip, path)
        k = kind_of(obs); k < 0 && return
        tbl = obs.table
        f64(x) = (v = collect(Float64, x); push!(keep, v); v)
        ep = f64(tbl.epoch)
        if k == 0;     y1, y2, s1, s2 = f64(tbl.ra), f64(tbl.dec), f64(tbl.σ_ra), f64(tbl.σ_dec)
        elseif k == 1; y1, y2, s1, s2 = f64(tbl.pa), f64(tbl.sep), f64(tbl.σ_pa), f64(tbl.σ_sep)
        else;          y1, s1 = f64(tbl.rv), f64(tbl.σ_rv); y2 = s2 = nothing
        end
        cor = (k <= 1 && hasproperty(tbl, :cor)) ? f64(tbl.cor) : nothing
        θobs = foldl(getproperty, path; init=θ0)
        v(sym) = hasproperty(θobs, sym) ? Int32(col((path..., sym))) : Int32(-1)
        ptr(x) = isnothing(x) ? Ptr{Cdouble}(0) : pointer(x)
        push!(blocks, OctoObsBlock(k, Int32(ip - 1), length(ep), isnothing(cor) ? 0 : 1, ptr(ep), ptr(y1), ptr(y2), ptr(s1),
                                   ptr(s2), ptr(cor), v(:jitter), v(:platescale), v(:northangle), v(:offset),
                                   Int32(nameof(typeof(obs)) === :ObsPriorAstromONeil2019), Int32(-1), Int32(-1), Int32(0),
                                   Ptr{Cdouble}(0)))
        push!(offloaded, obs)
    end
    # reference summation order: planet observations first, then system observations (system.jl:223-236)
    for (ip, pl) in enumerate(system.planets), obs in pl.observations
        add_block(obs, ip, (:planets, pl.name, :observations, normalizename(likelihoodname(obs))))
    end
    for obs in system.observations
        add_block(obs, 0, (:observations, normalizename(likelihoodname(obs))))
    end
    layout = OctoLayout(P, length(names), ntuple(i -> idx[:plx][i], MAXP), ntuple(i -> idx[:a][i], MAXP),
        ntuple(i -> idx[:e][i], MAXP), ntuple(i -> idx[:i][i], MAXP), ntuple(i -> idx[:ω][i], MAXP),
        ntuple(i -> idx[:Ω][i], MAXP), ntuple(i -> idx[:tp][i], MAXP), ntuple(i -> idx[:M][i], MAXP),
        ntuple(i -> idx[:mass][i], MAXP),
        # this glue offloads Visual{KepOrbit} planets; ThieleInnesOrbit planets would set basis = 1 and idx_A..idx_G
        ntuple(_ -> Int32(0), MAXP), ntuple(_ -> Int32(-1), MAXP), ntuple(_ -> Int32(-1), MAXP),
        ntuple(_ -> Int32(-1), MAXP), ntuple(_ -> Int32(-1), MAXP))
    ctx = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep begin
        check(ccall((:octo_create, LIB), Cint,
                    (Ref{OctoConstants}, Ref{OctoLayout}, Ptr{OctoObsBlock}, Int32, Int32, Ref{Ptr{Cvoid}}),
                    OctoConstants(), layout, blocks, length(blocks), device, ctx))
    end
    # "rest" system: blank the offloaded observations exactly as prior_only_model does for all of them
    blank(obs) = obs in offloaded ? BlankLikelihood((obs.priors, obs.derived), likelihoodname(obs)) : obs
    planets = map(system.planets) do pl
        Planet(name=pl.name, basis=Octofitter._planet_orbit_type(pl), variables=(pl.priors, pl.derived),
               observations=map(blank, pl.observations))
    end
    rest_sys = System(name=system.name, variables=(system.priors, system.derived), companions=planets,
                      observations=map(blank, system.observations))
    rest = LogDensityModel(rest_sys; verbosity=0)
    paths = copy(names)
    extract = θnt -> [foldl(getproperty, p; init=θnt) for p in paths]
    m = B200Model(model, rest, ctx[], length(names), extract, keep, model.D)
    finalizer(x -> ccall((:octo_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.ctx), m)
    return m
end

inputs(m::B200Model, θ_t) = m.extract(m.model.arr2nt(m.model.invlink(θ_t)))

# ---- single chain (stock AdvancedHMC / Pigeons call pattern)
function LogDensityProblems.logdensity(m::B200Model, θ_t::AbstractVector)
    lp = m.rest.ℓπcallback(θ_t)
    isfinite(lp) || return lp
    x = Float64.(inputs(m, θ_t)); ll = Ref{Cdouble}(0)
    check(ccall((:octo_logp, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Int64, Int64, Ref{Cdouble}), m.ctx, x, 1, 1, ll))
    return lp + ll[]
end
function LogDensityProblems.logdensity_and_gradient(m::B200Model, θ_t::AbstractVector)
    lp, glp = m.rest.∇ℓπcallback(θ_t); glp = copy(glp)          # the reference aliases its gradient buffer
    isfinite(lp) || return lp, glp
    x = Float64.(inputs(m, θ_t))
    J = ForwardDiff.jacobian(t -> inputs(m, t), θ_t)             # n_in x D, no epoch loop
    ll = Ref{Cdouble}(0); g = Vector{Cdouble}(undef, m.n_in)
    check(ccall((:octo_logp_grad, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Int64, Int64, Ref{Cdouble}, Ptr{Cdouble}),
                m.ctx, x, 1, 1, ll, g))
    return lp + ll[], glp .+ J' * g
end
# ---- batch of chains: θ_t is D x N (AdvancedHMC vectorised mode); returns N values and a D x N gradient
function LogDensityProblems.logdensity_and_gradient(m::B200Model, Θ::AbstractMatrix)
    N = size(Θ, 2)
    X = Matrix{Cdouble}(undef, N, m.n_in); Js = Vector{Matrix{Float64}}(undef, N)
    lps = Vector{Float64}(undef, N); G = Matrix{Float64}(undef, m.D, N)
    Threads.@threads for c in 1:N
        θ = view(Θ, :, c)
        lp, glp = m.rest.∇ℓπcallback(collect(θ)); lps[c] = lp; G[:, c] .= glp
        X[c, :] .= inputs(m, θ); Js[c] = ForwardDiff.jacobian(t -> inputs(m, t), collect(θ))
    end
    ll = Vector{Cdouble}(undef, N); g = Matrix{Cdouble}(undef, N, m.n_in)        # column-major N x n_in == the ABI layout
    check(ccall((:octo_logp_grad, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Int64, Int64, Ptr{Cdouble}, Ptr{Cdouble}),
                m.ctx, X, N, N, ll, g))
    for c in 1:N
        G[:, c] .+= Js[c]' * view(g, c, :)
    end
    return lps .+ ll, G
end
LogDensityProblems.dimension(m::B200Model) = m.D
LogDensityProblems.capabilities(::Type{<:B200Model}) = LogDensityProblems.LogDensityOrder{1}()
(m::B200Model)(θ_t) = LogDensityProblems.logdensity(m, θ_t)        # Pigeons target (ext/OctofitterPigeonsExt:10-12)

end # module
