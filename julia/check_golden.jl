# check_golden.jl — closes the one leg this repository cannot: evaluate the committed golden vectors
# (tests/golden/case_*.json, produced by the 60-digit mpmath restatement) with the UNTOUCHED reference, wherever
# Julia >= 1.10 with Octofitter 8.3 + OctofitterRadialVelocity + PlanetOrbits exists.  Not executed in the build image
# (no Julia there); written against the reference sources cited below.
#
#   julia --project=<env with Octofitter> julia/check_golden.jl tests/golden
#
# For every case it rebuilds the observation objects from the JSON tables, constructs the orbits and their solutions
# exactly as the generated `ln_like` does (src/likelihoods/system.jl:116-118, 156-170), calls the reference's own
# `ln_like(obs, ctx)` methods (relative-astrometry.jl:166-253, rv-absolute.jl:172-204, rv-absolute-margin.jl:140-185,
# rv-relative.jl:177-211, prior-observable.jl:78-137) through hand-built observation contexts (src/variables.jl:22-76)
# and prints the difference to the JSON's `ll`; the gradient is checked with ForwardDiff over the same closure.
using Octofitter, OctofitterRadialVelocity, PlanetOrbits, ForwardDiff, JSON
using Octofitter: PlanetObservationContext, SystemObservationContext, ln_like, mjup2msol

col(b, k) = b[k] === nothing ? nothing : Float64.(b[k])
novars = Octofitter.@variables begin end

function build_obs(b)
    k = b["kind"]
    if k == 0
        cols = (; epoch=col(b, "epoch"), ra=col(b, "y1"), dec=col(b, "y2"), σ_ra=col(b, "s1"), σ_dec=col(b, "s2"))
        b["cor"] === nothing || (cols = merge(cols, (; cor=col(b, "cor"))))
        obs = PlanetRelAstromObs(Table(; cols...); name=b["name"], variables=novars)
    elseif k == 1
        cols = (; epoch=col(b, "epoch"), pa=col(b, "y1"), sep=col(b, "y2"), σ_pa=col(b, "s1"), σ_sep=col(b, "s2"))
        b["cor"] === nothing || (cols = merge(cols, (; cor=col(b, "cor"))))
        obs = PlanetRelAstromObs(Table(; cols...); name=b["name"], variables=novars)
    elseif k == 2
        obs = StarAbsoluteRVObs(Table(; epoch=col(b, "epoch"), rv=col(b, "y1"), σ_rv=col(b, "s1")); name=b["name"], variables=novars)
    elseif k == 3
        obs = MarginalizedStarAbsoluteRVObs(Table(; epoch=col(b, "epoch"), rv=col(b, "y1"), σ_rv=col(b, "s1")); name=b["name"], variables=novars)
    elseif k == 4
        obs = PlanetRelativeRVObs(Table(; epoch=col(b, "epoch"), rv=col(b, "y1"), σ_rv=col(b, "s1")); name=b["name"], variables=novars)
    else
        return nothing            # kind 5 (HGCA) needs the catalogue object: compare through the full model instead
    end
    return get(b, "obs_prior", 0) == 1 ? ObsPriorAstromONeil2019(obs) : obs
end

# θ_obs of a table: the kernel-input columns it names (idx_* = 0-based column or -1)
function θ_obs(b, x)
    nt = (;)
    for (k, key) in ((:jitter, "idx_jitter"), (:platescale, "idx_platescale"), (:northangle, "idx_northangle"), (:offset, "idx_offset"))
        b[key] >= 0 && (nt = merge(nt, NamedTuple{(k,)}((x[b[key] + 1],))))
    end
    return nt
end

function total_ll(case, x)
    planets = case["layout"]["planets"]
    blocks = [b for b in case["blocks"] if b["kind"] != 5]
    obs = map(build_obs, blocks)
    g(p, k) = x[p[k] + 1]
    orbits = Tuple(get(p, "basis", 0) == 1 ?
        ThieleInnesOrbit(; A=g(p, "A"), B=g(p, "B"), F=g(p, "F"), G=g(p, "G"), e=g(p, "e"), tp=g(p, "tp"), M=g(p, "M"), plx=g(p, "plx")) :
        Visual{KepOrbit}(; a=g(p, "a"), e=g(p, "e"), i=g(p, "i"), ω=g(p, "w"), Ω=g(p, "W"), tp=g(p, "tp"), M=g(p, "M"), plx=g(p, "plx"))
        for p in planets)
    θ_planets = Tuple(p["mass"] >= 0 ? (; mass=x[p["mass"] + 1]) : (;) for p in planets)
    θ_system = (; planets=NamedTuple{Tuple(Symbol("p$i") for i in eachindex(planets))}(θ_planets))
    # every planet solved at every epoch of every table, tables concatenated in block order
    epochs = reduce(vcat, [col(b, "epoch") for b in blocks]; init=Float64[])
    sols = Tuple([orbitsolve(o, t) for t in epochs] for o in orbits)
    ll = zero(eltype(x)); start = 0
    for (b, o) in zip(blocks, obs)
        ctx = b["planet"] >= 0 ?
            PlanetObservationContext(θ_system, θ_planets[b["planet"] + 1], θ_obs(b, x), orbits, sols, b["planet"] + 1, start) :
            SystemObservationContext(θ_system, θ_obs(b, x), orbits, sols, start)
        ll += ln_like(o, ctx)
        start += length(b["epoch"])
    end
    return ll
end

for f in sort(filter(n -> startswith(n, "case_") && endswith(n, ".json"), readdir(ARGS[1])))
    case = JSON.parsefile(joinpath(ARGS[1], f))
    any(b -> b["kind"] == 5, case["blocks"]) && (println(rpad(f, 32), "skipped (HGCA: needs the catalogue object)"); continue)
    x = Float64.(case["x"])
    ll = total_ll(case, x)
    grad = ForwardDiff.gradient(v -> total_ll(case, v), x)
    rel(a, b) = abs(a - b) / max(abs(b), 1e-300)
    gerr = maximum(abs.(grad .- Float64.(case["grad"])) ./ max.(abs.(Float64.(case["grad"])), 1e-3 * maximum(abs.(case["grad"]))))
    println(rpad(f, 32), "ll = ", ll, "  golden = ", case["ll"], "  rel = ", rel(ll, case["ll"]), "  max grad err = ", gerr)
end
