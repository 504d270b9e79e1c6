# ref_cpu_baseline.jl — times the UNTOUCHED reference on the C2 workload wherever Julia >= 1.10 with
# Octofitter 8.3 + OctofitterRadialVelocity exists (not in the build image).  Emits bench.py's JSON schema.
using Octofitter, OctofitterRadialVelocity, Distributions, BenchmarkTools, JSON, DelimitedFiles
# inputs: the same seeded tables workloads.py generates (export with `python -c "import workloads, json; ..."`)
tab = JSON.parsefile(ARGS[1])            # {"astrom": {...columns...}, "rv": {...}, "theta": [[...], ...]}
astrom = PlanetRelAstromObs(Table(; (Symbol(k) => Float64.(v) for (k, v) in tab["astrom"])...), name="astrom")
rv = StarAbsoluteRVObs(Table(; (Symbol(k) => Float64.(v) for (k, v) in tab["rv"])...), name="rv",
    variables=@variables begin
        offset ~ Normal(150, 100)
        jitter ~ LogUniform(0.1, 100)
    end)
b = Planet(name="b", basis=Visual{KepOrbit}, observations=[astrom], variables=@variables begin
    a ~ LogUniform(1, 100); e ~ Uniform(0, 0.99); i ~ Sine(); ω ~ Uniform(0, 2pi); Ω ~ Uniform(0, 2pi)
    tp ~ Uniform(40000, 60000); mass ~ LogUniform(0.1, 100)
end)
sys = System(name="synthetic", companions=[b], observations=[rv], variables=@variables begin
    M ~ truncated(Normal(1.2, 0.1), lower=0.1); plx ~ truncated(Normal(50, 0.02), lower=0.1)
end)
model = Octofitter.LogDensityModel(sys; verbosity=0)
θ = model.link(Octofitter.guess_starting_position(model, 100)[1])
t1 = @belapsed $(model.∇ℓπcallback)($θ)
E = length(astrom.table.epoch) + length(rv.table.epoch)
println(JSON.json(Dict("impl" => "reference", "metric" => "epoch*chain logp-grad evals/s", "value" => E / t1,
    "unit" => "evals/s", "cpu_baseline" => Dict("kind" => "reference", "cores" => 1, "sample" => "1 chain x $E epochs, @belapsed"))))
