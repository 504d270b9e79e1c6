import sys; sys.path.insert(0, '.')

import numpy as np, octofitter_jl_b200 as octo
astrom = octo.PlanetRelAstromObs(octo.Table(epoch=[50000, 50120, 50240], ra=[-505.8, -502.6, -498.2], dec=[-66.9, -37.5, -7.9],
                                            σ_ra=[10.] * 3, σ_dec=[10.] * 3), name="GPI")
b = octo.Planet(name="b", observations=[astrom], variables={
    "a": octo.Uniform(0, 100), "e": octo.Uniform(0, 0.99), "i": octo.Sine(), "ω": octo.UniformCircular(),
    "Ω": octo.UniformCircular(), "θ": octo.UniformCircular(), "tp": octo.θ_at_epoch_to_tperi("θ", 50000)})
system = octo.System(name="Tutoria", companions=[b], variables={
    "M": octo.truncated(octo.Normal(1.2, 0.1), lower=0.1), "plx": octo.truncated(octo.Normal(50.0, 0.02), lower=0.1)})
model = octo.LogDensityModel(system)
chain = octo.octofit(model, np.random.default_rng(0), n_chains=256, adaptation=300, iterations=300)
print(dict(zip(chain["names"], np.median(chain["theta"][-100:].reshape(-1, model.D), axis=0))))
