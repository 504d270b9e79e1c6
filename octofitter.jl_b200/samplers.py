"""Replica-batched explorers over the device log-posterior.

The reference samples ONE chain at a time (AdvancedHMC NUTS, src/sampling.jl:412-423; Pigeons SliceSampler per
replica, ext/OctofitterPigeonsExt:70-72).  On the GPU the natural unit is a batch of chains advanced in lockstep:
every leapfrog step is one `ℓπcallback_grad` call for all chains.  These drivers are deliberately small — the
samplers stay the caller's business; they exist so the path can be exercised end to end.
"""
from __future__ import annotations

import numpy as np


def diagonal_metric(model, theta, h=1e-5):
    """Diagonal inverse mass from the curvature at `theta` (D,): 1 / |∂²ℓπ/∂θ_j²| by central differences of the
    device gradient — all 2D probes are one batched call.  (The reference adapts a dense metric with Stan's windowed
    adaptation, src/sampling.jl:335-345; this is the minimal stand-in for tests and examples.)"""
    th = np.asarray(theta, dtype=np.float64)
    D = th.shape[0]
    probes = np.repeat(th[None, :], 2 * D, axis=0)
    for j in range(D):
        probes[2 * j, j] += h; probes[2 * j + 1, j] -= h
    _, g = model.ℓπcallback_grad(np.asfortranarray(probes))
    hjj = np.array([-(g[2 * j, j] - g[2 * j + 1, j]) / (2 * h) for j in range(D)])
    return 1.0 / np.maximum(np.abs(hjj), 1e-8)


def batched_hmc(model, theta0, n_iter, *, step_size=0.02, n_leapfrog=16, rng=None, inv_mass=None):
    """Static-trajectory HMC on `model.ℓπcallback_grad` for all chains at once.

    theta0: (n_chains, D) unconstrained start.  Returns dict(theta [n_iter, n_chains, D], logpost [n_iter, n_chains],
    accept_rate, n_gradient_calls).  inv_mass: optional diagonal inverse mass (D,).
    """
    rng = np.random.default_rng() if rng is None else rng
    th = np.array(theta0, dtype=np.float64, order="F")
    n, D = th.shape
    im = np.ones(D) if inv_mass is None else np.asarray(inv_mass, dtype=np.float64)
    lp, g = model.ℓπcallback_grad(th)
    lp, g = lp.copy(), g.copy()
    out_th = np.empty((n_iter, n, D)); out_lp = np.empty((n_iter, n))
    acc_total, calls = 0.0, 1
    for it in range(n_iter):
        p = rng.standard_normal((n, D)) / np.sqrt(im)
        h0 = -lp + 0.5 * np.sum(p * p * im, axis=1)
        q, gq, lq = th.copy(), g.copy(), lp.copy()
        p = p + 0.5 * step_size * gq
        for k in range(n_leapfrog):
            q = q + step_size * p * im
            lq, gq = model.ℓπcallback_grad(np.asfortranarray(q))
            calls += 1
            gq = np.where(np.isfinite(lq)[:, None], gq, 0.0)
            p = p + (step_size if k < n_leapfrog - 1 else 0.5 * step_size) * gq
        with np.errstate(over="ignore", invalid="ignore"):
            h1 = -lq + 0.5 * np.sum(p * p * im, axis=1)
        dh = h0 - h1
        accept = np.isfinite(lq) & (np.log(rng.uniform(size=n)) < dh)
        th[accept] = q[accept]; g[accept] = gq[accept]; lp[accept] = lq[accept]
        acc_total += accept.mean()
        out_th[it] = th; out_lp[it] = lp
    return {"theta": out_th, "logpost": out_lp, "accept_rate": acc_total / n_iter, "n_gradient_calls": calls}


def device_hmc(model, theta0, n_iter, *, step_size=0.02, n_leapfrog=16, inv_mass=None, seed=0, keep_samples=True):
    """The same explorer as `batched_hmc`, resident on the device (C ABI `octo_hmc_run`): the whole run — n_iter
    transitions x n_leapfrog leapfrogs for all chains — is enqueued on one stream and synchronised once.  Returns the
    same dict as `batched_hmc` plus `theta_final`, `logpost_final` and the per-chain `accept` rates.  Counter-based
    randomness: the result is a pure function of (theta0, step_size, n_leapfrog, inv_mass, seed)."""
    import ctypes as C
    th = np.array(theta0, dtype=np.float64, order="F")
    n, D = th.shape
    im = None if inv_mass is None else np.ascontiguousarray(inv_mass, dtype=np.float64)
    samples = np.empty((n_iter, D, n)) if keep_samples else None
    lps = np.empty((n_iter, n)) if keep_samples else None
    th_f = np.empty((n, D), order="F"); lp_f = np.empty(n); acc = np.empty(n)
    p = lambda a: None if a is None else a.ctypes.data
    model._check(model._lib.octo_hmc_run(model._h, th.ctypes.data, n, n, int(n_iter), int(n_leapfrog), float(step_size), p(im),
                                         C.c_uint64(int(seed)), p(samples), p(lps), th_f.ctypes.data, lp_f.ctypes.data, acc.ctypes.data))
    return {"theta": None if samples is None else np.ascontiguousarray(samples.transpose(0, 2, 1)), "logpost": lps,
            "theta_final": th_f, "logpost_final": lp_f, "accept": acc, "accept_rate": float(acc.mean()),
            "n_gradient_calls": n_iter * n_leapfrog + 1}


def device_parallel_tempering(model, theta0, ladder, n_rounds, *, n_iter=1, n_leapfrog=8, step_size=0.02, inv_mass=None, seed=0):
    """Device-resident parallel tempering (C ABI `octo_pt_hmc_run`): chain c starts on rung c of `ladder`; every round
    is n_iter tempered HMC transitions for all chains, one deterministic even-odd swap round and a re-evaluation at the
    new weights, all enqueued on one stream.  Returns final states, tempered log posterior, raw ln_like, weight and rung
    of every chain, swap acceptance rate per adjacent pair and the trace of the chain on the last rung."""
    import ctypes as C
    th = np.array(theta0, dtype=np.float64, order="F")
    n, D = th.shape
    lad = np.ascontiguousarray(ladder, dtype=np.float64)
    assert lad.shape == (n,)
    im = None if inv_mass is None else np.ascontiguousarray(inv_mass, dtype=np.float64)
    th_f = np.empty((n, D), order="F"); lp = np.empty(n); ll = np.empty(n); beta = np.empty(n)
    rung = np.empty(n, dtype=np.int32); swaps = np.empty(n - 1); cold = np.empty((n_rounds, D)); acc = np.empty(n)
    p = lambda a: None if a is None else a.ctypes.data
    model._check(model._lib.octo_pt_hmc_run(model._h, th.ctypes.data, n, n, lad.ctypes.data, int(n_rounds), int(n_iter),
                                            int(n_leapfrog), float(step_size), p(im), C.c_uint64(int(seed)), th_f.ctypes.data,
                                            lp.ctypes.data, ll.ctypes.data, beta.ctypes.data, rung.ctypes.data, swaps.ctypes.data,
                                            cold.ctypes.data, acc.ctypes.data))
    # pair i is proposed on the rounds of its parity
    proposals = np.array([(n_rounds + (1 - (i & 1))) // 2 for i in range(n - 1)], dtype=np.float64)
    return {"theta_final": th_f, "logpost_tempered": lp, "loglike": ll, "beta": beta, "rung": rung,
            "swap_accept": swaps / np.maximum(proposals, 1.0), "swap_counts": swaps, "cold_trace": cold, "accept": acc}


def device_parallel_tempering_dist(model, pt, theta0_local, ladder_all, n_rounds, *, n_iter=1, n_leapfrog=8, step_size=0.02,
                                   inv_mass=None, seed=0):
    """`device_parallel_tempering` with the ladder sharded over the ranks of `pt` (an `octo.ParallelTempering` with
    backend "nccl" — one process per GPU — or "local"): C ABI `octo_pt_hmc_run_dist`.  Collective: every rank calls it
    with its own replicas (chains [rank n_local, (rank + 1) n_local) of the R) and the same full ladder.  Per round one
    resident-explorer launch, one ncclAllGather of R x 2 float64 and one decision kernel, all in stream order.  Returns
    this rank's chains plus the (replicated) swap statistics and the trace of the last rung."""
    import ctypes as C
    th = np.array(theta0_local, dtype=np.float64, order="F")
    n, D = th.shape
    lad = np.ascontiguousarray(ladder_all, dtype=np.float64)
    R = n * pt.world
    assert lad.shape == (R,) and n == pt.n_local
    im = None if inv_mass is None else np.ascontiguousarray(inv_mass, dtype=np.float64)
    th_f = np.empty((n, D), order="F"); lp = np.empty(n); ll = np.empty(n); beta = np.empty(n)
    rung = np.empty(n, dtype=np.int32); swaps = np.empty(R - 1); cold = np.empty((n_rounds, D)); acc = np.empty(n)
    p = lambda a: None if a is None else a.ctypes.data
    model._check(model._lib.octo_pt_hmc_run_dist(model._h, th.ctypes.data, n, n, lad.ctypes.data, int(n_rounds), int(n_iter),
                                                 int(n_leapfrog), float(step_size), p(im), C.c_uint64(int(seed)), th_f.ctypes.data,
                                                 lp.ctypes.data, ll.ctypes.data, beta.ctypes.data, rung.ctypes.data, swaps.ctypes.data,
                                                 cold.ctypes.data, acc.ctypes.data))
    proposals = np.array([(n_rounds + (1 - (i & 1))) // 2 for i in range(R - 1)], dtype=np.float64)
    return {"theta_final": th_f, "logpost_tempered": lp, "loglike": ll, "beta": beta, "rung": rung,
            "swap_accept": swaps / np.maximum(proposals, 1.0), "swap_counts": swaps, "cold_trace": cold, "accept": acc}


def hmc_random(model, seed, it, chain, D):
    """(z[D], u): the standard normals and the accept-step uniform `octo_hmc_run` uses for transition `it` of `chain`."""
    import ctypes as C
    z = np.empty(D); u = C.c_double()
    model._lib.octo_hmc_random(C.c_uint64(int(seed)), int(it), int(chain), int(D), z.ctypes.data, C.byref(u))
    return z, u.value


def octofit(model, rng=None, *, n_chains=256, adaptation=300, iterations=300, target_accept=0.8, n_leapfrog=12,
            n_init=100_000, windows=6, seed=None, verbosity=0):
    """`octofit(model; adaptation, iterations)` (src/sampling.jl:300-470) for a batch of chains on the device.

    The reference runs one NUTS chain with Stan-style windowed adaptation (AdvancedHMC `StanHMCAdaptor`: dual-averaging
    step size + a metric estimated in expanding windows).  Here `n_chains` chains of static-trajectory HMC run in
    lockstep on the device (`octo_hmc_run`): start = best of `n_init` prior draws (guess_starting_position,
    src/initialization.jl:14-66) plus a small scatter; adaptation = `windows` expanding windows, after each of which the
    diagonal inverse mass is set to the pooled per-coordinate variance of ALL the window's draws and the step size is
    rescaled towards `target_accept`, then a short terminal window that re-tunes the step size at the final metric (as
    Stan's does); then `iterations` sampling transitions.  Returns a dict like the reference's chain:
    `theta_t` / `theta` (natural space) [iterations, n_chains, D], `logpost`, `names`, and an `info` dict with the
    acceptance rate, the adapted step size / inverse mass and the gradient-call count."""
    rng = np.random.default_rng() if rng is None else rng
    seed = int(rng.integers(1 << 62)) if seed is None else int(seed)
    D = model.D
    params, _ = model.guess_starting_position(rng, N=n_init)
    start = model.link(params)
    inv_mass = diagonal_metric(model, start)
    th = np.asfortranarray(start[None, :] + 0.1 * np.sqrt(inv_mass)[None, :] * rng.standard_normal((n_chains, D)))
    eps, calls = 0.1, 0
    # expanding windows (Stan: each twice the previous); every window ends with a metric and step-size update
    sizes = np.maximum(1, (adaptation * 2.0 ** np.arange(windows) / (2.0 ** windows - 1)).astype(int))
    def tune(n_it, tag, sub):
        nonlocal th, eps, calls
        r = device_hmc(model, th, n_it, step_size=eps, n_leapfrog=n_leapfrog, inv_mass=inv_mass, seed=seed + 1000 * tag + sub)
        calls += r["n_gradient_calls"]
        th = r["theta_final"]
        eps *= float(np.clip(np.exp(1.5 * (r["accept_rate"] - target_accept)), 0.5, 2.0))
        if verbosity >= 2:
            print(f"adapt window {tag}.{sub}: {n_it} it, accept {r['accept_rate']:.2f}, step {eps:.3g}")
        return r
    for k, w in enumerate(sizes):
        # a few short runs inside the window tune the step size at the current metric; the metric update at the end of
        # the window pools the draws of ALL of them
        pooled = [tune(max(1, int(w) // 3), k, sub)["theta"].reshape(-1, D) for sub in range(3)]
        draws = np.concatenate(pooled, axis=0)
        var = draws.var(axis=0)
        if np.all(np.isfinite(var)) and np.all(var > 0):
            # Stan's regularisation of the variance estimate
            nw = draws.shape[0]
            inv_mass = (nw / (nw + 5.0)) * var + 1e-3 * (5.0 / (nw + 5.0))
    # terminal fast window (Stan's): the step size was tuned for the previous metric — re-tune it at the final one
    for sub in range(3):
        tune(max(2, int(sizes[0])), windows, sub)
    r = device_hmc(model, th, iterations, step_size=eps, n_leapfrog=n_leapfrog, inv_mass=inv_mass, seed=seed + 999_983)
    calls += r["n_gradient_calls"]
    theta_t = r["theta"]
    nat = model.invlink(theta_t.reshape(-1, D)).reshape(theta_t.shape)
    return {"theta_t": theta_t, "theta": nat, "logpost": r["logpost"], "names": model.spec.theta_names,
            "info": {"sampler": "device_hmc", "n_chains": n_chains, "adaptation": int(sizes.sum()) + 3 * max(2, int(sizes[0])), "iterations": iterations,
                     "accept_rate": r["accept_rate"], "step_size": eps, "inv_mass": inv_mass, "n_leapfrog": n_leapfrog,
                     "n_gradient_calls": calls, "seed": seed}}


def batched_parallel_tempering(model, model_ref_logp, pt, theta0, n_rounds, *, step_size=0.02, n_leapfrog=8, rng=None,
                               inv_mass=None):
    """Tempered HMC explorer + deterministic even-odd swaps (octo.ParallelTempering) for the LOCAL replicas of a rank.

    model: parameterised LogDensityModel (target ℓπ, prior included); model_ref_logp(θ_t) -> (ℓ_ref, ∇ℓ_ref) of the
    tempering reference (the prior-only model, ext/OctofitterPigeonsExt:61-67).  Replica r explores
    (1-β) ℓ_ref + β ℓ_target with the β it currently holds; swaps exchange β indices, not states.
    """
    rng = np.random.default_rng() if rng is None else rng
    th = np.array(theta0, dtype=np.float64, order="F")
    n, D = th.shape
    assert n == pt.n_local
    im = np.ones(D) if inv_mass is None else np.asarray(inv_mass, dtype=np.float64)
    swaps = []
    for rnd in range(n_rounds):
        beta = pt.local_betas()[:, None]

        def tempered(q):
            lt, gt = model.ℓπcallback_grad(np.asfortranarray(q))
            lr, gr = model_ref_logp(q)
            ok = np.isfinite(lt)
            lt = np.where(ok, lt, -np.inf); gt = np.where(ok[:, None], gt, 0.0)
            return (1 - beta[:, 0]) * lr + beta[:, 0] * lt, (1 - beta) * gr + beta * gt, lr, lt
        lp, g, lr, lt = tempered(th)
        p = rng.standard_normal((n, D)) / np.sqrt(im)
        h0 = -lp + 0.5 * np.sum(p * p * im, axis=1)
        q, gq = th.copy(), g.copy()
        p = p + 0.5 * step_size * gq
        for k in range(n_leapfrog):
            q = q + step_size * p * im
            lq, gq, lrq, ltq = tempered(q)
            gq = np.where(np.isfinite(lq)[:, None], gq, 0.0)
            p = p + (step_size if k < n_leapfrog - 1 else 0.5 * step_size) * gq
        with np.errstate(over="ignore", invalid="ignore"):
            h1 = -lq + 0.5 * np.sum(p * p * im, axis=1)
        accept = np.isfinite(lq) & (np.log(rng.uniform(size=n)) < h0 - h1)
        th[accept] = q[accept]; lr = np.where(accept, lrq, lr); lt = np.where(accept, ltq, lt)
        swaps.append(pt.swap_round(lr, lt))
    return {"theta": th, "swap_accept": np.array(swaps)}


def batched_slice_sampler(model, theta0, n_iter, *, rng=None, w=10.0, max_steps=20, n_passes=3, beta=None, keep_samples=True):
    """Replica-batched, VALUE-ONLY slice sampler over the device log posterior: the explorer the reference gives Pigeons
    (`Pigeons.default_explorer(::LogDensityModel) = SliceSampler()`, ext/OctofitterPigeonsExt:70-72; Pigeons' defaults
    w = 10, n_passes = 3).  All replicas move in lockstep: one iteration is `n_passes` sweeps over the D coordinates; a
    coordinate update is Neal (2003) slice sampling — vertical level, interval by stepping out (at most `max_steps`
    expansions; Pigeons doubles, which needs an extra acceptance test per proposal — same invariant distribution), then
    shrinkage — and every evaluation of the interval ends / proposals of ALL replicas is ONE value-only launch (K1v).

    beta: per-replica tempering weights (None = 1): replica r targets prior terms + beta_r * ln_like, the path Pigeons
    tempers along.  Returns theta [n_iter, n, D] (if kept), the final states, their tempered density, raw ln_like and
    the number of batched evaluations."""
    rng = np.random.default_rng() if rng is None else rng
    th = np.array(theta0, dtype=np.float64, order="F")
    n, D = th.shape
    bet = None if beta is None else np.asarray(beta, dtype=np.float64)
    evals = 0

    def density(q):
        """tempered log density and raw ln_like of a [m, D] batch (rows beyond n cycle over the replicas)"""
        nonlocal evals
        evals += 1
        lp = model.ℓπcallback(np.asfortranarray(q))
        if bet is None:
            return lp, None
        ll = model.ln_like_of_theta(np.asfortranarray(q))
        b = np.resize(bet, q.shape[0])
        with np.errstate(invalid="ignore"):
            out = np.where(np.isfinite(lp) & np.isfinite(ll), lp - (1.0 - b) * ll, -np.inf)
        return out, ll
    cur, cur_ll = density(th)
    if not np.all(np.isfinite(cur)):
        raise ValueError("slice sampling needs finite starting densities")
    out = np.empty((n_iter, n, D)) if keep_samples else None
    for it in range(n_iter):
        for _ in range(n_passes):
            for j in range(D):
                y = cur - rng.exponential(size=n)                     # log of the vertical level
                x0 = th[:, j].copy()
                L = x0 - w * rng.uniform(size=n); R = L + w
                # stepping out: both ends of every replica in one launch, until they are outside the slice
                J = np.floor(max_steps * rng.uniform(size=n)).astype(int); K = (max_steps - 1) - J
                growL, growR = np.ones(n, bool), np.ones(n, bool)
                while growL.any() or growR.any():
                    q = np.concatenate([th, th], axis=0)
                    q[:n, j] = L; q[n:, j] = R
                    f, _ = density(q)
                    insideL, insideR = f[:n] > y, f[n:] > y
                    growL &= insideL & (J > 0); growR &= insideR & (K > 0)
                    L = np.where(growL, L - w, L); J = J - growL
                    R = np.where(growR, R + w, R); K = K - growR
                # shrinkage: propose inside [L, R], shrink towards x0 on rejection
                todo = np.ones(n, bool)
                new_x, new_f, new_ll = x0.copy(), cur.copy(), None if cur_ll is None else cur_ll.copy()
                for _shrink in range(200):
                    if not todo.any():
                        break
                    x1 = L + rng.uniform(size=n) * (R - L)
                    q = th.copy(); q[:, j] = np.where(todo, x1, x0)
                    f, fl = density(q)
                    ok = todo & (f > y)
                    new_x = np.where(ok, x1, new_x); new_f = np.where(ok, f, new_f)
                    if fl is not None:
                        new_ll = np.where(ok, fl, new_ll)
                    todo &= ~ok
                    L = np.where(todo & (x1 < x0), x1, L); R = np.where(todo & (x1 >= x0), x1, R)
                th[:, j] = new_x; cur = new_f
                if new_ll is not None:
                    cur_ll = new_ll
        if keep_samples:
            out[it] = th
    return {"theta": out, "theta_final": th, "logdensity": cur, "loglike": cur_ll, "n_evaluations": evals}


def batched_slice_parallel_tempering(model, pt, theta0, n_rounds, *, rng=None, w=10.0, n_passes=3, max_steps=20):
    """The Pigeons-shaped workload with its real explorer: every round, each LOCAL replica of this rank takes one
    slice-sampling iteration at the weight it currently holds (value-only launches, all local replicas per launch),
    then one deterministic even-odd swap round over all ranks (`octo.ParallelTempering.swap_round`: the library's
    ncclAllGather of (l_ref, l_target) when `pt` was built with backend "nccl").  Swaps exchange weights, not states."""
    rng = np.random.default_rng() if rng is None else rng
    th = np.array(theta0, dtype=np.float64, order="F")
    assert th.shape[0] == pt.n_local
    swaps, evals = [], 0
    for rnd in range(n_rounds):
        r = batched_slice_sampler(model, th, 1, rng=rng, w=w, n_passes=n_passes, max_steps=max_steps, beta=pt.local_betas(),
                                  keep_samples=False)
        th = r["theta_final"]; evals += r["n_evaluations"]
        lp = model.ℓπcallback(np.asfortranarray(th)); ll = r["loglike"]
        swaps.append(pt.swap_round(lp - ll, lp))          # l_ref = prior terms, l_target = l_ref + ln_like
    return {"theta": th, "swap_accept": np.array(swaps), "n_evaluations": evals}


def octofit_rejection(model, rng=None, *, draws=100_000, batch=65536, verbosity=0):
    """Rejection sampling from the prior, batched on the device (octofit_rejection, src/sampling.jl:168-258).

    Draw `draws` samples from the priors, evaluate ln_like for all of them (`octo_loglike_theta`, one fused launch per
    batch instead of one CPU call per draw), accept draw i with probability exp(ll_i - max ll).  Returns a dict with
    the accepted natural-space samples (`theta`, [n_accepted, D]), their `loglike` and `logpost`, and the run
    statistics the reference stores in the chain info (`draws`, `n_accepted`, `acceptance_rate`)."""
    rng = np.random.default_rng() if rng is None else rng
    theta = model.sample_priors(rng, draws)
    theta_t = model.link(theta)
    ll = np.empty(draws)
    for lo in range(0, draws, batch):
        ll[lo:lo + batch] = model.ln_like_of_theta(theta_t[lo:lo + batch])
    ll[~np.isfinite(ll)] = -np.inf
    max_ll = ll.max()
    if not np.isfinite(max_ll):
        raise RuntimeError(f"All {draws} prior samples produced non-finite log-likelihoods. Check your model and priors.")
    u = rng.random(draws)
    accepted = np.flatnonzero((ll > -np.inf) & (u < np.exp(ll - max_ll)))
    if accepted.size == 0:
        raise RuntimeError(f"No samples were accepted out of {draws} draws. The posterior may be extremely concentrated "
                           "relative to the prior. Consider increasing `draws` or using a different sampler.")
    rate = accepted.size / draws
    if verbosity >= 1 and rate < 0.001:
        import warnings
        warnings.warn(f"Very low acceptance rate ({100 * rate:.2g}%). Consider HMC for more efficient sampling.")
    return {"theta": theta[accepted], "theta_t": theta_t[accepted], "loglike": ll[accepted],
            "logpost": model.ℓπcallback(theta_t[accepted]), "names": model.spec.theta_names,
            "info": {"sampler": "rejection", "draws": draws, "n_accepted": int(accepted.size), "acceptance_rate": rate}}

