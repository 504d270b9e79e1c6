#!/usr/bin/env python
"""bench.py — epoch·logp-grad evaluations per second of the hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port on the host cores
    python bench.py --sweep                                   # extra: C5 epoch sweep lines (not the contract line)

Workload at N=1: BASELINE.json configs[1] ("C2": 1 planet, 200 astrometry+RV epochs — 100 RA/Dec + 100
star-RV with offset and jitter — x 1024 chains).  A "step" is one value+gradient evaluation of the whole batch
(one leapfrog's worth of work for 1024 chains).  N>1: every rank runs its own 1024 chains (chains are
independent, no data-path collective) => weak scaling; value = all ranks' pairs / max-over-ranks time.

`value`: inputs resident in HBM, one kernel per step, K steps launched back to back between one pair of CUDA
events on the launching stream (barrier + synchronize on both sides); every step reads its own input set from a
pool larger than L2, so no step finds its inputs cached.  The region is measured N_REGIONS times and the median is
reported (all regions are listed).  `roofline.frac` counts EXECUTED FP64 flop (profiles/sass_flops.json).
N > 1 additionally runs BASELINE's C4 — 64 tempered replicas sharded over the ranks with the NCCL swap — and reports it
under `pt_nccl`.  `roofline.kernel_ms_isolated` is the latency of one cold
launch (own event pair, whole L2 flushed before it).  `e2e`: the same metric through the public host API (`LogDensityModel.
ln_like_and_gradient` -> C ABI `octo_logp_grad`) with HOST buffers: pack + H2D + kernel + D2H inside the
timed region, wall clock around K synchronous calls.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# ALGORITHMIC FP64 flop per (epoch x chain) pair — SURVEY.md §8(d)'s per-unit figure (libm-style cost model:
# add/mul = 1, div = 8, sqrt = 8, cbrt = 24, sincos = 48, log = 24): one Kepler solve = 75 + 7*8 + 8 + 24 + 2*48 =
# 259; astrometry projection + chi^2 + adjoint = 49; RV = 45 + div + log = 77.  The roofline's `achieved` uses this.
L2_BYTES = 126 * 1024 * 1024       # B200 L2
N_REGIONS = 11                     # timed regions of K steps each; the median is reported
WORKLOAD = "C2: 1 planet, 100 RA/Dec astrometry + 100 star-RV (offset, jitter) epochs x 1024 chains per GPU"
F_ALG = {"astrom": 308.0, "rv": 336.0, "extra_solve": 259.0}
# EXECUTED FP64 flop per pair of THIS kernel, from ncu SASS counts (DFMA = 2, DMUL/DADD = 1; profiles/r01_*):
# the FP32 Markley starter, the branch-free sincos/rcp and the single sincos per solve make it ~1.8x leaner than
# the algorithmic figure.  Reported next to the roofline as `executed`.
F_EXEC_FALLBACK = {"astrom": 150.0, "astrom_jitter": 305.0, "rv": 143.0, "rv_jitter": 206.0, "rv_margin": 219.0}
FP64_PEAK_FALLBACK_TFLOPS = 37.2     # nominal 148 SM x 64 DFMA/clk x 2 x 1.965 GHz, used only if the probe fails


def executed_flop_table():
    """EXECUTED FP64 flop per pair of this build's epoch loops (2 DFMA + DMUL + DADD), counted from the library's SASS by
    profiles/tools/sass_flops.py (run by __graft_entry__.build()).  This is what `roofline.frac` is made of."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "sass_flops.json")))["kernels"]["thr"]
        return {k: v["flop"] for k, v in d.items()}, "profiles/sass_flops.json (SASS op counts of this build)"
    except Exception:
        return dict(F_EXEC_FALLBACK), "fallback constants (profiles/sass_flops.json missing)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true", help="diagnostic: keep L2 warm between steps (not the contract number)")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def pin_host_thread(local_rank, local_world):
    """Give every rank its own slice of the host cores (all ranks of a box otherwise share one affinity mask and their
    launch threads migrate and contend: the e2e path is host-bound)."""
    try:
        cpus = sorted(os.sched_getaffinity(0))
        per = max(1, len(cpus) // max(1, local_world))
        mine = cpus[local_rank * per:(local_rank + 1) * per] or cpus
        os.sched_setaffinity(0, mine)
        return mine
    except (AttributeError, OSError):
        return None


def fp64_peak(device):
    """Measured FP64 FMA peak (TFLOP/s) — MEASURED_PEAKS.json has no FP64 entry."""
    so = os.path.join(ROOT, "profiles", "tools", "libfp64_peak.so")
    try:
        lib = C.CDLL(so)
        lib.fp64_peak_tflops.restype = C.c_double
        lib.fp64_peak_tflops.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
        ms = C.c_double()
        v = lib.fp64_peak_tflops(device, 2000, 5, C.byref(ms))
        if v > 0:
            return v, "measured (profiles/tools/fp64_peak.cu, DFMA chains, best of 5)"
    except OSError:
        pass
    return FP64_PEAK_FALLBACK_TFLOPS, "fallback (nominal)"


def measured_hbm():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


def flops_per_launch(spec, n_chains, table=None):
    """FP64 flop of one launch: per table, epochs x (kind figure + one extra solve per additional planet)."""
    table = table or F_ALG
    P = len(spec.layout_dict["planets"])
    fl = 0.0
    for b in spec.block_dicts:
        kind = "astrom" if b["kind"] <= 1 else "rv"
        if "rv_jitter" in table:            # the executed-flop table distinguishes the specialised loops
            jit = b.get("idx_jitter", -1) >= 0
            kind = ("astrom_jitter" if jit else "astrom") if b["kind"] <= 1 else ("rv_margin" if b["kind"] == 3 else ("rv_jitter" if jit else "rv"))
        n_solves = P if b["kind"] in (2, 3) else 1 + sum(1 for j, pl in enumerate(spec.layout_dict["planets"])
                                                         if j != b["planet"] and pl.get("mass", -1) >= 0)
        extra = table.get("extra_solve", 0.85 * table[kind])
        fl += len(b["epoch"]) * (table[kind] + (n_solves - 1) * extra)
    return fl * n_chains


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the C2 launch from the committed `ncu --set full` capture of this
    round (profiles/r02_ncu_c2.json, written by profiles/tools/ncu_summary.py --json), per launch; None when absent."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_c2.json")))
        return d["dram_bytes_read"] + d["dram_bytes_write"], "profiles/r02_ncu_c2.json (ncu --set full of the C2 launch)"
    except Exception:
        return None, "no ncu capture committed for this build"


def c4_parallel_tempering(octo, workloads, torch, dist, rank, world, local, rounds=60):
    """BASELINE config C4 across the ranks of this run: 64 tempered replicas block-partitioned over the GPUs.
    (a) host-API swap rounds through libocto_b200's own NCCL all-gather (octo_pt_swap_round), every decision checked
        against octo_pt_decide fed with the values gathered independently over torch.distributed;
    (b) the device-ordered run (octo_pt_hmc_run_dist: resident explorer + ncclAllGather + decision kernel per round, no host
        in the loop), checked against the single-GPU run of all 64 replicas on rank 0 (octo_pt_hmc_run)."""
    R = 64
    nl = R // world
    spec, x = workloads.config("C4")
    model = octo.LogDensityModel(spec, device=local)
    lad = np.linspace(0.0, 1.0, R) ** 2
    pt = octo.ParallelTempering(R, rank=rank, world=world, seed=21, beta=lad, backend="nccl", model=model)
    check = octo.ParallelTempering(R, seed=21, beta=lad, backend="local")
    rng = np.random.default_rng(7)
    ref_all = rng.normal(-30, 3, R)
    ok, t_rounds = True, []
    xs = np.asfortranarray(x[pt.local_slice])
    for rnd in range(rounds):
        xs[:, spec.column("b.a")] *= 1.0 + 1e-4 * np.cos(rnd + np.arange(nl))          # the states move between rounds
        tgt = model.ln_like(xs)
        t0 = time.perf_counter()
        acc = pt.swap_round(ref_all[pt.local_slice], tgt)
        t_rounds.append(time.perf_counter() - t0)
        parts = [torch.empty(nl, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(parts, torch.from_numpy(tgt).cuda())
        acc0 = check.swap_round(ref_all, torch.cat(parts).cpu().numpy())
        ok = ok and np.array_equal(acc, acc0) and np.array_equal(pt.chain_of_replica, check.chain_of_replica)
    pt.close(); model.close()
    # (b) device-ordered, with the standard priors attached
    spec_p, th_p = workloads.one_planet_with_priors(100, 0, R, seed=4)
    model_p = octo.LogDensityModel(spec_p, device=local)
    ptd = octo.ParallelTempering(R, rank=rank, world=world, seed=22, beta=lad, backend="nccl", model=model_p)
    im = np.full(spec_p.D, 1e-4)
    kw = dict(n_iter=1, n_leapfrog=8, step_size=1e-3, inv_mass=im)
    octo.device_parallel_tempering_dist(model_p, ptd, th_p[ptd.local_slice], lad, 3, seed=1, **kw)
    dist.barrier()
    t0 = time.perf_counter()
    res = octo.device_parallel_tempering_dist(model_p, ptd, th_p[ptd.local_slice], lad, rounds, seed=2, **kw)
    t_dist = time.perf_counter() - t0
    same, t_single = True, None
    if rank == 0:
        octo.device_parallel_tempering(model_p, th_p, lad, 3, seed=1, **kw)
        t0 = time.perf_counter()
        one = octo.device_parallel_tempering(model_p, th_p, lad, rounds, seed=2, **kw)
        t_single = time.perf_counter() - t0
        same = bool(np.array_equal(one["swap_counts"], res["swap_counts"]) and np.array_equal(one["cold_trace"], res["cold_trace"])
                    and np.array_equal(one["theta_final"][:nl], res["theta_final"]))
    ptd.close(); model_p.close()
    # (the first round pays NCCL's lazy connection set-up: reported separately)
    tt = torch.tensor([float(np.median(t_rounds[1:])), t_dist, 0.0 if ok else 1.0, t_rounds[0]], dtype=torch.float64, device="cuda")
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank != 0:
        return None
    return {"what": "C4: 64 tempered replicas, %d per GPU on %d GPUs, %d rounds" % (nl, world, rounds),
            "ranks": world, "replicas_per_rank": nl, "rounds": rounds,
            "host_api": {"us_per_round": float(tt[0]) * 1e6, "first_round_us": float(tt[3]) * 1e6, "decisions_match_octo_pt_decide": bool(tt[2] == 0.0),
                         "what": "octo_pt_swap_round: host pairs -> H2D -> ncclAllGather (inside libocto_b200) -> D2H -> decisions; "
                                 "checked every round against octo_pt_decide on values gathered over torch.distributed"},
            "device_ordered": {"us_per_round": float(tt[1]) / rounds * 1e6, "us_per_round_single_gpu_all_replicas": t_single / rounds * 1e6,
                               "identical_to_single_gpu_run": same, "mean_swap_accept": float(np.mean(res["swap_accept"])),
                               "what": "octo_pt_hmc_run_dist: per round {resident explorer: 1 tempered HMC transition x 8 leapfrogs, "
                                       "ncclAllGather of 64 x 2 float64, decision kernel on every rank}, one stream, one "
                                       "synchronisation per run; wall clock of the call, max over ranks"}}


def cpu_baseline(spec, x, threads, budget_s=12.0):
    """The oracle port (value + dual-number gradient) on the host cores, bounded sample of the same workload."""
    import octofitter_jl_b200 as octo
    from oracle import oracle_py
    orc = oracle_py.Oracle(spec.packed, octo.default_constants())
    n = x.shape[0]
    orc.logp_grad(x[: max(1, n // 8)], threads=threads)
    best, reps, t_all = 1e30, 0, time.perf_counter()
    while reps < 3 or (time.perf_counter() - t_all < budget_s and reps < 50):
        t0 = time.perf_counter()
        orc.logp_grad(x, threads=threads)
        best = min(best, time.perf_counter() - t0)
        reps += 1
    one = 1e30
    for _ in range(3):
        t0 = time.perf_counter()
        orc.logp_grad(x, threads=1)
        one = min(one, time.perf_counter() - t0)
    return {"value": n * spec.total_epochs / best, "unit": "epoch*chain logp-grad evals/s", "cores": threads,
            "kind": "port", "sample": f"full batch ({n} chains x {spec.total_epochs} epochs), best of {reps} passes",
            "value_1_thread": n * spec.total_epochs / one}


def run_reference(args):
    """CPU arm: the reference's algorithm as restated by oracle/ (kind 'port' — no Julia in this image), all host
    threads, same workload/metric.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import workloads
    from oracle import oracle_py
    import octofitter_jl_b200 as octo
    spec, x = workloads.config("C2")
    threads = oracle_py.max_threads()
    orc = oracle_py.Oracle(spec.packed, octo.default_constants())
    warm = max(3, args.warmup)
    for _ in range(min(warm, 20)):           # (a CPU pass needs no more warm-up than that; the line reports the flag's value)
        orc.logp_grad(x, threads=threads)
    steps = max(1, min(args.steps, 2000))
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.logp_grad(x, threads=threads)
    dt = (time.perf_counter() - t0) / steps
    pairs = x.shape[0] * spec.total_epochs
    v = pairs / dt
    line = {"impl": "reference", "metric": "epoch*chain logp-grad evals/s", "value": v, "unit": "evals/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "chains_per_gpu": int(x.shape[0]), "epochs": int(spec.total_epochs), "n_in": int(spec.n_in),
                       "note": "CPU arm: rank 0 only, one batch of 1024 chains per step on all host threads"},
            "cpu_baseline": {"value": v, "unit": "evals/s", "cores": threads, "kind": "port",
                             "sample": f"{steps} passes over the full batch ({x.shape[0]} chains x {spec.total_epochs} epochs)"},
            "e2e": {"value": v, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def time_device(model, d_in, d_ll, d_g, n, steps, warmup, torch, flush, grad=True):
    st = torch.cuda.current_stream()
    gp = d_g.data_ptr() if grad else 0
    for _ in range(warmup):
        model.enqueue_device(d_in.data_ptr(), n, n, d_ll.data_ptr(), gp, st.cuda_stream)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        if flush is not None:
            flush.zero_()                      # > L2 (126 MB): the next step starts from a cold L2
        a.record(st)
        model.enqueue_device(d_in.data_ptr(), n, n, d_ll.data_ptr(), gp, st.cuda_stream)
        b.record(st)
    torch.cuda.synchronize()
    return np.array([a.elapsed_time(b) for a, b in ev])      # ms


def time_stream(model, d_in_all, d_ll_all, d_g_all, n, steps, warmup, torch, flush, first=0, barrier=None):
    """The timed region of the contract: `steps` launches back to back on the launching stream between ONE pair of
    CUDA events.  Every launch reads its own input set and writes its own output set; the pool of sets is larger than
    L2 and L2 is flushed once before the region, so no step finds its inputs cached (observation tables and code
    stay resident, as they do across a sampler's evaluations)."""
    st = torch.cuda.current_stream()
    n_sets = d_in_all.shape[0]
    # raw pointers of every set, computed once: the timed loop is one C call per step (tensor indexing per step would
    # make the host, not the GPU, the bottleneck of a 14 us step)
    base = (d_in_all.data_ptr(), d_ll_all.data_ptr(), d_g_all.data_ptr())
    stride = (d_in_all[0].numel() * 8, d_ll_all[0].numel() * 8, d_g_all[0].numel() * 8)
    ptrs = [tuple(b0 + ((first + k) % n_sets) * s0 for b0, s0 in zip(base, stride)) for k in range(warmup + steps)]
    lib, h, sh = model._lib, model._h, st.cuda_stream
    for k in range(warmup):
        model.enqueue_device(ptrs[k][0], n, n, ptrs[k][1], ptrs[k][2], sh)
    torch.cuda.synchronize()
    flush.zero_()
    torch.cuda.synchronize()
    if barrier:
        barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for k in range(warmup, warmup + steps):
        pi, pl, pg = ptrs[k]
        if lib.octo_logp_grad_device(h, pi, n, n, pl, pg, sh):
            raise RuntimeError(lib.octo_last_error().decode())
    b.record(st)
    torch.cuda.synchronize()
    if barrier:
        barrier()
    return a.elapsed_time(b)            # ms for all steps


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    import torch
    import torch.distributed as dist
    import octofitter_jl_b200 as octo
    import workloads

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    pin_host_thread(local, int(os.environ.get("LOCAL_WORLD_SIZE", world)))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    spec, x = workloads.config("C2")          # every rank: its own 1024 chains (seed offset by rank)
    if rank > 0:
        spec, x = workloads.one_planet(100, 100, 1024, seed=2 + 1000 * rank)
    n, n_in, E = x.shape[0], spec.n_in, spec.total_epochs
    model = octo.LogDensityModel(spec, device=local)
    d_in = torch.from_numpy(np.ascontiguousarray(x.T)).cuda()          # [n_in, n] row-major == [n x n_in] column-major
    d_ll = torch.empty(n, dtype=torch.float64, device="cuda")
    d_g = torch.empty((n_in, n), dtype=torch.float64, device="cuda")
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    # pool of input/output sets for the timed region: > 1.5 x L2 in total, one set per step (reused only after the
    # whole pool went by); set 0 is the workload itself, the others are 1e-7-relative rescalings of it
    warm = max(3, args.warmup)
    set_bytes = 8 * n * (2 * n_in + 1)
    n_sets = max(warm + args.steps, int(1.5 * L2_BYTES / set_bytes) + 1) if not args.no_flush else 1
    n_sets = min(n_sets, 8192)
    scale = 1.0 + 1e-7 * torch.arange(n_sets, dtype=torch.float64, device="cuda")
    d_in_all = d_in.unsqueeze(0) * scale[:, None, None]
    d_ll_all = torch.empty((n_sets, n), dtype=torch.float64, device="cuda")
    d_g_all = torch.empty((n_sets, n_in, n), dtype=torch.float64, device="cuda")

    sampler = ClockSampler(local)
    barrier = dist.barrier if world > 1 else None
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    # The timed region of the contract — exactly K steps between one event pair, barrier + synchronize on both sides —
    # measured N_REGIONS times (a 20-step region of a 14 us step is a 0.3 ms sample); the line reports the MEDIAN region
    # (max over ranks per region) and lists them all.
    launches0 = model.kernel_launches
    regions = [time_stream(model, d_in_all, d_ll_all, d_g_all, n, args.steps, warm if r == 0 else 3, torch, flush,
                           first=r * (args.steps + 3), barrier=barrier) for r in range(N_REGIONS)]
    launches = (model.kernel_launches - launches0 - warm - 3 * (N_REGIONS - 1)) // N_REGIONS
    torch.cuda.synchronize()
    if world > 1:
        rt = torch.tensor(regions, dtype=torch.float64, device="cuda")
        dist.all_reduce(rt, op=dist.ReduceOp.MAX)
        regions = [float(v) for v in rt]
    t_dev = float(np.median(regions)) * 1e-3
    # secondary: every launch alone between its own event pair, L2 flushed before each (latency of one cold call)
    ms = time_device(model, d_in, d_ll, d_g, n, max(20, args.steps // 4), 3, torch, None if args.no_flush else flush)
    ms_val = time_device(model, d_in, d_ll, d_g, n, max(20, args.steps // 4), 3, torch, None if args.no_flush else flush, grad=False)

    # end-to-end through the public host API: inputs in pinned host memory, outputs read back every step
    x_pin = model.pinned_empty(x.shape); x_pin[...] = x
    out = (model.pinned_empty(n), model.pinned_empty((n, n_in)))
    for _ in range(max(3, min(args.warmup, 10))):
        model.ln_like_and_gradient(x_pin, out=out)
    if world > 1:
        dist.barrier()
    # (host-timed regions of K steps are a few milliseconds: like the device regions, N_REGIONS of them, median reported)
    def host_regions(fn):
        ts = []
        for _ in range(N_REGIONS):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        return float(np.median(ts)), ts
    def sync_steps():
        global_out = None
        for _ in range(args.steps):
            global_out = model.ln_like_and_gradient(x_pin, out=out)
        return global_out
    ll_h, g_h = sync_steps()
    t_e2e_sync, _ = host_regions(sync_steps)
    # the same steps through the asynchronous halves of the call (octo_logp_grad_begin / octo_wait), DEPTH independent
    # evaluations in flight: every step still copies its inputs host -> device and its (ll, gradient) device -> host
    # inside the timed region; the copies of one step overlap the kernel of another
    DEPTH = 3
    slots = [(model.pinned_empty(x.shape), (model.pinned_empty(n), model.pinned_empty((n, n_in)))) for _ in range(DEPTH)]
    for xi, _ in slots:
        xi[...] = x
    def pipelined(k_steps):
        pend = []
        for k in range(k_steps):
            xi, oi = slots[k % DEPTH]
            if len(pend) == DEPTH:
                pend.pop(0).wait()
            pend.append(model.ln_like_and_gradient_begin(xi, out=oi))
        for h in pend:
            h.wait()
    pipelined(max(3, min(args.warmup, 10)))
    if world > 1:
        dist.barrier()
    t_e2e, e2e_regions = host_regions(lambda: pipelined(args.steps))
    assert np.array_equal(slots[0][1][0], ll_h) and np.array_equal(slots[(args.steps - 1) % DEPTH][1][1], g_h)
    # the same call with ordinary (pageable) numpy arrays, staged through the library's pinned buffers
    t0 = time.perf_counter()
    for _ in range(args.steps):
        model.ln_like_and_gradient(x)
    t_e2e_pageable = time.perf_counter() - t0
    # the same tables with the reference's standard priors attached: full ℓπ(θ_t), ∇ℓπ(θ_t) on the device
    # (K0 forward + K1 + K0 backward), host θ_t in, host (lp, ∇) out
    spec_p, th_p = workloads.one_planet_with_priors(100, 100, n, seed=2 + 1000 * rank)
    model_p = octo.LogDensityModel(spec_p, device=local)
    thp = model_p.pinned_empty(th_p.shape); thp[...] = th_p
    out_p = (model_p.pinned_empty(n), model_p.pinned_empty((n, spec_p.D)))
    for _ in range(5):
        model_p.ℓπcallback_grad(thp, out=out_p)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        model_p.ℓπcallback_grad(thp, out=out_p)
    t_post = time.perf_counter() - t0
    n0 = model_p._lib.octo_kernel_launches(model_p._h)
    model_p.ℓπcallback_grad(thp, out=out_p)
    post_launches = model_p._lib.octo_kernel_launches(model_p._h) - n0
    # the same log posterior driven by the device-resident HMC explorer (octo_hmc_run): whole run on one stream
    hmc_iters, hmc_leap = 100, 10            # (a 1000-leapfrog call: its fixed ~150 us of allocation, copies and the final sync are 0.15 us per leapfrog)
    im_p = np.full(spec_p.D, 1e-4)
    octo.device_hmc(model_p, th_p, 2, step_size=1e-3, n_leapfrog=hmc_leap, inv_mass=im_p, seed=1, keep_samples=False)
    t0 = time.perf_counter()
    hres = octo.device_hmc(model_p, th_p, hmc_iters, step_size=1e-3, n_leapfrog=hmc_leap, inv_mass=im_p, seed=2, keep_samples=False)
    t_hmc = time.perf_counter() - t0
    # C4-shaped parallel tempering on the same model: 64 replicas, every round 1 tempered transition of 8 leapfrogs +
    # one swap round + re-evaluation, device-resident (octo_pt_hmc_run)
    pt_n, pt_rounds, pt_leap = 64, 200, 8
    lad = np.linspace(0.0, 1.0, pt_n) ** 3
    octo.device_parallel_tempering(model_p, th_p[:pt_n], lad, 2, n_iter=1, n_leapfrog=pt_leap, step_size=1e-3, inv_mass=im_p, seed=3)
    t0 = time.perf_counter()
    pres = octo.device_parallel_tempering(model_p, th_p[:pt_n], lad, pt_rounds, n_iter=1, n_leapfrog=pt_leap, step_size=1e-3, inv_mass=im_p, seed=4)
    t_pt = time.perf_counter() - t0
    clocks = sampler.stop()

    if world > 1:
        tt = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_e2e = float(tt[0])
    pt_nccl = c4_parallel_tempering(octo, workloads, torch, dist, rank, world, local) if world > 1 else None
    # sanity: device path and host path agree (set 0 of the pool is the workload itself)
    if not os.environ.get("OCTO_B200_LIB"):
        # (d_ll was last written by the value-only kernel: its own instantiation, same sums to the last bit or two)
        assert np.allclose(d_ll.cpu().numpy(), ll_h, rtol=1e-13, atol=0), "device-resident and host-API results differ"
        assert np.array_equal(d_ll_all[0].cpu().numpy(), ll_h), "timed-region results differ from the host API's"

    if rank == 0:
        pairs_step = n * E * world
        value = pairs_step * args.steps / t_dev
        kern_ms = t_dev / args.steps * 1e3          # average launch duration over the timed region
        peak, peak_how = fp64_peak(local)
        f_exec, f_exec_how = executed_flop_table()
        fl_exec = flops_per_launch(spec, n, f_exec)             # EXECUTED FP64 flop of one launch: what the roofline fraction is
        fl_model = flops_per_launch(spec, n)                    # SURVEY's libm-style planning figure, secondary
        achieved = fl_exec / (kern_ms * 1e-3) / 1e12
        hbm_peak, hbm_how = measured_hbm()
        traffic, traffic_how = ncu_traffic()
        alg_bytes = 8.0 * (n * n_in * 2 + n) + 8.0 * 6 * E
        line = {
            "metric": "epoch*chain logp-grad evals/s", "value": value, "unit": "evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": t_dev / args.steps * 1e3,
            "regions": {"n": N_REGIONS, "ms_per_step": [r / args.steps for r in regions], "reported": "median",
                        "what": "each region = exactly K steps between one CUDA-event pair (barrier + synchronize on both sides, max "
                                "over ranks); consecutive steps are INDEPENDENT evaluations (own input and output sets) launched back "
                                "to back on one stream - the throughput of a batch of independent evaluations, not the latency of a "
                                "dependent chain (that is hmc_device.us_per_leapfrog)"},
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "chains_per_gpu": n, "epochs": E, "n_in": n_in, "launch_geometry": list(model.launch_geometry(n)),
                       "l2": "inputs larger than L2: every timed step reads its own input set and writes its own output set "
                             "(pool of %d sets, %.0f MB > 126 MB L2; one 256 MiB flush before the timed region); the 9.6 KB "
                             "observation tables and the code stay cached, as across a sampler's evaluations" % (n_sets, n_sets * set_bytes / 1e6),
                       "timing": "one CUDA event pair on the launching stream around the K back-to-back launches, barrier + "
                                 "synchronize on both sides, max over ranks; value = pairs x K / region"},
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "what": "EXECUTED FP64 flop of one launch (2 DFMA + DMUL + DADD per pair, counted from this build's SASS) / "
                                 "the kernel's average launch duration over the timed region / measured DFMA peak",
                         "flop_per_pair": {k: f_exec[k] for k in ("astrom", "rv_jitter") if k in f_exec}, "flop_source": f_exec_how,
                         "traffic": traffic, "traffic_source": traffic_how, "peak_source": peak_how,
                         "frac_cost_model": fl_model / (kern_ms * 1e-3) / 1e12 / peak,
                         "cost_model": {"tflops": fl_model / (kern_ms * 1e-3) / 1e12, "flop_per_pair": F_ALG,
                                        "what": "SURVEY.md 8(d) planning weights (libm-style: sincos = 48, div = 8, ...): flop the "
                                                "kernel does NOT execute; kept as the secondary figure"},
                         "kernel_ms": kern_ms,
                         "kernel_ms_isolated": {"mean": float(np.mean(ms)), "min": float(ms.min()),
                                                "what": "each launch alone between its own event pair, L2 (code and tables "
                                                        "included) flushed before it; an empty kernel measures 6.1 us this way"},
                         "hbm": {"achieved_gbs": alg_bytes / (kern_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                                 "peak_source": hbm_how, "algorithmic_bytes": alg_bytes}},
            "e2e": {"value": pairs_step * args.steps / t_e2e, "unit": "evals/s",
                    "h2d_bytes_per_step": n * n_in * 8, "d2h_bytes_per_step": n * (n_in + 1) * 8,
                    "ms_per_step": t_e2e / args.steps * 1e3,
                    "ms_per_step_synchronous_call": t_e2e_sync / args.steps * 1e3,
                    "ms_per_step_pageable_host_arrays": t_e2e_pageable / args.steps * 1e3,
                    "in_flight": DEPTH,
                    "regions_ms_per_step": [t / args.steps * 1e3 for t in e2e_regions],
                    "api": "LogDensityModel.ln_like_and_gradient_begin(pinned host ndarray, out=pinned).wait() -> C ABI "
                           "octo_logp_grad_begin / octo_wait, %d independent evaluations in flight: every step moves its own inputs host -> device "
                           "and its (ll, gradient) device -> host inside the timed region — page-locked buffers of this size are read and "
                           "written by the kernel in place over PCIe (no separate copy launches; OCTO_B200_ZEROCOPY_MAX=0 restores the "
                           "H2D copy); `ms_per_step_synchronous_call` is one blocking octo_logp_grad per step (the latency of a single call)" % DEPTH},
            "value_only": {"what": "K1v, logp without gradient (Pigeons slice sampler / prior search), same workload, device-resident",
                       "value": n * E * world / (float(np.mean(ms_val)) * 1e-3), "unit": "evals/s", "ms_per_step": float(np.mean(ms_val))},
        "logpost_e2e": {"what": "full log-posterior + gradient w.r.t. the unconstrained vector (priors, bijectors, UniformCircular, "
                                "θ_at_epoch_to_tperi on device), same tables, D = %d, pinned host buffers; %d launch(es) per step" % (spec_p.D, post_launches),
                        "value": pairs_step * args.steps / t_post, "unit": "evals/s", "ms_per_step": t_post / args.steps * 1e3},
        "hmc_device": {"what": "octo_hmc_run on the same model: %d transitions x %d leapfrogs for %d chains, every launch on one "
                               "stream, one synchronisation at the end (wall clock of the call, copies included)" % (hmc_iters, hmc_leap, n),
                       "us_per_leapfrog": t_hmc / (hmc_iters * hmc_leap) * 1e6,
                       "value": pairs_step / world * hmc_iters * hmc_leap / t_hmc, "unit": "evals/s", "accept_rate": hres["accept_rate"]},
        "pt_device": {"what": "octo_pt_hmc_run: %d replicas (C4's count) on the same tables, %d rounds of {1 tempered HMC transition x %d "
                              "leapfrogs, even-odd swap round, re-evaluation}, one stream, one synchronisation" % (pt_n, pt_rounds, pt_leap),
                      "us_per_round": t_pt / pt_rounds * 1e6, "mean_swap_accept": float(np.mean(pres["swap_accept"]))},
        "gpu_launches": int(launches),
        **({"pt_nccl": pt_nccl} if pt_nccl else {}),
            "clocks": clocks,
        }
        if not args.no_cpu_baseline:
            from oracle import oracle_py
            line["cpu_baseline"] = cpu_baseline(spec, x, oracle_py.max_threads())
        print(json.dumps(line))
        if args.sweep:
            sweep(octo, workloads, torch, peak, local)
    model.close()
    model_p.close()
    if world > 1:
        dist.destroy_process_group()


def sweep(octo, workloads, torch, peak, device):
    """C5: epoch sweep x 4096 chains, device-resident, written to gpurun_out/sweep.jsonl."""
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out = open(os.path.join(ROOT, "gpurun_out", "sweep.jsonl"), "w")
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for k, E in enumerate([10, 32, 100, 316, 1000, 3162, 10000, 31623, 100000]):
        spec, x = workloads.one_planet(E, 0, 4096, seed=5 + k)
        model = octo.LogDensityModel(spec, device=device)
        n, n_in = x.shape
        d_in = torch.from_numpy(np.ascontiguousarray(x.T)).cuda()
        d_ll = torch.empty(n, dtype=torch.float64, device="cuda")
        d_g = torch.empty((n_in, n), dtype=torch.float64, device="cuda")
        steps = 20 if E <= 10000 else 5
        ms = time_device(model, d_in, d_ll, d_g, n, steps, 3, torch, flush)
        kms = float(np.mean(ms))
        fl = flops_per_launch(spec, n, executed_flop_table()[0])          # EXECUTED flop (SASS counts), like roofline.frac
        fl_model = flops_per_launch(spec, n)                               # SURVEY's planning weights: secondary
        rec = {"epochs": E, "chains": n, "kernel_ms": kms, "evals_per_s": n * E / (kms * 1e-3),
               "fp64_tflops": fl / (kms * 1e-3) / 1e12, "frac_fp64_peak": fl / (kms * 1e-3) / 1e12 / peak,
               "frac_cost_model": fl_model / (kms * 1e-3) / 1e12 / peak, "geometry": list(model.launch_geometry_full(n))}
        out.write(json.dumps(rec) + "\n"); out.flush()
        print("sweep", json.dumps(rec), file=sys.stderr)
        model.close()
    # the other BASELINE.json configs (parity-test cases; timed here for the record, not bench lines)
    for name in ("C1", "C3", "C4"):
        spec, x = workloads.config(name)
        model = octo.LogDensityModel(spec, device=device)
        n, n_in = x.shape
        d_in = torch.from_numpy(np.ascontiguousarray(x.T)).cuda()
        d_ll = torch.empty(n, dtype=torch.float64, device="cuda")
        d_g = torch.empty((n_in, n), dtype=torch.float64, device="cuda")
        ms = time_device(model, d_in, d_ll, d_g, n, 50, 5, torch, flush)
        t0 = time.perf_counter()
        for _ in range(50):
            model.ln_like_and_gradient(x)
        t_e2e = (time.perf_counter() - t0) / 50
        rec = {"config": name, "chains": n, "epochs": spec.total_epochs, "planets": len(spec.layout_dict["planets"]),
               "kernel_ms": float(np.mean(ms)), "evals_per_s": n * spec.total_epochs / (float(np.mean(ms)) * 1e-3),
               "e2e_ms_pageable": t_e2e * 1e3, "geometry": list(model.launch_geometry(n))}
        out.write(json.dumps(rec) + "\n")
        print("sweep", json.dumps(rec), file=sys.stderr)
        model.close()
    out.close()


if __name__ == "__main__":
    main()
