"""The Julia glue (julia/OctofitterB200.jl) cannot be executed in this image (no julia binary).  What CAN be checked
here: its ABI structs, parsed from the .jl file, have the layout of include/octo_b200.h (through the ctypes mirror, whose
sizes are checked against the header elsewhere); every `ccall` names a symbol the header declares and passes as many
arguments as the C prototype has; the replay of its call sequence (tests/c_abi_julia_replay.c) at least compiles and
agrees with the header on every struct."""
import ctypes as C
import os
import re
import subprocess

import octofitter_jl_b200 as octo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JL = open(os.path.join(ROOT, "julia", "OctofitterB200.jl"), encoding="utf-8").read()
HDR = open(os.path.join(ROOT, "include", "octo_b200.h"), encoding="utf-8").read()


def _jl_structs():
    out = {}
    for m in re.finditer(r"^struct (Octo\w+)\n(.*?)^end", JL, re.S | re.M):
        fields = []
        for ln in m.group(2).splitlines():
            f = re.match(r"\s*(\w+)::([\w{},]+)", ln)
            if f:
                fields.append((f.group(1), f.group(2)))
        out[m.group(1)] = fields
    return out


def _size_align(t):
    if t == "Cdouble":
        return 8, 8
    if t == "Int32":
        return 4, 4
    if t.startswith("Ptr{"):
        return 8, 8
    m = re.match(r"NTuple\{(\w+),(\w+)\}", t)
    if m:
        n = 4 if m.group(1) == "MAXP" else int(m.group(1))
        s, a = _size_align(m.group(2))
        return n * s, a
    raise AssertionError(f"unknown Julia field type {t}")


def _layout(fields):
    off, amax, out = 0, 1, []
    for name, t in fields:
        s, a = _size_align(t)
        off = (off + a - 1) // a * a
        out.append((name, off, s))
        off += s; amax = max(amax, a)
    return out, (off + amax - 1) // amax * amax


def test_julia_structs_have_the_c_layout():
    js = _jl_structs()
    for name, ct in (("OctoConstants", octo.OctoConstants), ("OctoObsBlock", octo.OctoObsBlock), ("OctoLayout", octo.OctoLayout),
                     ("OctoPrior", octo.OctoPrior), ("OctoInputDef", octo.OctoInputDef)):
        lay, size = _layout(js[name])
        assert size == C.sizeof(ct), name
        assert [f for f, _, _ in lay] == [f for f, _ in ct._fields_], name           # same fields, same order
        for f, off, s in lay:
            assert getattr(ct, f).offset == off and getattr(ct, f).size == s, (name, f)
    assert re.search(r"const MAXP = 4\b", JL) and "#define OCTO_MAX_PLANETS 4" in HDR
    v = int(re.search(r"#define OCTO_ABI_VERSION (\d+)", HDR).group(1))
    assert int(re.search(r"const OCTO_ABI_VERSION = (\d+)", JL).group(1)) == v


def _header_prototypes():
    protos = {}
    for m in re.finditer(r"^\s*(?:const\s+)?[\w\*]+\s+\**\s*(octo_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", HDR, re.S | re.M):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    return protos


def test_every_ccall_matches_a_prototype_of_the_header():
    protos = _header_prototypes()
    assert set(protos) == set(octo.EXPORTED_SYMBOLS)
    calls = re.findall(r"ccall\(\(:?(\w+), LIB\), [\w{}]+,\s*\(([^()]*)\)", JL, re.S)
    assert len(calls) >= 12
    seen = set()
    for name, argt in calls:
        if name == "f":          # pt_hmc_run picks :octo_pt_hmc_run / :octo_pt_hmc_run_dist at run time (same signature)
            for real in ("octo_pt_hmc_run", "octo_pt_hmc_run_dist"):
                n = len([a for a in re.split(r",(?![^{]*\})", argt) if a.strip()])
                assert n == protos[real], (real, n, protos[real])
                seen.add(real)
            continue
        assert name in protos, f"ccall to {name}: not declared in include/octo_b200.h"
        n = len([a for a in re.split(r",(?![^{]*\})", argt) if a.strip()])
        assert n == protos[name], (name, n, protos[name])
        seen.add(name)
    for must in ("octo_create", "octo_logp", "octo_logp_grad", "octo_logp_grad_begin", "octo_wait", "octo_set_parameterization",
                 "octo_logpost_grad", "octo_hmc_run", "octo_pt_init", "octo_pt_hmc_run_dist", "octo_destroy", "octo_abi_version"):
        assert must in seen, must


def test_glue_offloads_only_what_the_kernel_implements():
    """a custom trend_function or a non-Visual{KepOrbit} orbit must keep the observation in Julia."""
    assert "function probe_trend(obs, θ_obs)" in JL
    assert re.search(r"StarAbsoluteRVObs\n\s+return \(isnothing\(obs\.gaussian_process\) && !isnothing\(probe_trend", JL)
    assert re.search(r"MarginalizedStarAbsoluteRVObs\n\s+return !isnothing\(probe_trend", JL)
    assert "basis_of(pl)" in JL and "AbsoluteVisual" in JL
    # the model type handed to the samplers is the reference's own
    assert "Octofitter.LogDensityModel(b200_system(system; device); kwargs...)" in JL
    assert "<: Octofitter.AbstractObs" in JL and "function Octofitter.ln_like(o::B200Likelihood, ctx::Octofitter.SystemObservationContext)" in JL


def test_replay_of_the_glue_compiles_and_agrees_on_struct_layout(tmp_path):
    exe = tmp_path / "replay"
    subprocess.run(["gcc", "-O1", "-Wall", "-Werror", "-o", str(exe), os.path.join(ROOT, "tests", "c_abi_julia_replay.c"), "-ldl"], check=True)
    r = subprocess.run([str(exe), octo.LIB_PATH], capture_output=True, text=True)
    # layout mismatch exits with 3 before anything else; without a GPU the first compute call (octo_create) fails with 1
    assert r.returncode in (0, 1), r.stderr
    if r.returncode == 1:
        assert "no CUDA device" in r.stderr or "CUDA" in r.stderr
