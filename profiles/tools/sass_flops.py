"""Executed FP64 operations per (epoch x chain) pair, counted from the SASS of libocto_b200.so ("the
planning weights are to be replaced by an op counter").

For the 1-planet gradient kernels — throughput instantiation k_kepler_like<true,1,false,…> and latency instantiation
k_kepler_like<true,1,true,…> — find the epoch loop of every table kind (the backward branch whose body holds the most
FP64 instructions), divide by the pairs in flight per iteration (one MUFU.RSQ per Kepler solve) and count DFMA
(2 flop), DMUL, DADD (1 flop each).  Writes profiles/sass_flops.json, which bench.py reads for `roofline.frac`.

The shipped kernels inline the lean loops (octo_kernels.cu, "INLINE OR OUT OF LINE"), where the loops of different tables
cannot be told apart by name.  So the per-kind figures are counted on an ANALYSIS BUILD of the same sources with every
loop out of line (-DOCTO_SEG_ALL_OOL, built here into a temporary file), and then checked against the shipped library:
each lean kind's figure must be found among the backward-branch loops of its 1-planet lean gradient kernels (same
arithmetic, +-2 %); the matching loop is recorded under "shipped_loops" (its instructions per pair are the shipped ones).

    python profiles/tools/sass_flops.py [path/to/libocto_b200.so]
"""
import json
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SO = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "octofitter.jl_b200", "lib", "libocto_b200.so")
OUT = os.path.join(ROOT, "profiles", "sass_flops.json")
FP64 = ("DFMA", "DMUL", "DADD")


def disassemble(so, which="octo_kernels_n1_lean"):
    """SASS of one embedded cubin (build.py names each after its wrapper source)"""
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", which, so], cwd=d, check=True, capture_output=True)
        cub = max((f for f in os.listdir(d) if f.endswith(".cubin")), key=lambda f: os.path.getsize(os.path.join(d, f)))
        return subprocess.run(["nvdisasm", os.path.join(d, cub)], capture_output=True, text=True, check=True).stdout.split("\n")


def analysis_build():
    """the same sources with every table loop out of line, one-planet kernels only"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("b", os.path.join(ROOT, "octofitter.jl_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    out = os.path.join(tempfile.mkdtemp(), "libocto_analysis.so")
    mod.build_variant(out, ["-DOCTO_NPT_ONLY1", "-DOCTO_SEG_ALL_OOL"])
    return out


def sections(lines):
    """kernel name -> lines of its .text section"""
    out, cur = {}, None
    for ln in lines:
        m = re.match(r"//-+ \.text\.(\S+) -+", ln)
        if m:
            cur = m.group(1); out[cur] = []
        elif ln.startswith("//-----"):
            cur = None
        elif cur:
            out[cur].append(ln)
    return out


def demangle(n):
    d = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    return re.sub(r"_INTERNAL_\w+::|\(anonymous namespace\)::|_GLOBAL__N__\w+::", "", d)


def subroutines(sec):
    """[(demangled name, [lines])] for the $-labelled subroutines of a kernel section"""
    idx = [(i, ln) for i, ln in enumerate(sec) if re.match(r"^\$[_A-Za-z0-9\$\.]+:", ln)]
    out = []
    for n, (i, ln) in enumerate(idx):
        name = ln.rstrip(":").split("$")[-1]
        j = idx[n + 1][0] if n + 1 < len(idx) else len(sec)
        out.append((demangle(name), sec[i:j]))
    return out


def loops(body):
    """every backward-branch loop: [(FP64 instructions, opcode histogram)]"""
    ins, labels = [], {}
    for ln in body:
        m = re.match(r"^(\.L_x_\d+):", ln)
        if m:
            labels[m.group(1)] = len(ins); continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?(\S+)\s*(.*?);", ln)
        if m:
            ins.append((m.group(1), m.group(2)))
    out = []
    for i, (op, args) in enumerate(ins):
        if not op.startswith("BRA"):
            continue
        m = re.search(r"\((\.L_x_\d+)\)", args)
        if m and m.group(1) in labels and labels[m.group(1)] <= i:
            ops = Counter(o.split(".")[0] if not o.startswith("MUFU") else o for o, _ in ins[labels[m.group(1)]:i + 1])
            out.append((sum(ops[k] for k in FP64), ops))
    return out


def main_loop(body):
    """the backward-branch loop with the most FP64 instructions: (opcode histogram, pairs per iteration)"""
    ls = loops(body)
    if not ls:
        return None
    ops = max(ls, key=lambda t: t[0])[1]
    return ops, max(1, ops.get("MUFU.RSQ", 1))


def main():
    res = {"how": "profiles/tools/sass_flops.py: FP64 instructions of the epoch loops per pair (flop = 2 DFMA + DMUL + DADD); per-kind "
                  "figures from an analysis build with the loops out of line, checked against the loops of the shipped kernels",
           "kernels": {}}
    want = {"thr": "k_kepler_likeILb1ELi1ELb0E", "lat": "k_kepler_likeILb1ELi1ELb1E"}
    kinds = {"astrom": "seg_astrom_ool<true, 1, 0,", "astrom_jitter": "seg_astrom_ool<true, 1, 1,", "rv": "seg_rv_ool<true, 1, false, false, false,",
             "rv_jitter": "seg_rv_ool<true, 1, false, true, false,", "rv_margin": "seg_rv_ool<true, 1, true, true, false,"}
    ana = analysis_build()
    for fam in ("octo_kernels_n1_lean", "octo_kernels_n1_full"):
        secs = sections(disassemble(ana, fam))
        for tag, key in want.items():
            sec = [v for k, v in secs.items() if key in k]
            if not sec:
                continue
            subs = subroutines(sec[0])
            for kind, pat in kinds.items():
                if kind in res["kernels"].get(tag, {}):
                    continue
                hit = [b for n, b in subs if n.startswith("void " + pat) or n.startswith(pat)]
                if not hit:
                    continue
                r = main_loop(hit[0])
                if not r:
                    continue
                ops, pairs = r
                n_all = sum(ops.values())
                res["kernels"].setdefault(tag, {})[kind] = {
                    "pairs_per_iteration": pairs, "instructions": round(n_all / pairs, 1),
                    "dfma": round(ops["DFMA"] / pairs, 1), "dmul": round(ops["DMUL"] / pairs, 1), "dadd": round(ops["DADD"] / pairs, 1),
                    "fp64_instructions": round(sum(ops[k] for k in FP64) / pairs, 1),
                    "flop": round((2 * ops["DFMA"] + ops["DMUL"] + ops["DADD"]) / pairs, 1)}
    # the shipped library: loops of the 1-planet lean gradient kernels (FL 1: no parameterisation stage)
    secs = sections(disassemble(SO, "octo_kernels_n1_lean"))
    res["shipped_loops"] = {}
    for tag, key in want.items():
        sec = [v for k, v in secs.items() if key in k and "Lb1ELi1E" in k[k.index(key) + len(key):]]
        if not sec:
            continue
        found = []
        for f, ops in loops(sec[0]):
            pairs = max(1, ops.get("MUFU.RSQ", 1))
            if f / pairs > 40:
                found.append({"pairs_per_iteration": pairs, "instructions": round(sum(ops.values()) / pairs, 1),
                              "fp64_instructions": round(f / pairs, 1), "flop": round((2 * ops["DFMA"] + ops["DMUL"] + ops["DADD"]) / pairs, 1)})
        res["shipped_loops"][tag] = {}
        for kind in ("astrom", "rv", "rv_jitter"):
            want_f = res["kernels"].get(tag, {}).get(kind, {}).get("flop")
            if not want_f:
                continue
            match = [x for x in found if abs(x["flop"] - want_f) <= 0.02 * want_f]
            if not match:
                raise SystemExit(f"{tag}/{kind}: {want_f} flop per pair in the analysis build, not found among the shipped loops {found}")
            res["shipped_loops"][tag][kind] = min(match, key=lambda x: x["instructions"])      # the inner loop (outer loops span more)
    json.dump(res, open(OUT, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
