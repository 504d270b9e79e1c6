"""Executed FP64 operations per (epoch x chain) pair, counted from the SASS of libocto_b200.so ("the
planning weights are to be replaced by an op counter").

For the 1-planet gradient kernels — throughput instantiation k_kepler_like<true,1,false> and latency instantiation
k_kepler_like<true,1,true> — find the epoch loop of every segment subroutine (the backward branch whose body holds the
most FP64 instructions), divide by the pairs in flight per iteration (one MUFU.RSQ per Kepler solve) and count
DFMA (2 flop), DMUL, DADD (1 flop each).  Writes profiles/sass_flops.json, which bench.py reads for `roofline.frac`.

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


def disassemble(so):
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "octo_kernels", so], cwd=d, check=True, capture_output=True)
        cub = max((f for f in os.listdir(d) if f.endswith(".cubin")), key=lambda f: os.path.getsize(os.path.join(d, f)))
        return subprocess.run(["nvdisasm", os.path.join(d, cub)], capture_output=True, text=True, check=True).stdout.split("\n")


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


def main_loop(body):
    """the backward-branch loop with the most FP64 instructions: (opcode histogram, pairs per iteration)"""
    ins, labels = [], {}
    for ln in body:
        m = re.match(r"^(\.L_x_\d+):", ln)
        if m:
            labels[m.group(1)] = len(ins); continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?(\S+)\s*(.*?);", ln)
        if m:
            ins.append((m.group(1), m.group(2)))
    best = None
    for i, (op, args) in enumerate(ins):
        if not op.startswith("BRA"):
            continue
        m = re.search(r"\((\.L_x_\d+)\)", args)
        if m and m.group(1) in labels and labels[m.group(1)] <= i:
            ops = Counter(o.split(".")[0] if not o.startswith("MUFU") else o for o, _ in ins[labels[m.group(1)]:i + 1])
            f = sum(ops[k] for k in FP64)
            if best is None or f > best[0]:
                best = (f, ops)
    if best is None:
        return None
    ops = best[1]
    return ops, max(1, ops.get("MUFU.RSQ", 1))


def main():
    secs = sections(disassemble(SO))
    res = {"how": "profiles/tools/sass_flops.py: FP64 instructions of the epoch loops in the SASS of libocto_b200.so, per pair "
                  "(flop = 2 DFMA + DMUL + DADD)", "kernels": {}}
    want = {"thr": "k_kepler_likeILb1ELi1ELb0E", "lat": "k_kepler_likeILb1ELi1ELb1E"}
    kinds = {"astrom": "seg_astrom<true, 1, 0,", "astrom_jitter": "seg_astrom<true, 1, 1,", "rv": "seg_rv<true, 1, false, false, false,",
             "rv_jitter": "seg_rv<true, 1, false, true, false,", "rv_margin": "seg_rv<true, 1, true, true, false,"}
    for tag, key in want.items():
        sec = [v for k, v in secs.items() if key in k]
        if not sec:
            continue
        subs = subroutines(sec[0])
        for kind, pat in kinds.items():
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
    json.dump(res, open(OUT, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
