"""Find loops (backward branches) in a cuobjdump -sass function and print their opcode histograms.
usage: sass_loops.py all.sass <function-substring> [min_len]"""
import re
import sys
from collections import Counter

txt = open(sys.argv[1]).read().split("Function : ")
fn = [t for t in txt if sys.argv[2] in t.split("\n")[0]][0]
minlen = int(sys.argv[3]) if len(sys.argv) > 3 else 40
ins = []
for ln in fn.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", ln)
    if m:
        ins.append((int(m.group(1), 16), m.group(2)))
print("instructions:", len(ins), "bytes:", ins[-1][0] + 16)
addr2i = {a: i for i, (a, _) in enumerate(ins)}
loops = []
for i, (a, t) in enumerate(ins):
    m = re.search(r"\bBRA\b.*?0x([0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr2i and i - addr2i[tgt] >= minlen:
            loops.append((addr2i[tgt], i))
for lo, hi in loops:
    ops = Counter()
    for a, t in ins[lo:hi + 1]:
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        ops[t.split()[0].split(".")[0]] += 1
    n = hi - lo + 1
    f64 = sum(v for k, v in ops.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
    print(f"loop {ins[lo][0]:#x}..{ins[hi][0]:#x}: {n} instr, fp64={f64}")
    print("   ", ", ".join(f"{k}:{v}" for k, v in ops.most_common(30)))
