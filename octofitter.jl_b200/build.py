"""Build libocto_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "lib", "libocto_b200.so")
SOURCES = ["octo_kernels.cu", "octo_param.cu", "octo_hmc.cu", "octo_shim.cu"]
DEPS = SOURCES + ["octo_internal.h", "octo_param_dev.cuh", "octo_hmc_dev.cuh", os.path.join("..", "..", "include", "octo_b200.h")]
NVCC_CFLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
               "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
NVCC_LFLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC"]


def _compile_and_link(out, defines, log_path=None):
    """octo_kernels.cu is compiled once per planet-count instantiation and kernel family (-DOCTO_NPT=1|2|3|4
    -DOCTO_LEANSEL=0|1) — those eight objects and the other sources in parallel — then everything is linked into one shared library."""
    import tempfile
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    only1 = "-DOCTO_NPT_ONLY1" in defines
    # (source, extra flags, object); the kernel objects are built from one-line wrapper sources so that each embedded
    # cubin has its own name (cuobjdump -xelf, profiles/tools/sass_flops.py); line info still points at csrc/octo_kernels.cu
    units = [("octo_kernels_n%d_%s.cu" % (k, "lean" if l else "full"), ["-DOCTO_NPT=%d" % k, "-DOCTO_LEANSEL=%d" % l, "-I", CSRC], None)
             for k in ((1,) if only1 else (1, 2, 3, 4)) for l in (0, 1)]
    units += [(os.path.join(CSRC, s), [], None) for s in SOURCES if s != "octo_kernels.cu"]
    with tempfile.TemporaryDirectory() as tmp:
        for i, (src, extra, _) in enumerate(units):
            if not os.path.isabs(src):
                with open(os.path.join(tmp, src), "w") as f:
                    f.write('#include "%s"\n' % os.path.join(CSRC, "octo_kernels.cu"))
                src = os.path.join(tmp, src)
            units[i] = (src, extra, os.path.join(tmp, os.path.basename(src).replace(".cu", ".o")))

        def compile_one(u):
            src, extra, obj = u
            cmd = [nvcc] + NVCC_CFLAGS + list(defines) + extra + ["-c", src, "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            return cmd, r
        with ThreadPoolExecutor(max_workers=len(units)) as ex:
            results = list(ex.map(compile_one, units))
        log = ""
        for cmd, r in results:
            log += " ".join(cmd) + "\n" + r.stdout + r.stderr
        ok = all(r.returncode == 0 for _, r in results)
        if ok:
            cmd = [nvcc] + NVCC_LFLAGS + [u[2] for u in units] + ["-o", out, "-ldl"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            log += " ".join(cmd) + "\n" + r.stdout + r.stderr
            ok = r.returncode == 0
    if log_path:
        with open(log_path, "w") as f:
            f.write(log)
    if not ok:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building " + out)
    return log


def build_variant(out, defines):
    """Tuning builds (e.g. -DOCTO_LAT_ILP=3; add -DOCTO_NPT_ONLY1 for one-planet kernels only, a third of the compile
    time) into another path; select with OCTO_B200_LIB (profiles/tools/ab.sh)."""
    return _compile_and_link(out, defines)


def build(force=False, verbose=False):
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    newest = max(os.path.getmtime(os.path.join(CSRC, d)) for d in DEPS)
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= newest:
        return OUT
    log = _compile_and_link(OUT, [], os.path.join(HERE, "lib", "build.log"))
    if verbose:
        print(log)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
