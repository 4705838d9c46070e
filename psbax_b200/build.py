"""Build ``libpsbax_b200.so`` in-tree with nvcc for sm_100a.

    python -m psbax_b200.build [--force] [--verbose]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libpsbax_b200.so")
SOURCES = ["psbax_b200.cu", "simt.cu", "tensor_bf16.cu", "tensor_tf32.cu", "tensor_proj.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]
LINK_FLAGS = ["-shared", "-lcuda"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "psbax_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, trace: bool = False, variant: str = "",
          defines=()) -> str:
    """trace=True builds libpsbax_b200_trace.so with the clock64 pipeline trace compiled in
    (profiling aid; select it with PSBAX_LIB=<path>).  variant="name" + defines=("-DX=1", ...) builds
    libpsbax_b200_<name>.so for same-box A/B timing of an experiment."""
    out = OUT.replace(".so", "_trace.so") if trace else OUT
    if variant:
        out = OUT.replace(".so", f"_{variant}.so")
    if not force and not trace and not variant and not _stale():
        return OUT
    # one object per translation unit, compiled in parallel, then one link
    from concurrent.futures import ThreadPoolExecutor

    objdir = os.path.join(HERE, "build", variant or ("trace" if trace else "obj"))
    os.makedirs(objdir, exist_ok=True)
    flags = (NVCC_FLAGS + (["-DPSBAX_TC_TRACE=1"] if trace else []) + list(defines) +
             (["-Xptxas", "-v"] if verbose else []))

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        res = subprocess.run([_nvcc()] + flags + ["-c", "-o", obj, src], cwd=CSRC, capture_output=True, text=True)
        return src, obj, res

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    bad = False
    for src, _obj, res in results:
        if verbose or res.returncode != 0:
            sys.stderr.write(f"==== {src}\n" + res.stdout + res.stderr)
        bad = bad or res.returncode != 0
    if bad:
        raise RuntimeError("nvcc failed building libpsbax_b200.so")
    res = subprocess.run([_nvcc()] + NVCC_FLAGS + LINK_FLAGS + ["-o", out] + [o for _s, o, _r in results],
                         cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libpsbax_b200.so")
    return out


if __name__ == "__main__":
    _variant = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else ""
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, trace="--trace" in sys.argv,
                variant=_variant, defines=[a for a in sys.argv[1:] if a.startswith("-D")]))
