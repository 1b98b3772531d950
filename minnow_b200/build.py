"""Builds libminnow_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libminnow_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # bit-exactness with Go on amd64: no FMA contraction, IEEE div/sqrt, no fast-math
    "--fmad=false", "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-ffp-contract=off,-fno-fast-math,-O2",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp")))


def headers():
    inc = os.path.join(os.path.dirname(HERE), "include")
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".hpp"))]
    hs += [os.path.join(inc, f) for f in os.listdir(inc)]
    return hs


def build_library(force=False, verbose=False):
    """nvcc -> libminnow_b200.so; the translation units are compiled in parallel (one nvcc each) and linked."""
    deps = sources() + headers() + [os.path.abspath(__file__)]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = ["-D" + d for d in os.environ.get("MNW_DEFINES", "").split() if d]   # e.g. MNW_DEFINES=MNW_PIPE_DBG
    flags = NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else [])
    objdir = os.path.join(CSRC, "_build")
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        subprocess.check_call([nvcc] + flags + ["-c", "-o", obj, src])
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(sources()))) as pool:
        objs = list(pool.map(compile_one, sources()))
    subprocess.check_call([nvcc] + flags + ["-shared", "-o", LIB] + objs)
    return LIB


if __name__ == "__main__":
    import sys
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
