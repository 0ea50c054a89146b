"""Builds libmct_sm100a.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).  The translation units are
compiled in parallel (the persistent solver is instantiated per wavenumbers-per-thread in separate files)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
PROFILE = os.environ.get("MCT_PROFILE") == "1"      # instrumented build (phase counters), kept beside the product library
OBJDIR = os.path.join(LIBDIR, "obj_prof" if PROFILE else "obj")
LIB = os.path.join(LIBDIR, "libmct_sm100a_prof.so" if PROFILE else "libmct_sm100a.so")
SOURCES = ["abi.cu", "layout.cu", "eval.cu", "td_sc.cu", "td_sc_k1.cu", "td_sc_k2.cu", "td_sc_k4.cu", "td_sc_k8.cu", "td_grid.cu", "msd.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",            # reference-order arithmetic: FMAs only where the code asks for them with fma()
    "-Xcompiler", "-fPIC,-O3,-fvisibility=hidden,-Wall",
]


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "mct.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = ["-DMCT_PROFILE"] if PROFILE else []
    extra += ["-Xptxas", "-v"] if verbose else []
    extra += os.environ.get("MCT_NVCC_EXTRA", "").split()

    def compile_one(src):
        obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        r = subprocess.run([nvcc] + NVCC_FLAGS + extra + ["-c", "-o", obj, os.path.join(CSRC, src)], capture_output=True, text=True)
        return src, obj, r

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    for src, obj, r in results:
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise subprocess.CalledProcessError(r.returncode, f"nvcc {src}")
    # --cudart shared: one CUDA runtime per process, shared with torch when both are loaded
    subprocess.check_call([nvcc, "-shared", "--cudart", "shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + [obj for _, obj, _ in results])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
