"""Build the CUDA extension in-tree: embiggen_b200/libb2e.so (sm_100a only).

Every .cu is compiled to an object under csrc/_obj/ (in parallel, only when stale) and the
objects are linked into the shared library, so touching one kernel file costs one compile."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB_PATH = os.path.join(_HERE, "libb2e.so")
SOURCES = ["b2e_api.cu", "walk_kernels.cu", "sgns_kernels.cu", "sgns_pipe.cu", "graph_build.cu", "glove.cu",
           "edge_pred.cu", "exchange.cu", "alias_build.cu"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++",
]


def _headers():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(_HERE, "..", "include", "b2e.h"))
    return deps


def _stale(target, deps) -> bool:
    if not os.path.exists(target):
        return True
    built = os.path.getmtime(target)
    return any(os.path.getmtime(d) > built for d in deps)


def needs_build() -> bool:
    sources = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    return _stale(LIB_PATH, sources + _headers())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    headers = _headers()
    sources = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]

    def compile_one(name):
        source = os.path.join(CSRC, name)
        target = os.path.join(OBJ, name[:-3] + ".o")
        if force or _stale(target, [source] + headers):
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", source, "-o", target]
            subprocess.run(cmd, check=True)
        return target

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        objects = list(pool.map(compile_one, sources))
    subprocess.run([nvcc, "-shared", "-cudart", "shared", "-gencode", "arch=compute_100a,code=sm_100a",
                    "-ccbin", "/usr/bin/g++", "-o", LIB_PATH] + objects,
                   check=True)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
