"""Build libagmv_b200.so (CUDA kernels + C-ABI) and libagmv_dropin.so (the
reference-facing C API) in-tree for sm_100a. nvcc cross-compiles without a GPU."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-Xcompiler", "-Wno-unused-function"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force=False, verbose=False):
    lib = os.path.join(HERE, "libagmv_b200.so")
    cu = [os.path.join(CSRC, "api.cu")]
    deps = cu + [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "agmv_b200.h")]
    if force or _newer(lib, deps):
        cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-shared", "-o", lib] + cu + ["-lcudart", "-lpthread"]
        subprocess.run(cmd, check=True)
    dropin_src = os.path.join(CSRC, "agmv_dropin.c")
    if os.path.exists(dropin_src):
        dlib = os.path.join(HERE, "libagmv_dropin.so")
        if force or _newer(dlib, [dropin_src, os.path.join(ROOT, "include", "agmv_dropin.h"), lib]):
            subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-Wall", "-I", os.path.join(ROOT, "include"), "-o", dlib, dropin_src,
                            "-L", HERE, "-lagmv_b200", "-Wl,-rpath,$ORIGIN", "-lm"], check=True)
    return lib


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built")
