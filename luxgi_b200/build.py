"""In-tree build of libluxddgi.so (sm_100a only).  Usable as `python -m luxgi_b200.build`.

nvcc cross-compiles without a GPU; the resulting .so sits next to this file (git-ignored, shipped to the GPU box by gpurun).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libluxddgi.so")
SOURCES = ["ddgi_kernels.cu", "ddgi_engine.cpp"]
HEADERS = ["ddgi_kernels.h", "ddgi_math.cuh", "march_kernel.inc", "blend_tc.inc", "blend_lists.inc", "blend_umma.inc", os.path.join("..", "..", "include", "luxddgi.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",            # numerics contract: FMA only where __fmaf_rn is written
    "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-O2,-Wall,-fvisibility=hidden,-ffp-contract=off",  # host float code (SDF build chunk lists) must not contract either
    "-Xptxas", "-v",
]


def nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    env = dict(os.environ)
    env.pop("CXX", None)  # the image exports a g++ wrapper nvcc does not need
    env.pop("CC", None)
    flags = list(NVCC_FLAGS)
    cmd = [nvcc(), "-shared", "-o", LIB] + flags + ["-x", "cu"] + [os.path.join(CSRC, f) for f in SOURCES]
    cmd += ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []
    res = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = res.stdout
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libluxddgi.so")
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
