"""Build libdcb.so (CUDA kernels + C ABI) in-tree for sm_100a.

    python -m decombinator_b200.build [--force]

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels with gpurun.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdcb.so")
SYNTH_LIB = os.path.join(HERE, "libdcbsynth.so")   # the synthetic read generator alone (bench.py's reference arm maps only this + oracle/)
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

SOURCES = ["decombine.cu", "collapse.cu", "tagset.cpp", "pack.cpp", "fastq.cpp", "group.cpp", "synth.cpp", "error.cpp"]

FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function", "-shared", "-Xptxas", "-v",
]


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(os.path.dirname(HERE), "include", "dcb.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    cmd = [NVCC] + FLAGS + ["-I", os.path.join(os.path.dirname(HERE), "include"), "-o", LIB] + sources() + ["-lpthread", "-lz"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libdcb.so")
    with open(os.path.join(HERE, "build.log"), "w") as fh:
        fh.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
    build_synth(force=True)
    return LIB


def build_synth(force=False):
    """libdcbsynth.so: csrc/synth.cpp alone (plain C++, no CUDA)."""
    src = [os.path.join(CSRC, "synth.cpp"), os.path.join(CSRC, "error.cpp")]
    if not force and os.path.exists(SYNTH_LIB) and all(os.path.getmtime(x) <= os.path.getmtime(SYNTH_LIB) for x in src):
        return SYNTH_LIB
    cmd = ["g++", "-O3", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(os.path.dirname(HERE), "include"), "-o", SYNTH_LIB] + src + ["-lpthread"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ failed building libdcbsynth.so")
    return SYNTH_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
