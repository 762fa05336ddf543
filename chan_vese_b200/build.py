"""In-tree build of the sm_100a shared library (nvcc cross-compiles without a GPU).

    python -m chan_vese_b200.build [--force]

The library lands in chan_vese_b200/lib/libchan_vese_b200.so (git-ignored, travels with gpurun snapshots).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libchan_vese_b200.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "chan_vese_b200.h")
SOURCES = ["api.cu", "csv_kernels.cu", "pm_kernels.cu", "f32_kernels.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "static"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [HEADER, os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra_flags=()):
    """Compile the CUDA extension for sm_100a if it is missing or older than its sources."""
    if not force and not _stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    extra_flags = list(extra_flags) + os.environ.get("CVB_EXTRA_NVCC_FLAGS", "").split()  # tuning sweeps only
    cmd = [_nvcc()] + NVCC_FLAGS + list(extra_flags) + ["-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    # the image exports CC/CXX pointing at a relocated gcc; nvcc's default host compiler (g++ on PATH) is fine
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return LIB_PATH


CLI_SRC = os.path.join(os.path.dirname(HERE), "cli", "chan_vese.cpp")
CLI_BIN = os.path.join(os.path.dirname(HERE), "bin", "chan_vese")


def build_cli(force=False):
    """g++ -std=c++14 host front-end (the reference's bin/chan_vese surface) linked against the C-ABI library."""
    build()
    if (not force and os.path.exists(CLI_BIN) and os.path.getmtime(CLI_BIN) >= max(os.path.getmtime(CLI_SRC), os.path.getmtime(HEADER),
                                                                                    os.path.getmtime(LIB_PATH))):
        return CLI_BIN
    os.makedirs(os.path.dirname(CLI_BIN), exist_ok=True)
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [gxx, "-std=c++14", "-O2", "-Wall", "-Wextra", "-I", os.path.dirname(HEADER), "-o", CLI_BIN, CLI_SRC, "-L", LIB_DIR,
           "-lchan_vese_b200", "-Wl,-rpath," + LIB_DIR]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + res.stdout + res.stderr)
    return CLI_BIN


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
    print(build_cli(force="--force" in sys.argv))
