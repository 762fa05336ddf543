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
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]
OBJ_DIR = os.path.join(HERE, "lib", "obj")


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _headers():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + [HEADER, os.path.abspath(__file__)]


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [HEADER, os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra_flags=(), out=None):
    """Compile the CUDA extension for sm_100a if it is missing or older than its sources: one object per .cu file
    (compiled in parallel, only the stale ones), linked into one shared library.  `out`: alternative library path (tuning
    variants built with extra flags; their objects are not cached)."""
    extra_flags = list(extra_flags) + os.environ.get("CVB_EXTRA_NVCC_FLAGS", "").split()  # tuning sweeps only
    variant = bool(extra_flags) or out is not None
    lib_path = out or LIB_PATH
    if not force and not variant and not _stale():
        return lib_path
    obj_dir = OBJ_DIR if not variant else lib_path + ".obj"
    os.makedirs(obj_dir, exist_ok=True)
    os.makedirs(os.path.dirname(lib_path), exist_ok=True)
    hdr_time = max(os.path.getmtime(h) for h in _headers())
    jobs = []
    for src in SOURCES:
        sp, op = os.path.join(CSRC, src), os.path.join(obj_dir, src[:-3] + ".o")
        if force or variant or not os.path.exists(op) or os.path.getmtime(op) < max(os.path.getmtime(sp), hdr_time):
            jobs.append([_nvcc()] + NVCC_FLAGS + extra_flags + ["-c", sp, "-o", op])
    # the image exports CC/CXX pointing at a relocated gcc; nvcc's default host compiler (g++ on PATH) is fine
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max(1, len(jobs))) as ex:
        for cmd, res in zip(jobs, ex.map(lambda c: subprocess.run(c, capture_output=True, text=True), jobs)):
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
            if res.returncode != 0:
                raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    link = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-o", lib_path] + \
           [os.path.join(obj_dir, src[:-3] + ".o") for src in SOURCES] + ["-ldl"]
    if verbose:
        print(" ".join(link), file=sys.stderr)
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    return lib_path


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
