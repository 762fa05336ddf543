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


STAMP = os.path.join(LIB_DIR, "build_stamp")


def _digest(extra_flags=()):
    """Hash of everything the library is made of.  Staleness is decided by CONTENT, not by modification times: a
    snapshot of the tree (gpurun, a checkout) copies files in arbitrary order, and several ranks import this module at the
    same moment -- none of them may decide to rebuild a library that is already right."""
    import hashlib
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)) + [HEADER, os.path.abspath(__file__)]:
        path = f if os.path.isabs(f) else os.path.join(CSRC, f)
        if os.path.isfile(path):
            h.update(os.path.basename(path).encode())
            with open(path, "rb") as fh:
                h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS + list(extra_flags)).encode())
    return h.hexdigest()


def _stale(lib_path=LIB_PATH, stamp=STAMP, extra_flags=()):
    if not os.path.exists(lib_path) or not os.path.exists(stamp):
        return True
    with open(stamp) as fh:
        return fh.read().strip() != _digest(extra_flags)


def build(force=False, verbose=False, extra_flags=(), out=None):
    """Compile the CUDA extension for sm_100a if it is missing or was built from other sources: one object per .cu file
    (compiled in parallel), linked into one shared library that is moved into place atomically, under a file lock (ranks
    of one job may call this concurrently).  `out`: alternative library path (tuning variants built with extra flags)."""
    import fcntl
    extra_flags = list(extra_flags) + os.environ.get("CVB_EXTRA_NVCC_FLAGS", "").split()  # tuning sweeps only
    lib_path = out or LIB_PATH
    stamp = STAMP if out is None else lib_path + ".stamp"
    if not force and not _stale(lib_path, stamp, extra_flags):
        return lib_path
    os.makedirs(os.path.dirname(lib_path), exist_ok=True)
    with open(lib_path + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not _stale(lib_path, stamp, extra_flags):  # another process built it while we waited
            return lib_path
        obj_dir = OBJ_DIR if out is None and not extra_flags else lib_path + ".obj"
        os.makedirs(obj_dir, exist_ok=True)
        jobs = [[_nvcc()] + NVCC_FLAGS + extra_flags + ["-c", os.path.join(CSRC, src), "-o", os.path.join(obj_dir, src[:-3] + ".o")]
                for src in SOURCES]
        # the image exports CC/CXX pointing at a relocated gcc; nvcc's default host compiler (g++ on PATH) is fine
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(len(jobs)) as ex:
            for cmd, res in zip(jobs, ex.map(lambda c: subprocess.run(c, capture_output=True, text=True), jobs)):
                if verbose:
                    print(" ".join(cmd), file=sys.stderr)
                if res.returncode != 0:
                    raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
        tmp = lib_path + ".tmp.%d" % os.getpid()
        link = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-o", tmp] + \
               [os.path.join(obj_dir, src[:-3] + ".o") for src in SOURCES] + ["-ldl"]
        if verbose:
            print(" ".join(link), file=sys.stderr)
        res = subprocess.run(link, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
        os.replace(tmp, lib_path)
        with open(stamp + ".tmp", "w") as fh:
            fh.write(_digest(extra_flags))
        os.replace(stamp + ".tmp", stamp)
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
