"""Builds libapyib_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libapyib_b200.so")
SOURCES = ["contract.cu", "contract_tma.cu", "stream.cu", "dets.cu", "dets_tpm.cu", "dets_pairs.cu", "dets_pairs_k2_small.cu", "dets_pairs_k2_large.cu", "lemma.cu", "capi.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]
OBJDIR = os.path.join(LIBDIR, "obj")


STAMP = os.path.join(LIBDIR, "build.sha256")


def source_hash():
    """sha256 over every file of csrc/, the public header and the compiler flags (content, not mtimes: a snapshot
    copied to another box keeps its contents but not necessarily its timestamps)."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [os.path.join(HERE, "..", "include", "apyib_b200.h")]
    for f in files:
        if os.path.isfile(f):
            h.update(os.path.basename(f).encode())
            h.update(open(f, "rb").read())
    return h.hexdigest()


def _stale():
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    try:
        return open(STAMP).read().strip() != source_hash()
    except OSError:
        return True


def build_locked(force=False, verbose=False):
    """build() serialised across processes (torchrun ranks importing the package at the same time): one builds,
    the others wait on the lock and then find the library up to date."""
    import fcntl
    os.makedirs(LIBDIR, exist_ok=True)
    with open(os.path.join(LIBDIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return build(force=force, verbose=verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    flags = NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else [])

    def compile_one(src):                      # one nvcc per translation unit, all of them concurrently
        obj = os.path.join(OBJDIR, os.path.splitext(src)[0] + ".o")
        res = subprocess.run([nvcc] + flags + ["-c", os.path.join(CSRC, src), "-o", obj], capture_output=True, text=True)
        return src, obj, res

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    for src, obj, res in results:
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("nvcc failed compiling %s" % src)
        if verbose:
            print(res.stdout + res.stderr)
    tmp = LIB + ".tmp.%d" % os.getpid()
    res = subprocess.run([nvcc] + NVCC_FLAGS + ["-shared"] + [obj for _, obj, _ in results] + ["-o", tmp],
                         capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libapyib_b200.so")
    os.replace(tmp, LIB)
    with open(STAMP + ".tmp.%d" % os.getpid(), "w") as f:
        f.write(source_hash() + "\n")
    os.replace(STAMP + ".tmp.%d" % os.getpid(), STAMP)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
