"""Builds libb200nn.so IN-TREE with nvcc for sm_100a (cross-compiles without a GPU).

    python -m neuralnetlib_b200._build [--force] [--verbose]

Objects go to neuralnetlib_b200/_build/, the shared library to neuralnetlib_b200/_lib/libb200nn.so
(git-ignored, but shipped to the GPU box by gpurun). A source is recompiled only when its content hash
(or a header's) changed.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
BUILD = os.path.join(PKG, "_build")
LIBDIR = os.path.join(PKG, "_lib")
LIB = os.path.join(LIBDIR, "libb200nn.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libb200nn.so cannot be built (there is no CPU fallback)")


def _hash(paths) -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    return h.hexdigest()


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(ROOT, "include", "b200nn.h"))
    return sorted(hs)


def _deps(src):
    """The source plus the local headers it includes, transitively (`#include "x"` relative to csrc/)."""
    import re
    seen, todo = set(), [src]
    while todo:
        f = todo.pop()
        if f in seen or not os.path.exists(f):
            continue
        seen.add(f)
        with open(f) as fh:
            for inc in re.findall(r'^\s*#include\s+"([^"]+)"', fh.read(), flags=re.M):
                todo.append(os.path.normpath(os.path.join(os.path.dirname(f), inc)))
    return sorted(seen)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    nvcc = _nvcc()
    jobs, objs = [], []
    for src in sources():
        obj = os.path.join(BUILD, os.path.basename(src)[:-3] + ".o")
        stamp = obj + ".sha"
        digest = _hash(_deps(src))
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == digest:
            continue
        jobs.append((src, obj, stamp, digest))

    def compile_one(job):
        src, obj, stamp, digest = job
        cmd = [nvcc] + NVCC_FLAGS + ["-DB200NN_BUILDING", "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = r.stdout + r.stderr
        with open(obj + ".log", "w") as f:
            f.write(" ".join(cmd) + "\n" + log)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{log}")
        with open(stamp, "w") as f:
            f.write(digest)
        return src, log

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for src, log in ex.map(compile_one, jobs):
                if verbose:
                    print(f"[b200nn build] {os.path.basename(src)}\n{log}")
    if jobs or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcuda"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    lib = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(lib)
