"""Builds libbm_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libbm_b200.so")
STAMP = LIB + ".sha256"        # digest of the sources the library was built from (git-ignored, travels with the .so)
SOURCES = ["bm_api.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "--expt-relaxed-constexpr"]


def _source_digest() -> str:
    """sha256 over the flags, every file under csrc/ and the public header: what the built library depends on.  Content, not
    mtimes -- a checkout, a snapshot copy to another box or a `touch` must not trigger (or hide the need for) a rebuild."""
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    paths = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC))
    paths.append(os.path.join(os.path.dirname(HERE), "include", "bm_b200.h"))
    for p in paths:
        if os.path.isfile(p):
            h.update(os.path.basename(p).encode())
            with open(p, "rb") as f:
                h.update(f.read())
    return h.hexdigest()


def is_current() -> bool:
    if not (os.path.isfile(LIB) and os.path.isfile(STAMP)):
        return False
    with open(STAMP) as f:
        return f.read().strip() == _source_digest()


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and is_current():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
        [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB, "-lcudart"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libbm_b200.so")
    if verbose:
        sys.stderr.write(res.stderr)
    with open(STAMP, "w") as f:
        f.write(_source_digest() + "\n")
    return LIB


if __name__ == "__main__":
    print(build_library(force=True, verbose="-v" in sys.argv))
