"""In-tree build of libdvae_b200.so (sm_100a only) with plain nvcc; no torch headers involved.

`python -m dvae_b200.build` or `__graft_entry__.build()`.  Objects are cached under csrc/build/ and
rebuilt when a source or header is newer.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(PKG_DIR), "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libdvae_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_mtime() -> float:
    return max([os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC)
                if f.endswith((".cuh", ".h"))] + [0.0])


def build(verbose: bool = False, force: bool = False) -> str:
    os.makedirs(os.path.join(CSRC, "build"), exist_ok=True)
    hdr_m = _headers_mtime()
    jobs = []
    objs = []
    for src in _sources():
        s = os.path.join(CSRC, src)
        o = os.path.join(CSRC, "build", src[:-3] + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_m):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        cmd = [NVCC, *ARCH_FLAGS, *COMMON, "-c", s, "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {s}:\n{r.stdout}\n{r.stderr}")
        with open(o[:-2] + ".ptxas.log", "w") as f:
            f.write(r.stderr)
        if verbose:
            print(f"[dvae_b200.build] compiled {os.path.basename(s)}")

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(compile_one, jobs))
    if jobs or not os.path.exists(LIB_PATH):
        cmd = [NVCC, *ARCH_FLAGS, "-shared", "-o", LIB_PATH, *objs]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(f"[dvae_b200.build] linked {LIB_PATH}")
    return LIB_PATH


if __name__ == "__main__":
    build(verbose=True, force="--force" in sys.argv)
