"""Builds cuvs_b200/lib/libcuvs_c.so — the C-ABI drop-in — with nvcc for sm_100a, in-tree.

    python -m cuvs_b200.build [--force] [--verbose]

Objects are cached under cuvs_b200/lib/obj and rebuilt when a source or any header changes.
nvcc cross-compiles without a GPU; the resulting .so travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import concurrent.futures as cf
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
OBJ_DIR = os.path.join(LIB_DIR, "obj")
LIB = os.path.join(LIB_DIR, "libcuvs_c.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall,-Wno-unused-function",
          "--expt-relaxed-constexpr", "-ccbin", "/usr/bin/g++",
          "-I", os.path.join(ROOT, "include"), "-I", SRC]


def _newest_header_mtime() -> float:
    hs = glob.glob(os.path.join(SRC, "*.hpp")) + glob.glob(os.path.join(SRC, "*.cuh")) + \
        glob.glob(os.path.join(ROOT, "include", "**", "*.h"), recursive=True) + [os.path.abspath(__file__)]
    return max(os.path.getmtime(h) for h in hs)


def _compile(src: str, obj: str, verbose: bool) -> str:
    cmd = [NVCC] + ARCH + CFLAGS + ["-c", src, "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {os.path.basename(src)}:\n{r.stdout}\n{r.stderr}")
    return (r.stdout + r.stderr) if verbose else ""


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    sources = sorted(glob.glob(os.path.join(SRC, "*.cu")))
    hdr_m = _newest_header_mtime()
    jobs, objs = [], []
    for s in sources:
        o = os.path.join(OBJ_DIR, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_m):
            jobs.append((s, o))
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            futs = {ex.submit(_compile, s, o, verbose): s for s, o in jobs}
            for f in cf.as_completed(futs):
                out = f.result()
                if verbose and out:
                    print(f"--- {os.path.basename(futs[f])}\n{out}")
    if jobs or not os.path.exists(LIB):
        cmd = [NVCC] + ARCH + ["-shared", "-ccbin", "/usr/bin/g++", "-o", LIB] + objs + ["-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
