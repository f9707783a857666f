"""Builds libdissc_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdissc_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale() -> bool:
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every csrc/*.cu (one nvcc process per file, in parallel) and link the shared library."""
    if not force and not _stale():
        return LIB
    return _build(LIB, os.path.join(HERE, "build"), [], force, verbose)


def build_variant(tag: str, defines, verbose: bool = False) -> str:
    """A/B builds of compile-time constants: `libdissc_b200_<tag>.so` compiled with extra -D flags, selected at run time with
    DISSC_LIB=<path> (dissc_b200/_lib.py).  Measurement aid only."""
    out = os.path.join(HERE, f"libdissc_b200_{tag}.so")
    return _build(out, os.path.join(HERE, f"build_{tag}"), [f"-D{d}" for d in defines], True, verbose)


def _build(LIB: str, objdir: str, extra, force: bool, verbose: bool) -> str:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(objdir, exist_ok=True)
    objs, procs = [], []
    hdr_t = max([os.path.getmtime(h) for h in glob.glob(os.path.join(CSRC, "*.cuh"))
                 + glob.glob(os.path.join(HERE, "..", "include", "*.h"))] + [0.0])
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if not force and os.path.isfile(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_t):
            continue
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd)))
    for src, p in procs:
        if p.wait() != 0:
            raise RuntimeError(f"nvcc failed on {src}")
    subprocess.check_call([nvcc, "-shared", "-o", LIB, *objs, "-lcudart"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
