"""Builds libubu_sw.so (CUDA, sm_100a only) in-tree with nvcc. No torch involved.

Two translation units (the Smith-Waterman path and the k-mer graph build) are compiled side by side and linked
into the one shared library the C ABI (include/ubu_sw.h, include/ubu_kmer.h) lives in."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
UNITS = {
    "ubu_sw.cu": ["sw_common.cuh", "sw_fwd16.cuh", "sw_fwd16s.cuh", "sw_tb.cuh", "sw_tbstrip.cuh", "sw_long.cuh", "sw_post.cuh"],
    "ubu_kmer.cu": ["sw_common.cuh"],
    "ubu_pack.cu": [],
}
HEADERS = [os.path.join(ROOT, "include", f) for f in ("ubu_sw.h", "ubu_kmer.h")]
OBJ_DIR = os.path.join(ROOT, "build", "obj")
LIB = os.path.join(HERE, "libubu_sw.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]


def nvcc_path() -> str:
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _deps(unit: str):
    return [os.path.join(CSRC, unit)] + [os.path.join(CSRC, d) for d in UNITS[unit]] + HEADERS


def _obj(unit: str) -> str:
    return os.path.join(OBJ_DIR, unit.replace(".cu", ".o"))


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def needs_build() -> bool:
    return any(_stale(LIB, _deps(u)) for u in UNITS)


def _compile(unit: str, extra, verbose: bool):
    cmd = [nvcc_path(), *NVCC_FLAGS, *extra, "-c", "-o", _obj(unit), os.path.join(CSRC, unit)]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    return unit, subprocess.run(cmd, capture_output=True, text=True)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJ_DIR, exist_ok=True)
    extra = os.environ.get("UBU_NVCC_EXTRA", "").split()
    todo = [u for u in UNITS if force or extra or _stale(_obj(u), _deps(u))]
    with ThreadPoolExecutor(max_workers=len(UNITS)) as ex:
        for unit, r in ex.map(lambda u: _compile(u, extra, verbose), todo):
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError(f"nvcc failed compiling {unit}")
            if verbose:
                sys.stderr.write(r.stderr)
    r = subprocess.run([nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB, *[_obj(u) for u in UNITS]],
                       capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking libubu_sw.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
