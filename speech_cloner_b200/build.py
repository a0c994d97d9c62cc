"""Build ``libspeechdsp.so`` in-tree with nvcc for sm_100a (B200).

``python -m speech_cloner_b200.build`` or ``__graft_entry__.build()``.  The .so is git-ignored but
travels with the gpurun snapshot; nothing is JIT-compiled at run time.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "csrc", "speechdsp.cu")
OUT = os.path.join(_HERE, "libspeechdsp.so")
DEPS = [os.path.join(_HERE, "csrc", f) for f in
        ("speechdsp.cu", "common.cuh", "dft20.cuh", "fft400.cuh", "fe_kernels.cuh", "gl_kernels.cuh",
         "generic_kernels.cuh", "fe_ws.cuh", "phn_kernels.cuh", "sampler_kernels.cuh")] + [os.path.join(_HERE, "..", "include", "speechdsp.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def source_hash() -> str:
    """sha256 over the kernel sources and the C header: pins committed ncu numbers (profiles/*_traffic.json) to the code
    they were captured from."""
    import hashlib
    h = hashlib.sha256()
    base = os.path.join(_HERE, "csrc")
    for name in sorted(os.listdir(base)):
        if name.endswith((".cu", ".cuh")):
            h.update(name.encode())
            h.update(open(os.path.join(base, name), "rb").read())
    h.update(open(os.path.join(_HERE, "..", "include", "speechdsp.h"), "rb").read())
    return h.hexdigest()


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def up_to_date() -> bool:
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(d) <= t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return OUT
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libspeechdsp.so")
    if verbose:
        sys.stderr.write(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
